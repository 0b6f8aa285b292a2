"""Plain-Python statement of the delta transport format (csrc/spx_pack.cu, "dpack"): an
encoder and a decoder written from the record description, used to check the C host decoder
on CPU and the CUDA encoder on the GPU.  Test infrastructure only."""
import numpy as np

TILE = 256
SEGMENT = 8192
WIDTHS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 32]


def _code_of_width(w):
    for c, ww in enumerate(WIDTHS):
        if w <= ww:
            return c
    raise AssertionError


def _lattice(v, p):
    """q and the round-trip verdict of one float32 value (not NaN, not -0.0)."""
    qf = np.rint(np.float32(v) * p)
    if not abs(float(qf)) < 2147483520.0:
        return 0, False
    q = int(qf)
    return q, (np.float32(q) / p).view(np.uint32) == np.float32(v).view(np.uint32)


def _zigzag(d):
    return ((d << 1) ^ (d >> 31)) & 0xFFFFFFFF


def _wrap(d):
    return (d + 2 ** 31) % 2 ** 32 - 2 ** 31


def _groups(z):
    """(codes of the 32 groups, payload bytes) of 256 zigzag values."""
    codes, pay = [], bytearray()
    for l in range(32):
        zz = [int(v) for v in z[l * 8:(l + 1) * 8]]
        c = _code_of_width(max(zz).bit_length())
        w = WIDTHS[c]
        codes.append(c)
        acc = 0
        for j, v in enumerate(zz):
            acc |= v << (j * w)
        pay += acc.to_bytes(w, 'little')
    return codes, pay


def _size(codes, pay):
    return len(pay) + (sum(1 for c in codes if c) + 1) // 2


def encode_tile(x, p):
    """Record (bytes) of one tile; x float32 [n <= 256]."""
    n = x.size
    bits = x.view(np.uint32)
    is_nan = np.isnan(x)
    is_nz = bits == 0x80000000
    q = np.zeros(TILE, dtype=np.int64)
    valid = np.zeros(TILE, dtype=bool)
    bad = False
    for c in range(n):
        if is_nan[c]:
            continue
        if is_nz[c]:
            valid[c] = True
            continue
        q[c], ok = _lattice(x[c], p)
        if ok:
            valid[c] = True
        else:
            bad = True
    raw = np.zeros(TILE, dtype=np.float32)
    raw[:n] = x
    raw_rec = bytes([2]) + raw.tobytes()
    if bad:
        return raw_rec
    if not valid.any():
        return bytes([0])
    base = int(q[np.argmax(valid)])
    f = base
    d1 = np.zeros(TILE, dtype=np.int64)
    for c in range(TILE):
        if valid[c]:
            d1[c] = _wrap(int(q[c]) - f)
            f = int(q[c])
    z = [_zigzag(int(d)) for d in d1]
    codes, pay = _groups(z)
    order2 = 0
    if valid.all():                                   # second differences compete
        d2 = [_wrap(int(d1[c]) - (int(d1[c - 1]) if c else 0)) for c in range(TILE)]
        z2 = [_zigzag(d) for d in d2]
        codes2, pay2 = _groups(z2)
        if _size(codes2, pay2) < _size(codes, pay):
            codes, pay, order2 = codes2, pay2, 1
    has_nan, has_nz = bool(is_nan.any()), bool(is_nz.any())
    if len(pay) == 0 and not has_nan and not has_nz:
        return bytes([3]) + np.uint32(base & 0xFFFFFFFF).tobytes()
    b = bytearray([1 | (4 if has_nan else 0) | (8 if has_nz else 0) | (order2 << 4)])
    b += np.uint32(base & 0xFFFFFFFF).tobytes()
    nzg = sum(1 << l for l in range(32) if codes[l])
    b += np.uint32(nzg).tobytes()
    nz_codes = [c for c in codes if c] + [0]
    b += bytes(nz_codes[2 * i] | (nz_codes[2 * i + 1] << 4) for i in range((len(nz_codes) - 1 + 1) // 2))
    for flag, m in ((has_nan, is_nan), (has_nz, is_nz)):
        if flag:
            mm = np.zeros(TILE, dtype=bool)
            mm[:n] = m
            b += np.packbits(mm, bitorder='little').tobytes()
    b += pay
    return raw_rec if len(b) >= len(raw_rec) else bytes(b)


def encode(fld, decimals):
    """fld float32 [n_rows, row_len] (rounded) -> (seg_off uint32 [n_rows * segments],
    payload uint8); segments in row order."""
    fld = np.ascontiguousarray(fld, dtype=np.float32)
    n_rows, row_len = fld.shape
    p = np.float32(10.0 ** decimals)
    segs = (row_len + SEGMENT - 1) // SEGMENT
    offs = np.zeros(n_rows * segs, dtype=np.uint32)
    out = bytearray()
    for r in range(n_rows):
        for s in range(segs):
            offs[r * segs + s] = len(out) // 4
            for c0 in range(s * SEGMENT, min(row_len, (s + 1) * SEGMENT), TILE):
                out += encode_tile(fld[r, c0:c0 + TILE], p)
                out += bytes((-len(out)) % 4)             # every record padded to 4 bytes
    return offs, np.frombuffer(bytes(out), dtype=np.uint8).copy()


def segment_sizes(fld, decimals):
    """4-byte words every segment needs, in (row, segment) order."""
    offs, payload = encode(fld, decimals)
    return np.diff(np.append(offs.astype(np.int64), payload.size // 4))


def decode(offs, payload, n_rows, row_len, decimals):
    """The inverse, in plain Python (small fields only)."""
    p = np.float32(10.0 ** decimals)
    segs = (row_len + SEGMENT - 1) // SEGMENT
    b = payload
    out = np.empty((n_rows, row_len), dtype=np.float32)
    for r in range(n_rows):
        for s in range(segs):
            at = int(offs[r * segs + s]) * 4
            for c0 in range(s * SEGMENT, min(row_len, (s + 1) * SEGMENT), TILE):
                n = min(TILE, row_len - c0)
                o = out[r, c0:c0 + n]
                mode = int(b[at])
                kind = mode & 3
                if kind == 0:
                    o[:] = np.nan
                    at += 4
                    continue
                if kind == 2:
                    o[:] = np.frombuffer(bytes(b[at + 1:at + 1 + 4 * n]), dtype=np.float32)
                    at += 4 + 4 * TILE
                    continue
                f = int(np.frombuffer(bytes(b[at + 1:at + 5]), dtype=np.int32)[0])
                if kind == 3:
                    o[:] = np.float32(f) / p
                    at += 8
                    continue
                nzg = int(np.frombuffer(bytes(b[at + 5:at + 9]), dtype=np.uint32)[0])
                n_nz = bin(nzg).count('1')
                nib = b[at + 9:at + 9 + (n_nz + 1) // 2]
                pos = at + 9 + (n_nz + 1) // 2
                bm_nan = bm_nz = None
                if mode & 4:
                    bm_nan = np.unpackbits(b[pos:pos + 32], bitorder='little').astype(bool)
                    pos += 32
                if mode & 8:
                    bm_nz = np.unpackbits(b[pos:pos + 32], bitorder='little').astype(bool)
                    pos += 32
                d1 = 0
                k = 0
                for l in range(32):
                    w = 0
                    if (nzg >> l) & 1:
                        w = WIDTHS[(int(nib[k // 2]) >> ((k & 1) * 4)) & 15]
                        k += 1
                    acc = int.from_bytes(bytes(b[pos:pos + w]), 'little')
                    pos += w
                    for j in range(8):
                        cell = l * 8 + j
                        if cell >= n:
                            break
                        z = (acc >> (j * w)) & ((1 << w) - 1) if w else 0
                        d = (z >> 1) ^ -(z & 1)
                        if mode & 16:
                            d1 = _wrap(d1 + d)
                            f = _wrap(f + d1)
                        else:
                            f = _wrap(f + d)
                        if bm_nan is not None and bm_nan[cell]:
                            o[cell] = np.nan
                        elif bm_nz is not None and bm_nz[cell]:
                            o[cell] = -0.0
                        else:
                            o[cell] = np.float32(f) / p
                at = (pos + 3) // 4 * 4
    return out


def synth_field(rng, n_rows, row_len, decimals=2, nan_frac=0.05, smooth=True):
    """A rounded float32 test field: smooth rows with plateaus, NaN patches, -0.0, and a few
    values that are not on the lattice."""
    x = np.linspace(0, 6 * np.pi, row_len)
    fld = np.empty((n_rows, row_len), dtype=np.float32)
    for r in range(n_rows):
        row = 20.0 * np.sin(x * rng.uniform(0.2, 3.0) + rng.uniform(0, 6)) * rng.uniform(0, 1)
        if not smooth:
            row = row + rng.normal(0, 50, row_len)
        row[np.abs(row) < 3.0] = 0.0                        # plateaus
        fld[r] = np.round(row.astype(np.float32), decimals)
    if nan_frac > 0:
        m = rng.random((n_rows, row_len)) < nan_frac
        fld[m] = np.nan
        if row_len > 700:
            fld[0, 256:512] = np.nan                        # an all-NaN tile
            fld[-1, :300] = np.nan                          # leading NaN run across a tile
    return fld
