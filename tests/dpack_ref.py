"""Plain-Python statement of the delta transport format (csrc/spx_pack.cu, "dpack"): an
encoder and a decoder written from the record description, used to check the C host decoder
on CPU and the CUDA encoder on the GPU.  Test infrastructure only."""
import numpy as np

TILE = 256
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


def encode(fld, decimals):
    """fld float32 [n_rows, row_len] -> (tile_off uint32 [n_rows * tiles], payload uint8)."""
    fld = np.ascontiguousarray(fld, dtype=np.float32)
    n_rows, row_len = fld.shape
    p = np.float32(10.0 ** decimals)
    tiles = (row_len + TILE - 1) // TILE
    offs = np.zeros(n_rows * tiles, dtype=np.uint32)
    words = []
    n_words = 0
    for r in range(n_rows):
        for t in range(tiles):
            x = fld[r, t * TILE:(t + 1) * TILE]
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
            rec = None
            if bad:
                raw = np.zeros(TILE, dtype=np.float32)
                raw[:n] = x
                rec = np.concatenate([np.array([2], np.uint32), raw.view(np.uint32)])
            elif not valid.any():
                rec = np.array([0], np.uint32)
            else:
                base = int(q[np.argmax(valid)])
                f = base
                z = np.zeros(TILE, dtype=np.uint64)
                for c in range(TILE):
                    d = 0
                    if valid[c]:
                        d = (int(q[c]) - f + 2 ** 31) % 2 ** 32 - 2 ** 31     # wrap to int32
                        f = int(q[c])
                    z[c] = ((d << 1) ^ (d >> 31)) & 0xFFFFFFFF
                codes = []
                pay = bytearray()
                for l in range(32):
                    zz = [int(v) for v in z[l * 8:(l + 1) * 8]]
                    w = max(zz).bit_length()
                    c = _code_of_width(w)
                    w = WIDTHS[c]
                    codes.append(c)
                    acc = 0
                    for j, v in enumerate(zz):
                        acc |= v << (j * w)
                    pay += acc.to_bytes(w, 'little')
                has_nan, has_nz = bool(is_nan.any()), bool(is_nz.any())
                if len(pay) == 0 and not has_nan and not has_nz:
                    rec = np.array([3, base & 0xFFFFFFFF], np.uint32)
                else:
                    b = bytearray()
                    b += np.uint32(1 | (4 if has_nan else 0) | (8 if has_nz else 0)).tobytes()
                    b += np.uint32(base & 0xFFFFFFFF).tobytes()
                    b += bytes(codes[2 * i] | (codes[2 * i + 1] << 4) for i in range(16))
                    for flag, m in ((has_nan, is_nan), (has_nz, is_nz)):
                        if flag:
                            mm = np.zeros(TILE, dtype=bool)
                            mm[:n] = m
                            b += np.packbits(mm, bitorder='little').tobytes()
                    b += pay
                    b += bytes((-len(b)) % 4)
                    if len(b) // 4 >= 1 + TILE:
                        raw = np.zeros(TILE, dtype=np.float32)
                        raw[:n] = x
                        rec = np.concatenate([np.array([2], np.uint32), raw.view(np.uint32)])
                    else:
                        rec = np.frombuffer(bytes(b), dtype=np.uint32)
            offs[r * tiles + t] = n_words
            words.append(rec)
            n_words += rec.size
    payload = np.concatenate(words) if words else np.zeros(0, np.uint32)
    return offs, payload.view(np.uint8).copy()


def decode(offs, payload, n_rows, row_len, decimals):
    """The inverse, in plain Python (small fields only)."""
    p = np.float32(10.0 ** decimals)
    tiles = (row_len + TILE - 1) // TILE
    w32 = np.frombuffer(payload.tobytes(), dtype=np.uint32)
    raw8 = payload
    out = np.empty((n_rows, row_len), dtype=np.float32)
    for r in range(n_rows):
        for t in range(tiles):
            n = min(TILE, row_len - t * TILE)
            o = out[r, t * TILE:t * TILE + n]
            at = int(offs[r * tiles + t])
            mode = int(w32[at])
            kind = mode & 3
            if kind == 0:
                o[:] = np.nan
                continue
            if kind == 2:
                o[:] = w32[at + 1:at + 1 + n].view(np.float32)
                continue
            f = int(np.int32(w32[at + 1]))
            if kind == 3:
                o[:] = np.float32(f) / p
                continue
            b = raw8[at * 4:]
            nib = b[8:24]
            pos = 24
            bm_nan = bm_nz = None
            if mode & 4:
                bm_nan = np.unpackbits(b[pos:pos + 32], bitorder='little').astype(bool)
                pos += 32
            if mode & 8:
                bm_nz = np.unpackbits(b[pos:pos + 32], bitorder='little').astype(bool)
                pos += 32
            for l in range((n + 7) // 8):
                c = (int(nib[l // 2]) >> ((l & 1) * 4)) & 15
                w = WIDTHS[c]
                acc = int.from_bytes(bytes(b[pos:pos + w]), 'little')
                pos += w
                for j in range(min(8, n - l * 8)):
                    z = (acc >> (j * w)) & ((1 << w) - 1) if w else 0
                    d = (z >> 1) ^ -(z & 1)
                    f = (f + d + 2 ** 31) % 2 ** 32 - 2 ** 31
                    cell = l * 8 + j
                    if bm_nan is not None and bm_nan[cell]:
                        o[cell] = np.nan
                    elif bm_nz is not None and bm_nz[cell]:
                        o[cell] = -0.0
                    else:
                        o[cell] = np.float32(f) / p
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
