"""Delta transport (csrc/spx_pack.cu "dpack"): the C host decoder against the plain-Python
statement of the format (tests/dpack_ref.py).  CPU only; the CUDA encoder is checked
against the same decoder in tests/test_gpu_main.py."""
import ctypes as C

import numpy as np
import pytest

from spinterps_b200 import _lib
from tests import dpack_ref


def c_decode(offs, payload, n_rows, row_len, decimals, n_threads=1, row0=0):
    lib = _lib.load()
    tiles = lib.spx_dpack_segments(row_len)
    out = np.full((n_rows, row_len + 3), 7.0, dtype=np.float32)
    o = offs[row0 * tiles:]
    _lib.check(lib.spx_dunpack_rows_host(
        o.ctypes.data, payload.ctypes.data, payload.nbytes, n_rows, row_len, decimals,
        out.ctypes.data, out.strides[0] // 4, n_threads), 'dunpack')
    assert np.all(out[:, row_len:] == 7.0)
    return out[:, :row_len]


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32),
                          np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize('row_len', [1, 7, 256, 300, 1000, 8192 + 700])
@pytest.mark.parametrize('decimals', [0, 2, 3])
def test_decoder_matches_format(row_len, decimals):
    rng = np.random.default_rng(row_len * 10 + decimals)
    fld = dpack_ref.synth_field(rng, 5, row_len, decimals)
    fld[1, ::5][~np.isnan(fld[1, ::5])] = -0.0
    if row_len >= 300:
        fld[2, 10] = np.float32(1.23456789)       # off the lattice: raw tile
        fld[2, 290] = np.inf
        fld[3, :] = np.float32(12.5) if decimals else np.float32(12.0)   # constant tiles
        fld[4, 260:] = rng.normal(0, 1e6, row_len - 260).astype(np.float32).round(decimals)
    offs, payload = dpack_ref.encode(fld, decimals)
    # NaN payload bits are canonical (0x7FC00000) in both decoders; compare through isnan
    want = fld.copy()
    want[np.isnan(want)] = np.float32(np.nan)
    got_py = dpack_ref.decode(offs, payload, *fld.shape, decimals)
    got_c = c_decode(offs, payload, *fld.shape, decimals)
    for got in (got_py, got_c):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        fin = ~np.isnan(want)
        assert same_bits(got[fin], want[fin])
    # threaded decode and a row offset
    got_t = c_decode(offs, payload, 3, row_len, decimals, n_threads=3, row0=2)
    fin = ~np.isnan(want[2:])
    assert same_bits(got_t[fin], want[2:][fin])


def test_rate_on_a_smooth_field():
    rng = np.random.default_rng(0)
    fld = dpack_ref.synth_field(rng, 4, 2048, 2, nan_frac=0.0)
    offs, payload = dpack_ref.encode(fld, 2)
    assert payload.nbytes + offs.nbytes < 1.0 * fld.size          # < 1 byte per cell


def test_decoder_rejects_bad_offsets():
    rng = np.random.default_rng(1)
    fld = dpack_ref.synth_field(rng, 2, 600, 2, nan_frac=0.0)
    offs, payload = dpack_ref.encode(fld, 2)
    lib = _lib.load()
    out = np.empty_like(fld)
    bad = offs.copy()
    bad[1] = 0xFFFFFFFF            # second row
    assert lib.spx_dunpack_rows_host(bad.ctypes.data, payload.ctypes.data, payload.nbytes, 2, 600, 2,
                                     out.ctypes.data, 600, 1) != 0
    assert lib.spx_dunpack_rows_host(offs.ctypes.data, payload.ctypes.data, 8, 2, 600, 2,
                                     out.ctypes.data, 600, 1) != 0


def test_decoder_special_tiles():
    """Linear ramps (second differences all zero: empty groups with a running slope),
    quadratics, jumps that need 14 / 16 / 32-bit groups, wrap-around of the int32 chain."""
    c = np.arange(1500, dtype=np.float64)
    rows = [
        0.03 * c,                                        # ramp
        -0.07 * c + 5.0,
        0.0001 * (c - 700) ** 2,                          # quadratic
        np.where((c // 9) % 2 == 0, 120.0, -95.5),       # jumps of ~2e4 lattice steps
        np.where((c // 5) % 2 == 0, 2.0e7, -2.0e7),      # jumps of 4e9: wraps modulo 2^32
        np.cumsum(np.random.default_rng(5).integers(-3000, 3000, 1500)) * 0.01,
        np.r_[np.full(700, 3.25), 0.05 * c[:800]],       # plateau, then a ramp
    ]
    fld = np.round(np.array(rows).astype(np.float32), 2)
    offs, payload = dpack_ref.encode(fld, 2)
    modes = payload[(offs.astype(np.int64) * 4)]
    assert (modes & 16).any()                            # second differences are in use
    got_c = c_decode(offs, payload, *fld.shape, 2)
    got_py = dpack_ref.decode(offs, payload, *fld.shape, 2)
    assert same_bits(got_py, fld) and same_bits(got_c, fld)


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_decoder_fuzz_all_width_classes(seed):
    """Random rows that mix every regime inside single tiles -- constant stretches, smooth
    ramps, noise of every magnitude up to the int32 range, NaN and -0.0 cells, ragged row
    length -- through the Python encoder; the C decoder (AVX2 + BMI2 where the CPU has them,
    scalar under SPX_DUNPACK_SCALAR=1) must return the identical bits."""
    rng = np.random.default_rng(100 + seed)
    row_len = 2600 + 37 * seed
    rows = []
    for _ in range(10):
        x = np.zeros(row_len)
        c = 0
        level = rng.uniform(-50, 50)
        while c < row_len:
            n = int(rng.integers(1, 400))
            kind = rng.integers(0, 5)
            seg = np.full(n, level)
            if kind == 1:
                seg = level + np.cumsum(rng.uniform(-0.5, 0.5) + np.zeros(n))
            elif kind == 2:
                seg = level + rng.normal(0, 10.0 ** rng.integers(-2, 7), n)
            elif kind == 3:
                seg = level + 1e-3 * (np.arange(n) - n / 2) ** 2
            elif kind == 4:
                seg = rng.choice([-0.001, 0.0, 0.004, level], n)
            x[c:c + n] = seg[:row_len - c]
            level = float(x[min(c + n, row_len) - 1])
            if not np.isfinite(level) or abs(level) > 1e7:
                level = rng.uniform(-50, 50)
            c += n
        x = np.round(x.astype(np.float32), 2)
        x[rng.random(row_len) < rng.choice([0.0, 0.02, 0.3])] = np.nan
        rows.append(x)
    fld = np.array(rows, dtype=np.float32)
    offs, payload = dpack_ref.encode(fld, 2)
    got = c_decode(offs, payload, *fld.shape, 2, n_threads=2)
    ok = ~np.isnan(fld)
    assert np.array_equal(np.isnan(got), ~ok)
    assert same_bits(got[ok], fld[ok])
    modes = set()
    at = 0
    for m in payload[offs.astype(np.int64) * 4]:
        modes.add(int(m) & 3)
    assert modes                                             # at least the segment heads parsed
