"""The output container (reference interp/prepare.py:308-431) as
the self-contained NetCDF-4 / HDF5 writer produces it, checked from the BYTES by the
reader of nc4file.py: signature and superblock, dimensions as dimension scales, dtypes
('i8' time), (1, ny, nx) chunks, shuffle + deflate filters with the requested level,
DIMENSION_LIST references, every text attribute; threaded compression; re-opening;
grid-row writes; and the raw bytes of a chunk decoded with nothing but zlib."""
import struct
import zlib

import numpy as np
import pandas as pd
import pytest

from spinterps_b200 import ncwriter
from spinterps_b200.nc4file import Nc4Reader, Nc4Writer, SIG

pytestmark = pytest.mark.skipif(ncwriter.have_netcdf4(), reason='netCDF4 backend in use')

SETT = ['sett_index_type', 'sett_stns_min_dist_thrsh', 'sett_drft_rass', 'sett_idw_exps',
        'sett_ork_flag', 'sett_spk_flag', 'sett_edk_flag', 'sett_idw_flag', 'sett_nnb_flag',
        'sett_interp_flag_est_vars', 'sett_out_dir', 'sett_cell_size', 'sett_tbeg', 'sett_tend',
        'sett_tfreq', 'sett_algn_ras', 'sett_poly_shp', 'sett_ipoly_flag', 'sett_stn_bdist',
        'sett_cell_bdist', 'sett_poly_simplify_tol_ratio', 'sett_min_var_thr', 'sett_min_var_cut',
        'sett_max_var_cut', 'sett_max_steps_per_chunk', 'sett_min_vg_val', 'sett_neb_sel_mthd',
        'sett_n_nebs', 'sett_n_pies']       # interp/prepare.py:384-427


def _make(tmp_path, nt=150, ny=9, nx=13, level=3, dtype=np.float32):
    x = 500.0 + 1000.0 * np.arange(nx)
    y = 500.0 + 1000.0 * np.arange(ny)[::-1]
    tv = ncwriter.time_numbers(pd.date_range('2000-01-01', periods=nt, freq='D'),
                               'days since 1900-01-01', 'gregorian', 'D')
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)]
    sett = {k: f'v_{i}' for i, k in enumerate(SETT)}
    p = ncwriter.create(tmp_path / 'out.nc', x, y, tv, args, dtype, 'mm', 'precip',
                        'days since 1900-01-01', 'gregorian', level, sett)
    return p, x, y, tv, sett


def test_container_layout_from_the_bytes(tmp_path):
    p, x, y, tv, sett = _make(tmp_path)
    raw = open(p, 'rb').read()
    assert raw[:8] == SIG and raw[8] == 0                         # HDF5, superblock version 0
    assert struct.unpack_from('<Q', raw, 40)[0] == len(raw)       # end-of-file address
    f = Nc4Reader(p)
    # dimensions are dimension scales without a variable of their own
    assert f.dimensions == {'dimx': 13, 'dimy': 9, 'dimt': 150}
    for i, d in enumerate(('dimx', 'dimy', 'dimt')):
        a = f.datasets[d]['attrs']
        assert a['CLASS'] == 'DIMENSION_SCALE' and a['_Netcdf4Dimid'] == i
        assert a['NAME'] == ('This is a netCDF dimension but not a netCDF variable.%10d'
                             % f.dimensions[d])
    # coordinate variables: 'd', 'd', 'i8' (interp/prepare.py:325-354)
    assert f.datasets['X']['dtype'] == np.dtype('<f8') and f.datasets['X']['dims'] == ('dimx',)
    assert f.datasets['Y']['dtype'] == np.dtype('<f8') and f.datasets['Y']['dims'] == ('dimy',)
    assert f.datasets['time']['dtype'] == np.dtype('<i8') and f.datasets['time']['dims'] == ('dimt',)
    assert np.array_equal(f.read_var('X'), x) and np.array_equal(f.read_var('Y'), y)
    assert np.array_equal(f.read_var('time'), tv)
    assert f.datasets['time']['attrs'] == {'units': 'days since 1900-01-01', 'calendar': 'gregorian'}
    # fields: float32 (dimt, dimy, dimx), chunks (1, ny, nx), shuffle + zlib level 3
    for lab, std in (('OK', 'precip (OK)'), ('IDW_000', 'precip (IDW_exp_2.0)')):
        d = f.datasets[lab]
        assert d['dtype'] == np.dtype('<f4') and d['shape'] == (150, 9, 13)
        assert d['dims'] == ('dimt', 'dimy', 'dimx')
        assert d['layout'] == 'chunked' and d['chunk'] == (1, 9, 13)
        assert d['filters'] == [(2, (4,)), (1, (3,))]              # shuffle(4 bytes), deflate(3)
        assert d['attrs'] == {'units': 'mm', 'standard_name': std}
        assert d['chunks'] == {}                                  # nothing written yet
    # the 29 settings, Source, netCDF-4 provenance
    assert len(SETT) == 29
    for k in SETT:
        assert f.root_attrs[k] == sett[k]
    assert f.root_attrs['Source'] == str(p) and '_NCProperties' in f.root_attrs
    f.close()


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_threaded_chunk_writes_reopen_and_row_chunks(tmp_path, dtype):
    nt, ny, nx = 150, 9, 13
    p, *_ = _make(tmp_path, nt, ny, nx, level=1, dtype=dtype)
    rng = np.random.default_rng(1)
    full = np.round(rng.gamma(1.0, 5.0, size=(nt, ny, nx)), 2).astype(dtype)
    full[7, 2:4] = np.nan
    h = ncwriter.open_for_update(p)
    h.write('OK', slice(0, 100), 0, ny, full[:100])              # slabs of whole steps
    h.close()
    ncwriter.finalize(p)                                          # file closed ...
    h = ncwriter.open_for_update(p)                               # ... and re-opened
    for t in range(100, nt):                                      # step by step (vgs path)
        h.write('OK', t, 0, ny, full[t])
    h.write('IDW_000', slice(3, 6), 0, 4, full[3:6, 0:4])         # grid-row chunks
    h.write('IDW_000', slice(3, 6), 4, ny, full[3:6, 4:ny])
    assert np.array_equal(h.read('OK', 120), full[120])
    h.sync()
    h.close()
    ncwriter.finalize(p)
    f = Nc4Reader(p)
    assert sorted(f.datasets['OK']['chunks']) == list(range(nt))   # 3-level chunk B-tree
    for t in range(nt):
        assert np.array_equal(f.read_step('OK', t), full[t], equal_nan=True)
    for t in range(nt):
        got = f.read_step('IDW_000', t)
        if 3 <= t < 6:
            assert np.array_equal(got, full[t], equal_nan=True)
        else:
            assert np.isnan(got).all()
    # a chunk decoded with nothing but zlib + byte un-shuffle from its raw bytes
    addr, nbytes = f.datasets['OK']['chunks'][42]
    blob = open(p, 'rb').read()[addr:addr + nbytes]
    plain = np.frombuffer(zlib.decompress(blob), dtype=np.uint8)
    isz = np.dtype(dtype).itemsize
    vals = plain.reshape(isz, ny * nx).T.copy().view(np.dtype(dtype).newbyteorder('<')).reshape(ny, nx)
    assert np.array_equal(vals, full[42])
    f.close()


def test_packed_field_rows_go_straight_to_the_compressor(tmp_path):
    """The writer accepts the 2-byte transport form (transfer.PackedField): each step is
    decoded inside the compression worker; same file bytes as writing the float field."""
    from spinterps_b200 import _lib
    from spinterps_b200.transfer import PackedField
    nt, ny, nx = 12, 9, 13
    rng = np.random.default_rng(2)
    q = rng.integers(-50, 6000, size=(nt, ny * nx)).astype(np.int32)
    stride = (ny * nx + 7) // 8 * 8
    hdr = np.zeros(nt, dtype=_lib.PACK_ROW_DTYPE)
    hdr['qmin'] = q.min(axis=1)
    codes = np.zeros((nt, stride), dtype=np.uint16)
    codes[:, :ny * nx] = (q - hdr['qmin'][:, None]).astype(np.uint16)
    codes[5, 17] = 0xFFFF
    fld = q.astype(np.float32) / np.float32(100.0)
    fld[5, 17] = np.nan
    pf = PackedField(_lib.load(), hdr, codes, {}, ny * nx, 2)
    assert np.array_equal(pf.decode(), fld, equal_nan=True)
    (tmp_path / 'a').mkdir()
    (tmp_path / 'b').mkdir()
    pa, *_ = _make(tmp_path / 'a', nt, ny, nx)
    pb, *_ = _make(tmp_path / 'b', nt, ny, nx)
    wa = Nc4Writer(pa, 'r+', n_threads=3)
    wa.write_steps('OK', 0, pf)
    wa.close()
    wb = Nc4Writer(pb, 'r+', n_threads=3)
    wb.write_steps('OK', 0, fld.reshape(nt, ny, nx))
    wb.close()
    fa, fb = Nc4Reader(pa), Nc4Reader(pb)
    for t in range(nt):
        assert np.array_equal(fa.read_step('OK', t), fb.read_step('OK', t), equal_nan=True)
        assert np.array_equal(fa.read_step('OK', t).ravel(), fld[t], equal_nan=True)
    fa.close()
    fb.close()


def test_delta_field_rows_go_straight_to_the_compressor(tmp_path):
    """The same for the delta transport form (transfer.DeltaField): the compression workers
    decode their step concurrently, each on its own thread."""
    from spinterps_b200 import _lib
    from spinterps_b200.transfer import DeltaField
    from tests import dpack_ref
    nt, ny, nx = 12, 9, 37
    rng = np.random.default_rng(3)
    fld = dpack_ref.synth_field(rng, nt, ny * nx, 2)
    fld[4, 50:60] = -0.0
    offs, payload = dpack_ref.encode(fld, 2)
    df = DeltaField(_lib.load(), offs, payload, nt, ny * nx, 2)
    assert np.array_equal(df.decode().view(np.uint32)[~np.isnan(fld)],
                          fld.view(np.uint32)[~np.isnan(fld)])
    assert np.array_equal(df.row(7), fld[7], equal_nan=True)
    pa, *_ = _make(tmp_path, nt, ny, nx)
    wa = Nc4Writer(pa, 'r+', n_threads=4)
    wa.write_steps('OK', 0, df)
    wa.close()
    fa = Nc4Reader(pa)
    for t in range(nt):
        assert np.array_equal(fa.read_step('OK', t).ravel(), fld[t], equal_nan=True)
    fa.close()
