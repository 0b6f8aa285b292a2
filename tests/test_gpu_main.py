"""End to end through the reference-facing surface on the GPU: SpInterpSteps
(operator seam, the reference's 12-tuple in / 13-tuple out) and SpInterpMain
(setters -> verify -> interpolate -> output file + stats.csv)."""
import types
from threading import Lock

import numpy as np
import pandas as pd
import pytest

from oracle import spinterp_oracle as orc
from tests.golden_util import load_case, rel_err

pytestmark = pytest.mark.gpu


def test_steps_seam_matches_golden():
    from spinterps_b200.steps import SpInterpSteps
    case, outs = load_case('d_sk_ok_mask_rows')
    n_stn = case['stn_xs'].size
    labels = [f'S{i:05d}' for i in range(n_stn)]
    T = case['data'].shape[0]
    tidx = pd.date_range('2000-01-01', periods=T)
    main = types.SimpleNamespace(
        _vb=False, _n_cpus=1, _mp_flag=False,
        _crds_df=pd.DataFrame({'X': case['stn_xs'], 'Y': case['stn_ys']}, index=labels),
        _min_var_thr=case['min_var_thr'], _min_var_cut=case['min_var_cut'],
        _max_var_cut=case['max_var_cut'], _cntn_idxs=case['cntn_idxs'],
        _interp_crds_orig_shape=case['grid_shape'], _interp_x_crds_msh=case['cell_xs'],
        _interp_y_crds_msh=case['cell_ys'], _nc_file_path=None, _nc_nmrl_prcn=2,
        _neb_sel_mthd='all', _n_nebs=None, _n_pies=None, _min_vg_val=case['min_vg_val'],
        _interp_flag_est_vars=False, _intrp_dtype=np.float64)
    args = (pd.DataFrame(case['data'], index=tidx, columns=labels), 0, T, 1,
            case['interp_args'], Lock(), None, None,
            pd.Series(case['vgs'], index=tidx, dtype=object), pd.Series(np.arange(T), index=tidx),
            case['fld_beg_row'], case['fld_end_row'])
    out = SpInterpSteps(main)._get_all_interp_outputs(args)
    assert len(out) == 13 and out[6] == [a[2] for a in case['interp_args']]
    for lab, ref in outs.items():
        tol = 1e-12 if lab.startswith('IDW') else 1e-9
        assert rel_err(out[7][lab], ref, 0.2) <= tol, lab


def test_main_end_to_end(tmp_path):
    from spinterps_b200 import ncwriter
    from spinterps_b200.main import SpInterpMain
    rng = np.random.default_rng(4)
    n_stn, T = 30, 9
    idx = pd.date_range('2001-03-01', periods=T, freq='D')
    labs = [f'P{i:03d}' for i in range(n_stn)]
    vals = rng.gamma(1.0, 5.0, (T, n_stn))
    vals[rng.random((T, n_stn)) < 0.15] = np.nan
    data = pd.DataFrame(vals, index=idx, columns=labs)
    crds = pd.DataFrame({'X': rng.uniform(0, 6e4, n_stn), 'Y': rng.uniform(0, 5e4, n_stn)},
                        index=labs)
    vg = '0.1 Nug(0.0) + 0.9 Sph(20000)'
    vgs_ser = pd.Series([vg] * T, index=idx, dtype=object)

    m = SpInterpMain(False)
    m.set_data(data, crds)
    m.set_vgs_ser(vgs_ser)
    m.set_out_dir(tmp_path / 'run')
    m.set_netcdf4_parameters('precip.nc', 'mm', 'precipitation', 'days since 1900-01-01',
                             'gregorian', 2, 1)
    m.set_interp_time_parameters('2001-03-01', '2001-03-09', 'D', '%Y-%m-%d')
    m.set_neighbor_selection_method('all')
    m.set_misc_settings(cell_size=2000.0, min_cutoff_value=0.0, max_steps_per_chunk=4)
    m.set_cell_selection_mask(lambda x, y: (x - 3e4) ** 2 + (y - 2.5e4) ** 2 <= 2.4e4 ** 2,
                              polygon_cell_buffer_distance=1000.0)
    m.turn_ordinary_kriging_on()
    m.turn_inverse_distance_weighting_on([2])
    m.turn_nearest_neighbor_on()
    m.verify()
    m.interpolate()

    exp, _ = orc.interp_chunk(
        m._data_df.values, m._crds_df['X'].values, m._crds_df['Y'].values,
        m._interp_x_crds_msh, m._interp_y_crds_msh, m._interp_crds_orig_shape, m._interp_args,
        vgs=[vg] * T, cntn_idxs=m._cntn_idxs, min_var_cut=0.0, intrp_dtype=np.float32)
    h = ncwriter.open_for_read(m._nc_file_path)
    ny, nx = m._interp_crds_orig_shape
    for lab in ('OK', 'IDW_000', 'NNB'):
        ref = np.round(exp[lab], 2).reshape(T, ny, nx)
        for t in range(T):
            got = h.read(lab, t)
            assert np.array_equal(np.isnan(got), np.isnan(ref[t]))
            # both sides round f32 values to 2 decimals; a value on a rounding boundary
            # may land on the neighbouring cent
            assert np.nanmax(np.abs(got - ref[t])) <= 0.0101, (lab, t)
            assert np.nanmean(np.abs(got - ref[t]) > 1e-6) < 1e-3
    h.close()
    stats = pd.read_csv(tmp_path / 'run' / 'stats.csv', sep=';', index_col=0)
    assert stats.shape[0] == T
    for lab in ('data', 'OK', 'IDW_000', 'NNB'):
        for st in ('min', 'mean', 'max', 'std', 'count'):
            assert f'{lab}_{st}' in stats.columns
    assert np.allclose(stats['OK_count'].values, m._cntn_idxs.sum())
    # the statistics come from the GPU output stage; the reference computes them from
    # the file content with numpy (interp/main.py:503-520)
    h = ncwriter.open_for_read(m._nc_file_path)
    for lab in ('OK', 'IDW_000', 'NNB'):
        for t in range(T):
            f = np.asarray(h.read(lab, t), dtype=np.float64)
            for st in ('min', 'mean', 'max', 'std'):
                exp_v = getattr(np, f'nan{st}')(f)
                assert abs(stats[f'{lab}_{st}'].values[t] - exp_v) <= 2e-6 + 1e-6 * abs(exp_v), (
                    lab, t, st)
    h.close()


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_round_stats_kernel(dtype):
    """spx_round_stats_dev: bit-exact np.round in the field dtype + per-row statistics."""
    import torch
    from spinterps_b200.engine import ChunkEngine
    eng = ChunkEngine()
    rng = np.random.default_rng(5)
    for (T, n, ld, dec) in [(7, 10007, 10007, 2), (3, 4096, 4100, 1), (5, 33, 36, 3),
                            (4, 50000, 50000, None)]:
        a = (rng.gamma(1.0, 5.0, (T, ld)) * rng.choice([1.0, 100.0], (T, 1))).astype(dtype)
        a[rng.random((T, ld)) < 0.1] = np.nan
        a[1, :] = np.nan                      # a step without any value
        a[0, 5] = 0.125                       # ties: round half to even
        a[0, 6] = 0.375
        d = torch.from_numpy(a.copy()).cuda()
        view = d[:, :n]
        st = eng.round_and_stats(view, dec)
        got = d.cpu().numpy()
        ref = a.copy()
        if dec is not None:
            ref[:, :n] = np.round(a[:, :n], dec)
        assert np.array_equal(got, ref, equal_nan=True), (T, n, ld, dec)
        f = ref[:, :n].astype(np.float64)
        with np.errstate(all='ignore'), np.testing.suppress_warnings() as sup:
            sup.filter(RuntimeWarning)
            exp = np.stack([np.nanmin(f, 1), np.nanmean(f, 1), np.nanmax(f, 1), np.nanstd(f, 1),
                            np.isfinite(f).sum(1).astype(float)])
        assert np.array_equal(np.isnan(st), np.isnan(exp))
        assert np.array_equal(st[4], exp[4])
        ok = ~np.isnan(exp)
        assert np.all(np.abs(st[ok] - exp[ok]) <= 1e-11 * np.maximum(1.0, np.abs(exp[ok])))


@pytest.mark.parametrize('codec', ['delta', 'u16'])
@pytest.mark.parametrize('decimals,scale,special', [(2, 5.0, 'nan'), (1, 30.0, 'none'),
                                                    (3, 2.0, 'wide'), (0, 100.0, 'inf'),
                                                    (2, 5.0, 'unrounded')])
def test_packed_download_is_bit_exact(decimals, scale, special, codec):
    """The compact transports of rounded f32 fields (delta: spx_dpack_field_dev -> PCIe ->
    spx_dunpack_rows_host; u16: spx_pack_field_dev -> spx_unpack_field_host) return exactly
    the bytes of the rounded field: NaN rows and
    cells, negative values, rows whose range exceeds 16 bits, infinities, a field that was
    never rounded (every row falls back to raw floats), ragged row length."""
    import torch
    from spinterps_b200 import _lib
    from spinterps_b200.engine import ChunkEngine
    from spinterps_b200.transfer import PackedDownloader
    eng = ChunkEngine()
    rng = np.random.default_rng(11)
    T, G = 37, 30011
    f = (rng.gamma(1.0, scale, size=(T, G)) - 0.3 * scale).astype(np.float32)
    if special == 'nan':
        f[rng.random(f.shape) < 0.1] = np.nan
        f[5] = np.nan
    elif special == 'wide':
        f[3, 100] = 1.0e5          # range of the row > 65534 codes -> raw
        f[7] *= 1.0e6              # |q| beyond int32 -> raw
    elif special == 'inf':
        f[2, 17] = np.inf
        f[9, 3] = -np.inf
    d = torch.from_numpy(f).cuda()
    if special != 'unrounded':
        eng.round_and_stats(d, decimals)
        exp = np.round(f, decimals)
    else:
        exp = f
    assert np.array_equal(d.cpu().numpy(), exp, equal_nan=True)
    dl = PackedDownloader(eng.device, T, G, codec=codec)
    out = np.full((T, G), -7.0, dtype=np.float32)
    n_raw = dl.finish(dl.start(d, decimals), out)
    assert np.array_equal(out.view(np.uint32)[~np.isnan(exp)],
                          exp.view(np.uint32)[~np.isnan(exp)])       # incl. the sign of zero
    assert (np.signbit(exp) & (exp == 0)).any() or special != 'nan'   # -0.0 is exercised
    assert np.array_equal(np.isnan(out), np.isnan(exp))
    if codec == 'delta':
        # raw TILES instead of raw rows; a noise-like field (gamma noise, cell by cell) does
        # not fit 1.25 bytes per cell and goes through the 16-bit codec
        assert dl.d2h_bytes > 0 and dl.fallbacks in (0, 1)
        if special == 'unrounded':
            assert dl.fallbacks == 1         # every tile raw: > 4 bytes per cell
        return
    if special == 'unrounded':
        assert n_raw >= T - 1
    elif special == 'wide':
        assert n_raw == 2
    elif special == 'inf':
        assert n_raw == 2
    else:
        assert n_raw == 0
    # and through the public result() of a chunk
    assert dl.d2h_bytes >= T * 2 * G


def _dpack(d, row_len, decimals, flags=0, stats=False, cap=None):
    """spx_dpack_field_dev on a device tensor; returns (counters, seg_off, payload, stats)."""
    import ctypes as C
    import torch
    from spinterps_b200 import _lib
    lib = _lib.load()
    n_rows, pitch = d.shape
    segs = int(lib.spx_dpack_segments(row_len))
    if cap is None:
        cap = int(lib.spx_dpack_capacity(n_rows, row_len))
    offs = torch.zeros(n_rows * segs, dtype=torch.int32, device='cuda')
    pay = torch.full((int(lib.spx_dpack_capacity(n_rows, row_len)) + 64,), 0xAB,
                     dtype=torch.uint8, device='cuda')
    cnt = torch.full((2,), -1, dtype=torch.int64, device='cuda')
    st = torch.zeros((5, n_rows), dtype=torch.float64, device='cuda') if stats else None
    ws = torch.empty(max(1, int(lib.spx_dpack_stats_workspace(n_rows, row_len))),
                     dtype=torch.uint8, device='cuda') if stats else None
    _lib.check(lib.spx_dpack_field_dev(
        C.c_void_p(d.data_ptr()), n_rows, row_len, pitch, decimals, flags,
        C.c_void_p(st.data_ptr() if stats else None), C.c_void_p(ws.data_ptr() if stats else None),
        C.c_void_p(offs.data_ptr()), C.c_void_p(pay.data_ptr()), cap, C.c_void_p(cnt.data_ptr()),
        C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'dpack')
    torch.cuda.synchronize()
    return (cnt.cpu().numpy(), offs.cpu().numpy().view(np.uint32), pay.cpu().numpy(),
            st.cpu().numpy() if stats else None)


def _c_decode(offs, used, n_rows, row_len, decimals):
    from spinterps_b200 import _lib
    out = np.empty((n_rows, row_len), dtype=np.float32)
    _lib.check(_lib.load().spx_dunpack_rows_host(
        offs.ctypes.data, used.ctypes.data, used.nbytes, n_rows, row_len, decimals,
        out.ctypes.data, row_len, 2), 'dunpack')
    return out


@pytest.mark.parametrize('row_len,pitch', [(1000, 1000), (777, 779), (256, 256), (5, 8),
                                           (8192 + 300, 8192 + 300), (3 * 8192, 3 * 8192)])
def test_delta_encoder_follows_the_format(row_len, pitch):
    """spx_dpack_field_dev against the plain-Python statement of the record format
    (tests/dpack_ref.py): the records decode to the identical floats with the Python decoder
    AND the C decoder, and every segment takes exactly the bytes the Python encoder needs
    (NaN cells and tiles, -0.0, off-lattice values -> raw tiles, constant tiles, first and
    second differences, ragged and unaligned rows); a payload buffer that is too small is
    reported, never overrun."""
    import torch
    from tests import dpack_ref
    rng = np.random.default_rng(row_len)
    n_rows = 9 if row_len < 5000 else 4
    fld = dpack_ref.synth_field(rng, n_rows, row_len, 2)
    fld[1, ::5][~np.isnan(fld[1, ::5])] = -0.0
    fld[3, :] = np.float32(12.5)
    if row_len >= 300:
        fld[2, 10] = np.float32(1.23456789)
        fld[2, 290] = np.inf
        fld[n_rows - 1, 260:] = rng.normal(0, 1e6, row_len - 260).astype(np.float32).round(2)
        fld[0, 100:140] = np.float32(3.0e9)           # |q| beyond int32
    if row_len >= 1000:
        # smooth, fully valid tiles: second differences win
        fld[2, 512:1000] = np.round(np.linspace(0, 300, 488) ** 1.5, 2).astype(np.float32)
    want_off, want_pay = dpack_ref.encode(fld, 2)
    want_words = dpack_ref.segment_sizes(fld, 2)
    d = torch.full((n_rows, pitch), 99.0, dtype=torch.float32, device='cuda')
    d[:, :row_len] = torch.from_numpy(fld).cuda()
    cnt, offs, pay, _ = _dpack(d, row_len, 2)
    assert cnt[1] == 0 and cnt[0] * 4 == want_pay.nbytes
    assert np.all(pay[cnt[0] * 4:] == 0xAB)                      # nothing past the records
    used = pay[:cnt[0] * 4].copy()
    ok = ~np.isnan(fld)
    for got in (dpack_ref.decode(offs, used, n_rows, row_len, 2),
                _c_decode(offs, used, n_rows, row_len, 2)):
        assert np.array_equal(np.isnan(got), ~ok)
        assert np.array_equal(got.view(np.uint32)[ok], fld.view(np.uint32)[ok])
    # segment by segment the same bytes as the Python encoder (their order in the payload is free)
    segs = offs.size // n_rows
    for i in range(offs.size):
        a, b = int(offs[i]) * 4, int(want_off[i]) * 4
        n = int(want_words[i]) * 4
        assert np.array_equal(used[a:a + n], want_pay[b:b + n]), (i // segs, i % segs)
    if row_len >= 1000:
        assert (want_pay[[int(o) * 4 for o in want_off]] & 3 != 0).any()
    # too small a buffer: flagged, the needed size still reported, no write past the end
    small = (want_pay.nbytes // 2) // 4 * 4
    cnt, offs2, pay2, _ = _dpack(d, row_len, 2, cap=small)
    assert cnt[0] * 4 == want_pay.nbytes
    if offs.size > 1:
        assert cnt[1] == 1 and (offs2 == 0xFFFFFFFF).any()
    assert np.all(pay2[small:] == 0xAB)


@pytest.mark.parametrize('decimals,write_back', [(2, False), (2, True), (0, False), (3, True)])
def test_fused_output_stage_matches_the_separate_kernels(decimals, write_back):
    """SPX_DPACK_ROUND: one pass over the UNROUNDED field = np.round (bit-exact, incl. -0.0,
    NaN, inf and values beyond int32) + the per-step statistics of spx_round_stats_dev + the
    encoding."""
    import torch
    from spinterps_b200 import _lib
    from spinterps_b200.engine import ChunkEngine
    eng = ChunkEngine()
    rng = np.random.default_rng(decimals)
    T, G = 21, 20011
    x = np.linspace(0, 40, G)
    f = np.stack([8 * np.sin(x * rng.uniform(0.5, 2)) * rng.uniform(0, 2) + rng.uniform(-1, 4)
                  for _ in range(T)]).astype(np.float32)
    f[np.abs(f) < 0.004] *= -1.0                      # some round to -0.0
    f[rng.random(f.shape) < 0.03] = np.nan
    f[4] = np.nan
    f[6, 100] = np.inf
    f[7, 5000:5100] = 4.0e9
    f[8, 300:900] = 0.0
    exp = np.round(f, decimals)
    d_ref = torch.from_numpy(f).cuda()
    st_ref = eng.round_and_stats(d_ref, decimals)
    assert np.array_equal(d_ref.cpu().numpy().view(np.uint32)[~np.isnan(exp)],
                          exp.view(np.uint32)[~np.isnan(exp)])
    d = torch.from_numpy(f).cuda()
    flags = _lib.SPX_DPACK_ROUND | (_lib.SPX_DPACK_WRITE_BACK if write_back else 0)
    cnt, offs, pay, st = _dpack(d, G, decimals, flags=flags, stats=True)
    assert cnt[1] == 0
    got = _c_decode(offs, pay[:cnt[0] * 4].copy(), T, G, decimals)
    ok = ~np.isnan(exp)
    assert np.array_equal(np.isnan(got), ~ok)
    assert np.array_equal(got.view(np.uint32)[ok], exp.view(np.uint32)[ok])
    assert (np.signbit(exp) & (exp == 0)).any()
    after = d.cpu().numpy()
    want_after = exp if write_back else f
    assert np.array_equal(after.view(np.uint32)[ok], want_after.view(np.uint32)[ok])
    assert np.array_equal(st[4], st_ref[4])                       # finite counts
    assert np.array_equal(st[[0, 2]], st_ref[[0, 2]], equal_nan=True)   # min / max
    rows = np.isfinite(st_ref).all(axis=0)                        # rows without inf / all-NaN
    assert rows.sum() >= T - 3
    assert np.all(np.abs(st[:, rows] - st_ref[:, rows])
                  <= 1e-11 * np.maximum(1.0, np.abs(st_ref[:, rows])))
    assert np.isnan(st[:4, 4]).all() and st[4, 4] == 0 and st_ref[4, 4] == 0


def test_delta_download_of_a_smooth_field_is_small():
    """An interpolated field (the engine's own output) crosses PCIe in well under one byte
    per cell and comes back bit for bit."""
    import torch
    from spinterps_b200.engine import ChunkEngine
    from spinterps_b200.transfer import PackedDownloader
    from tests.synth import make_problem
    p = make_problem(5, 60, 24, 120, 1100, cell=1000.0, miss=0.2)
    eng = ChunkEngine()
    pend = eng.submit_chunk(interp_args=[('OK', None, 'OK')], vgs=['0.1 Nug(0.0) + 0.9 Sph(40000)'] * 24,
                            intrp_dtype=np.float32, round_decimals=2, field_stats=True, **p)
    flds, _ = pend.result(to_host=False)
    d = flds['OK']
    dl = PackedDownloader(eng.device, d.shape[0], d.shape[1], codec='delta')
    got = dl.download(d, 2)
    exp = d.cpu().numpy()
    ok = ~np.isnan(exp)
    assert np.array_equal(np.isnan(got), ~ok)
    assert np.array_equal(got.view(np.uint32)[ok], exp.view(np.uint32)[ok])
    assert dl.fallbacks == 0 and dl.d2h_bytes < 1.0 * exp.size
