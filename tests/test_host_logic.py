"""Host-side logic (no GPU): grouping indices bit-exact vs the oracle, variogram
helpers, sharding, output-file layout, configuration surface."""
import numpy as np
import pandas as pd
import pytest

from oracle import spinterp_oracle as orc
from spinterps_b200 import dist as sdist
from spinterps_b200 import ncwriter, vgs
from spinterps_b200.engine import availability_groups
from spinterps_b200.main import SpInterpMain


@pytest.mark.parametrize('seed,T,N,miss', [(0, 50, 6, 0.4), (1, 300, 40, 0.1), (2, 20, 3, 0.9),
                                           (3, 7, 9, 0.0)])
def test_availability_groups_bit_exact(seed, T, N, miss):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(T, N))
    d[rng.random((T, N)) < miss] = np.nan
    grp_of_step, grp_mask = availability_groups(~np.isnan(d))
    ref = orc.get_grps_in_time(d)          # interp/grps.py:57-101 restated
    assert len(ref) == grp_mask.shape[0]
    for k, (idx, mask) in enumerate(ref):
        assert np.array_equal(np.where(grp_mask[k])[0], idx)
        assert np.array_equal(grp_of_step == k, mask)


def test_nuggetness_and_cluster_match_oracle():
    cases = ['0.0 Nug(0.0)', '0.1 Nug(0.0) + 0.9 Sph(20000)', '0.00001 Nug(0.0) + 0.00002 Sph(20000)',
             '1.0 Exp(0.00001)', '0.5 Sph(10)+0.5 Gau(20)', 'nan']
    for c in cases:
        for mv in (0.0, 1e-4, 0.5):
            assert vgs.check_full_nuggetness(c, mv) == orc.check_full_nuggetness(c, mv), (c, mv)
    seq = ['a', 'b', 'a', 'c', 'b', 'a']
    got, exp = vgs.get_vgs_cluster(seq), orc.get_vgs_cluster(seq)
    assert list(got) == list(exp) == ['a', 'b', 'c']
    for k in got:
        assert np.array_equal(got[k], exp[k])


def test_vg_abs_bound_is_an_upper_bound():
    h = np.linspace(0, 1e5, 2001)
    for vg in ['0.1 Nug(0.0) + 0.9 Sph(20000)', '0.3 Nug(0.0) + 0.7 Hol(25000)',
               '0.1 Nug(0.0) + 0.002 Pow(0.5)', '2.5 Rng(1.0)', '1.0 Exp(20000) + 0.5 Lin(1000)']:
        vals = orc.get_theo_vg_vals(vg, h)
        assert np.nanmax(np.abs(vals)) <= vgs.vg_abs_bound(vg, 1e5) * (1 + 1e-12), vg


def test_shard_bounds():
    b = sdist.shard_bounds(10000, 8)
    assert b[0] == 0 and b[-1] == 10000 and np.all(np.diff(b) == 1250)
    b = sdist.shard_bounds(10, 4)
    assert np.array_equal(b, np.linspace(0, 10, 5, dtype=np.int64))    # ret_mp_idxs, misc.py:601
    b = sdist.shard_bounds(3, 8)                                      # more ranks than steps
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0)
    w = np.r_[np.full(100, 9.0), np.full(900, 1.0)]
    b = sdist.shard_bounds(1000, 2, w)
    assert b[0] == 0 and b[-1] == 1000 and abs(w[:b[1]].sum() - w[b[1]:].sum()) <= 9.0


def test_output_file_layout(tmp_path):
    """interp/prepare.py:290-432: dimensions, coordinate variables, per-label
    variables and sett_* attributes."""
    x = np.linspace(500, 9500, 10)
    y = np.linspace(7500, 500, 8)
    tr = pd.date_range('2000-01-01', periods=5, freq='D')
    tv = ncwriter.time_numbers(tr, 'days since 1900-01-01', 'gregorian', 'D')
    assert tv[0] == 36524 and np.all(np.diff(tv) == 1)
    assert np.array_equal(ncwriter.time_numbers(pd.date_range('2000-01-01', periods=3, freq='2D'),
                                                'days since 2000-01-01', 'gregorian', '2D'),
                          [0, 1, 2])   # divided by the numeric prefix of the frequency
    args = [('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0)]
    p = ncwriter.create(tmp_path / 'out.nc', x, y, tv, args, np.float32, 'mm', 'precip',
                        'days since 1900-01-01', 'gregorian', 1, {'sett_cell_size': 1000.0})
    h = ncwriter.open_for_update(p)
    fld = np.arange(16, dtype=np.float32).reshape(2, 8)
    h.write('OK', 2, 3, 5, np.tile(fld[:, :1], (1, 10)))
    h.sync()
    h.close()
    if ncwriter.have_netcdf4():
        pytest.skip('netCDF4 backend: layout checked by the netCDF4 library itself')
    ncwriter.finalize(p)
    from spinterps_b200.nc4file import Nc4Reader
    f = Nc4Reader(p)
    assert f.dimensions == {'dimx': 10, 'dimy': 8, 'dimt': 5}
    assert np.array_equal(f.read_var('X'), x) and np.array_equal(f.read_var('Y'), y)
    assert f.read_var('Y')[0] > f.read_var('Y')[-1]              # Y descending
    assert f.datasets['OK']['dims'] == ('dimt', 'dimy', 'dimx')
    assert f.datasets['OK']['attrs']['standard_name'] == 'precip (OK)'
    assert f.datasets['IDW_000']['attrs']['standard_name'] == 'precip (IDW_exp_2.0)'
    assert f.datasets['time']['attrs']['units'] == 'days since 1900-01-01'
    assert f.root_attrs['sett_cell_size'] == '1000.0'
    assert np.isnan(f.read_step('OK', 0)).all()
    assert np.array_equal(f.read_step('OK', 2)[3:5, 0], [0, 8])
    f.close()


def _frames(n_stn=6, T=4):
    rng = np.random.default_rng(0)
    idx = pd.date_range('2000-01-01', periods=T, freq='D')
    labs = [f'S{i}' for i in range(n_stn)]
    data = pd.DataFrame(rng.gamma(1, 5, (T, n_stn)), index=idx, columns=labs)
    crds = pd.DataFrame({'X': rng.uniform(0, 1e4, n_stn), 'Y': rng.uniform(0, 1e4, n_stn)},
                        index=labs)
    return data, crds


def test_main_configuration_surface(tmp_path):
    """Setter validation and call-order flags behave like interp/data.py."""
    m = SpInterpMain(False)
    with pytest.raises(AssertionError):
        m.interpolate()                                   # verify() first (main.py:76)
    with pytest.raises(AssertionError):
        m.verify()                                        # set_data first (data.py:749)
    data, crds = _frames()
    with pytest.raises(AssertionError):
        m.set_data(data.values, crds)                     # not a DataFrame
    m.set_data(data, crds)
    m.set_out_dir(tmp_path / 'o')
    with pytest.raises(AssertionError):
        m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', -1, 1)
    m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', 2, 1)
    with pytest.raises(AssertionError):
        m.set_interp_time_parameters('2000-01-05', '2000-01-01', 'D', '%Y-%m-%d')
    m.set_interp_time_parameters('2000-01-01', '2000-01-04', 'D', '%Y-%m-%d')
    with pytest.raises(AssertionError):
        m.set_neighbor_selection_method('bogus')
    with pytest.raises(AssertionError):
        m.set_neighbor_selection_method('nrst')           # n_neighbors missing
    m.set_neighbor_selection_method('all')
    with pytest.raises(AssertionError):
        m.set_misc_settings(min_vg_val=1)                 # must be a float (data.py:725)
    with pytest.raises(AssertionError):
        m.set_misc_settings(min_cutoff_value=5.0, max_cutoff_value=1.0)
    m = SpInterpMain(False)
    m.set_data(data, crds)
    m.set_out_dir(tmp_path / 'o')
    m.set_netcdf4_parameters('o.nc', 'mm', 'p', 'days since 1900-01-01', 'gregorian', 2, 1)
    m.set_interp_time_parameters('2000-01-01', '2000-01-04', 'D', '%Y-%m-%d')
    m.set_neighbor_selection_method('all')
    m.set_misc_settings(cell_size=1000.0)
    with pytest.raises(AssertionError):
        m.verify()                                        # no interpolation method turned on
    with pytest.raises(AssertionError):
        m.turn_ordinary_kriging_est_var_on()              # needs OK first (main.py:440)
    m.turn_inverse_distance_weighting_on([1, 2.5])
    m.turn_nearest_neighbor_on()
    m.verify()
    assert [a[2] for a in m._interp_args] == ['IDW_000', 'IDW_001', 'NNB']
    assert m._interp_crds_orig_shape[0] * m._interp_crds_orig_shape[1] == m._interp_x_crds_msh.size
    assert m._nc_y_crds[0] > m._nc_y_crds[-1]
    assert (tmp_path / 'o' / 'o.nc').exists()


# ---- native host planner (csrc/spx_plan.cu) against the NumPy index logic ----------
@pytest.mark.parametrize('seed,T,N,miss', [(0, 40, 7, 0.3), (1, 300, 65, 0.1), (2, 64, 128, 0.0),
                                           (3, 500, 201, 0.02), (4, 50, 3, 0.7)])
def test_native_avail_groups_match_numpy(seed, T, N, miss):
    from spinterps_b200 import _lib
    rng = np.random.default_rng(seed)
    d = rng.gamma(1.0, 5.0, size=(T, N))
    d[rng.random((T, N)) < miss] = np.nan
    d[T // 2] = d[0]                       # a repeated availability pattern
    if T > 45:
        d[7] = np.nan                      # a step without stations
    thr = 4.0
    avail = ~np.isnan(d)
    exp_gos, exp_mask = availability_groups(avail)
    copy = np.full(T * N, -1.0)
    gos, mask, grp_n, first, n_avail, flag = _lib.avail_groups(d, thr, data_copy=copy.ctypes.data)
    assert np.array_equal(gos, exp_gos) and gos.dtype == np.int32
    assert np.array_equal(mask, exp_mask) and mask.dtype == np.bool_
    assert np.array_equal(grp_n, exp_mask.sum(axis=1))
    assert np.array_equal(n_avail, avail.sum(axis=1))
    assert all(gos[f] == g and not (gos[:f] == g).any() for g, f in enumerate(first))
    with np.errstate(invalid='ignore'):
        assert np.array_equal(flag, (np.where(avail, d, -np.inf) >= thr).any(axis=1))
    assert np.array_equal(copy.reshape(T, N), d, equal_nan=True)
    # -inf threshold: every step with a station passes; strided input (row pitch > N)
    wide = np.full((T, N + 5), 7.0)
    wide[:, :N] = d
    gos2, mask2, _, _, n2, flag2 = _lib.avail_groups(wide[:, :N])
    assert np.array_equal(gos2, exp_gos) and np.array_equal(mask2, exp_mask)
    assert np.array_equal(flag2, n2 >= 1)
    # packed words instead of byte masks
    gos3, bits3, n3, _, _, _ = _lib.avail_groups(d, want_mask=False)
    assert bits3.dtype == np.uint64 and bits3.shape == (exp_mask.shape[0], (N + 63) // 64)
    assert np.array_equal(_lib.unpack_group_bits(bits3, N), exp_mask)
    assert np.array_equal(gos3, exp_gos) and np.array_equal(n3, exp_mask.sum(axis=1))


def _numpy_downdate_plan(grp_of_step, grp_mask, steps, rows):
    """The index arrays engine._solve_downdate builds with NumPy (one variogram)."""
    n_stn = grp_mask.shape[1]
    grp_n = grp_mask.sum(axis=1)
    grps = np.unique(grp_of_step[steps])
    sys_of_row = np.searchsorted(grps, grp_of_step[steps])
    r = (n_stn - grp_n[grps]).astype(np.int32)
    ridx = np.argsort(sys_of_row, kind='stable')
    cnt = np.bincount(sys_of_row, minlength=grps.size)
    beg = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    nsys, n_data = grps.size, steps.size
    rhs_off = beg + np.arange(nsys)
    n_rhs = n_data + nsys
    urow = np.empty(n_rhs, dtype=np.int32)
    rrow = np.empty(n_rhs, dtype=np.int64)
    rkind = np.zeros(n_rhs, dtype=np.int32)
    pos_data = np.arange(n_data) + np.repeat(np.arange(nsys), cnt)
    urow[pos_data] = np.arange(n_data)
    rrow[pos_data] = rows[ridx]
    pos_ones = rhs_off + cnt
    urow[pos_ones] = n_data + np.arange(nsys)
    rrow[pos_ones] = -1
    rkind[pos_ones] = 1
    return dict(
        sys_grp=grps.astype(np.int32), sys_r=r, sys_n=grp_n[grps].astype(np.int32),
        sys_miss_off=np.concatenate([[0], np.cumsum(r)])[:-1].astype(np.int64),
        sys_stn_off=np.concatenate([[0], np.cumsum(grp_n[grps])])[:-1].astype(np.int64),
        miss_list=np.concatenate([np.where(~grp_mask[g])[0] for g in grps] + [[]]).astype(np.int32),
        stn_list=np.concatenate([np.where(grp_mask[g])[0] for g in grps] + [[]]).astype(np.int32),
        sys_rhs_off=rhs_off.astype(np.int64), sys_rhs_cnt=(cnt + 1).astype(np.int32),
        rhs_urow=urow, rhs_row=rrow, rhs_kind=rkind,
        sys_order=np.argsort(-r, kind='stable').astype(np.int32),
        bt_data_step=steps[ridx].astype(np.int32), pos_ones=pos_ones.astype(np.int64))


@pytest.mark.parametrize('seed,T,N,miss', [(0, 60, 9, 0.3), (1, 400, 70, 0.05), (2, 30, 16, 0.0)])
def test_native_downdate_plan_matches_numpy(seed, T, N, miss):
    import ctypes as C
    from spinterps_b200 import _lib
    rng = np.random.default_rng(seed)
    d = rng.gamma(1.0, 5.0, size=(T, N))
    d[rng.random((T, N)) < miss] = np.nan
    d[T // 3] = d[1]
    d[T - 1] = d[1]
    gos, mask, grp_n, _, n_avail, _ = _lib.avail_groups(d)
    steps = np.where((n_avail >= 2) & (rng.random(T) < 0.8))[0].astype(np.int32)
    rows = rng.permutation(steps.size).astype(np.int64)
    exp = _numpy_downdate_plan(gos, mask, steps, rows)
    lib = _lib.load()
    nbytes = lib.spx_downdate_plan_host_bytes(steps.size)
    buf = np.zeros(nbytes, dtype=np.uint8)
    plan = _lib.spx_dd_plan()
    grp_n32 = grp_n.astype(np.int32)
    _lib.check(lib.spx_downdate_plan_host(
        gos.ctypes.data, grp_n32.ctypes.data, mask.shape[0], N,
        steps.ctypes.data, rows.ctypes.data, steps.size, buf.ctypes.data, nbytes, C.byref(plan)))
    assert plan.n_upload_bytes <= nbytes
    assert plan.n_upload_bytes <= plan.n_bytes <= lib.spx_downdate_plan_bytes(steps.size, N)
    # the device-filled station lists lie behind the uploaded prefix
    assert min(plan.off_miss_list, plan.off_stn_list) >= plan.n_upload_bytes
    assert plan.off_miss_list + 4 * plan.total_r <= plan.n_bytes
    assert plan.off_stn_list + 4 * plan.total_n <= plan.n_bytes
    ns, nd, nr = plan.n_sys, plan.n_data, plan.n_rhs
    assert (ns, nd, nr) == (exp['sys_grp'].size, steps.size, steps.size + exp['sys_grp'].size)
    assert plan.max_r == (int(exp['sys_r'].max()) if ns else 0)
    assert plan.total_r == exp['miss_list'].size and plan.total_n == exp['stn_list'].size

    def arr(name, dtype, n):
        off = getattr(plan, 'off_' + name)
        assert off % 16 == 0
        return buf[off:off + n * np.dtype(dtype).itemsize].view(dtype)
    for name, dtype, n in [('sys_grp', np.int32, ns), ('sys_r', np.int32, ns), ('sys_n', np.int32, ns),
                           ('sys_miss_off', np.int64, ns), ('sys_stn_off', np.int64, ns),
                           ('sys_rhs_off', np.int64, ns), ('sys_rhs_cnt', np.int32, ns),
                           ('rhs_urow', np.int32, nr), ('rhs_row', np.int64, nr),
                           ('rhs_kind', np.int32, nr), ('sys_order', np.int32, ns),
                           ('pos_ones', np.int64, ns)]:
        assert np.array_equal(arr(name, dtype, n), exp[name]), name
    bt = arr('bt_step', np.int32, nr)
    assert np.array_equal(bt[:nd], exp['bt_data_step'])
    assert np.array_equal(gos[bt[nd:]], exp['sys_grp'])      # any step of the system's group
    # too small a buffer is refused
    assert lib.spx_downdate_plan_host(
        gos.ctypes.data, grp_n32.ctypes.data, mask.shape[0], N,
        steps.ctypes.data, rows.ctypes.data, steps.size, buf.ctypes.data, 64, C.byref(plan)) != 0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the compiled reference of oracle/_ref, else
    the oracle port, on the host cores) prints ONE JSON line with the keys the driver
    reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1',
                        '--warmup', '0'], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'cell-steps/s' and d['value'] > 0
    assert d['metric'].startswith('interpolated cell-steps/s')
    assert d['higher_is_better'] is True and d['steps'] == 1 and d['warmup'] == 0
    from oracle import ref_runner
    kind = 'reference' if ref_runner.available() else 'port'   # oracle/_ref built or not
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and 'sample' in d['cpu_baseline']
    assert d['e2e'] == {'value': d['value'], 'unit': 'cell-steps/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_unpack_field_host_matches_numpy_division():
    """Host decode of the 2-byte field transport (spx_unpack_field_host): float32(qmin + code)
    / float32(10^d) with IEEE division, 0xFFFF -> NaN, raw rows untouched; threaded and
    single-threaded, unaligned row pitch."""
    import ctypes as C
    from spinterps_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    T, G, d = 23, 1003, 2
    stride = int(lib.spx_pack_stride(G))
    assert stride % 8 == 0 and stride >= G
    codes = rng.integers(0, 65536, size=(T, stride)).astype(np.uint16)
    codes[:, ::7] = 0xFFFF
    codes[:, 3::11] = 0xFFFE
    hdr = np.zeros(T, dtype=_lib.PACK_ROW_DTYPE)
    hdr['qmin'] = rng.integers(-40000, 40000, size=T)
    hdr['mode'][[4, 11]] = _lib.SPX_PACK_RAW
    p = np.float32(10.0 ** d)
    exp = (hdr['qmin'][:, None].astype(np.int64) + codes[:, :G].astype(np.int64)).astype(
        np.float32) / p
    exp[codes[:, :G] == 0xFFFF] = np.nan
    exp[codes[:, :G] == 0xFFFE] = -0.0
    for n_threads, ld in ((1, G), (4, G + 3)):
        out = np.full((T, ld), -1.0, dtype=np.float32)
        _lib.check(lib.spx_unpack_field_host(hdr.ctypes.data, codes.ctypes.data, T, G, d,
                                             out.ctypes.data, ld, n_threads), 'unpack')
        got = out[:, :G]
        keep = hdr['mode'] == _lib.SPX_PACK_U16
        assert np.array_equal(got[keep], exp[keep], equal_nan=True)
        assert np.array_equal(np.signbit(got[keep]), np.signbit(exp[keep]))
        assert (got[~keep] == -1.0).all() and (out[:, G:] == -1.0).all()


@pytest.mark.parametrize('seed,n,ny,nx,R,expect', [(2, 500, 1000, 1000, 20000.0, True),
                                                   (71, 220, 256, 280, 11000.0, True),
                                                   (73, 150, 160, 176, 20000.0, False),
                                                   (5, 3000, 300, 300, 2500.0, False)])
def test_station_clusters_match_scipy_components(seed, n, ny, nx, R, expect):
    """engine.station_clusters (the partition behind the sparse-covariance solve): same
    connected components as scipy's, None when one has more than 8 stations."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    from spinterps_b200.engine import station_clusters
    from tests.synth import make_problem
    p = make_problem(seed, n, 2, ny, nx, cell=1000.0)
    sx, sy = p['stn_xs'], p['stn_ys']
    pairs = cKDTree(np.column_stack([sx, sy])).query_pairs(R, output_type='ndarray')
    n_comp, ref = connected_components(
        coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n)), directed=False)
    lab = station_clusters(sx, sy, R, 8)
    if not expect:
        assert lab is None and np.bincount(ref).max() > 8
        return
    assert lab is not None and lab.max() + 1 == n_comp
    assert len(set(zip(lab.tolist(), ref.tolist()))) == n_comp        # same partition
    assert np.bincount(lab).max() <= 8
