"""Grid preparation: point-in-polygon cell / station selection and
drift raster sampling: hand-made known answers for the oracle (NumPy) and seeded
comparisons of the CUDA kernels with it (integer / index work: exact).  The fixtures
produced by the reference's own preparation code are in ``test_prep_golden.py``."""
import numpy as np
import pytest

from oracle import spinterp_oracle as orc


def _star(cx, cy, r_out, r_in, n=7, rot=0.3):
    ang = rot + np.pi * np.arange(2 * n) / n
    rad = np.where(np.arange(2 * n) % 2 == 0, r_out, r_in)
    return np.column_stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)])


def test_oracle_points_in_polygons_known_answers():
    sq = np.array([[0, 0], [10, 0], [10, 10], [0, 10]], float)
    xs = np.array([5, -1, 11, 5, 5, 10.5, 0.5, 9.999])
    ys = np.array([5, 5, 5, -0.5, 11, 10.5, 9.5, 0.001])
    assert orc.points_in_polygons(xs, ys, [sq]).tolist() == [
        True, False, False, False, False, False, True, True]
    # buffer 1: points closer than 1 to the boundary join (distance exactly 1 does not)
    assert orc.points_in_polygons(xs, ys, [sq], 1.0).tolist() == [
        True, False, False, True, False, True, True, True]
    # closed ring (repeated first vertex) and a concave polygon
    closed = np.vstack([sq, sq[:1]])
    assert np.array_equal(orc.points_in_polygons(xs, ys, [closed]),
                          orc.points_in_polygons(xs, ys, [sq]))
    ell = np.array([[0, 0], [10, 0], [10, 4], [4, 4], [4, 10], [0, 10]], float)
    assert orc.points_in_polygons([7, 2, 7], [7, 7, 2], [ell]).tolist() == [False, True, True]
    # union of two rings, area by counting on a fine grid (star: known area)
    st = _star(50.0, 40.0, 30.0, 12.0)
    g = np.arange(0.25, 100, 0.5)
    mx, my = np.meshgrid(g, g)
    m = orc.points_in_polygons(mx.ravel(), my.ravel(), [st])
    area = 0.5 * abs(np.dot(st[:, 0], np.roll(st[:, 1], -1)) - np.dot(st[:, 1], np.roll(st[:, 0], -1)))
    assert abs(m.sum() * 0.25 - area) / area < 0.01


def test_oracle_sample_raster_known_answers():
    ras = np.arange(12.0).reshape(3, 4)
    out = orc.sample_raster(ras, [0, 2, 3, 1, 0], [0, 3, 0, 1, -1], ndv=5.0)
    assert np.array_equal(out, [0.0, 11.0, np.nan, np.nan, np.nan], equal_nan=True)
    big = np.array([[-3.4028234663852886e38, 1.0]])
    out = orc.sample_raster(big, [0, 0], [0, 1], ndv=-3.4028234663852886e38)
    assert np.isnan(out[0]) and out[1] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize('buf', [0.0, 3500.0])
def test_points_in_polygons_kernel_matches_oracle(buf):
    from spinterps_b200 import prep
    rng = np.random.default_rng(3)
    rings = [_star(4.0e4, 5.0e4, 3.0e4, 1.2e4, n=9),
             _star(1.2e5, 9.0e4, 2.5e4, 2.0e4, n=150, rot=0.1),      # > 1 edge chunk
             np.array([[1.5e5, 1.0e4], [1.9e5, 1.0e4], [1.9e5, 3.0e4], [1.5e5, 3.0e4],
                       [1.5e5, 1.0e4]])]                               # closed ring
    g = 1000.0 * (np.arange(200) + 0.5)
    mx, my = np.meshgrid(g, g[::-1][:130])
    xs = np.concatenate([mx.ravel(), rng.uniform(0, 2e5, 777)])
    ys = np.concatenate([my.ravel(), rng.uniform(0, 1.3e5, 777)])
    exp = orc.points_in_polygons(xs, ys, rings, buf)
    got = prep.points_in_polygons(xs, ys, rings, buf)
    assert got.dtype == bool and np.array_equal(got, exp)
    assert 0.05 < got.mean() < 0.6


@pytest.mark.gpu
def test_sample_raster_kernel_matches_oracle():
    from spinterps_b200 import prep
    rng = np.random.default_rng(4)
    ras = rng.normal(500.0, 200.0, size=(300, 400))
    ras[rng.random(ras.shape) < 0.05] = -9999.0
    ras[10, 10] = -9999.00001            # np.isclose to the no-data value
    rows = rng.integers(-3, 303, 5000)
    cols = rng.integers(-3, 403, 5000)
    rows[:2], cols[:2] = 10, 10
    for ndv in (-9999.0, None):
        exp = orc.sample_raster(ras, rows, cols, ndv)
        got = prep.sample_raster(ras, rows, cols, ndv)
        assert np.array_equal(got, exp, equal_nan=True)
    # cell / station index helpers (interp/drift.py:175-188, :209-210)
    rr, cc = prep.drift_cell_indices(2, 4, 5, 6)
    assert rr.tolist() == [2, 2, 3, 3, 4, 4] and cc.tolist() == [5, 6, 5, 6, 5, 6]
    rr, cc = prep.drift_point_indices([1050.0, 1999.9], [8999.0, 8000.1], 1000.0, 9000.0, 100.0)
    assert rr.tolist() == [0, 9] and cc.tolist() == [0, 9]


@pytest.mark.gpu
def test_main_with_polygons_and_array_drift_raster(tmp_path):
    """SpInterpMain with polygon cell selection, station selection by buffer distance and
    an array drift raster (EDK): same fields as the oracle on the prepared arrays, and
    the prepared mask / drift equal the oracle's preparation."""
    import pandas as pd
    from spinterps_b200 import SpInterpMain, ncwriter
    rng = np.random.default_rng(8)
    n_stn, T = 45, 6
    xs = rng.uniform(0, 9e4, n_stn)
    ys = rng.uniform(0, 7e4, n_stn)
    labels = [f'S{i:05d}' for i in range(n_stn)]
    tidx = pd.date_range('2001-03-01', periods=T)
    vals = rng.gamma(1.0, 5.0, size=(T, n_stn))
    vals[rng.random(vals.shape) < 0.1] = np.nan
    data = pd.DataFrame(vals, index=tidx, columns=labels)
    crds = pd.DataFrame({'X': xs, 'Y': ys}, index=labels)
    vg = '0.1 Nug(0.0) + 0.9 Sph(30000)'
    rings = [_star(3.0e4, 3.0e4, 2.0e4, 1.0e4, n=6), _star(6.5e4, 4.5e4, 1.5e4, 1.2e4, n=8)]
    cs = 2000.0
    # drift raster covering more than the grid, with a no-data patch
    rx0, ry1 = -2.0e4, 1.0e5
    rr, cc = np.meshgrid(np.arange(70), np.arange(80), indexing='ij')
    elev = 300.0 + 0.004 * (rx0 + (cc + 0.5) * cs) + 0.002 * (ry1 - (rr + 0.5) * cs)
    elev[27:29, 42:45] = -9999.0            # inside the second polygon
    m = SpInterpMain(verbose=False)
    m.set_data(data, crds)
    m.set_vgs_ser(pd.Series([vg] * T, index=tidx))
    m.set_out_dir(tmp_path / 'run')
    m.set_netcdf4_parameters('precip.nc', 'mm', 'precipitation', 'days since 1900-01-01',
                             'gregorian', 2, 1)
    m.set_interp_time_parameters('2001-03-01', '2001-03-06', 'D', '%Y-%m-%d')
    m.set_neighbor_selection_method('all')
    m.set_misc_settings(cell_size=cs, min_cutoff_value=0.0)
    m.set_cell_selection_polygons(rings, 1.5e4, True, 3000.0)
    m.turn_external_drift_kriging_on([dict(values=elev, x_min=rx0, y_max=ry1, cell_size=cs,
                                           ndv=-9999.0)])
    m.turn_ordinary_kriging_on()
    m.verify()
    # preparation against the oracle
    keep = orc.points_in_polygons(xs, ys, rings, 1.5e4)
    assert 2 < keep.sum() < n_stn
    assert list(m._crds_df.index) == [s for s, k in zip(labels, keep) if k]
    allv = np.concatenate(rings)
    assert np.isclose(m._x_min, allv[:, 0].min() - 3000.0) and np.isclose(m._y_max, allv[:, 1].max() + 3000.0)
    ny, nx = m._interp_crds_orig_shape
    # the grid's row / column window is RASTER-relative (interp/prepare.py:163-173): with an
    # origin that is not aligned to the raster it spans max - min + 1 columns / rows
    assert ((m._x_min - rx0) / cs) % 1.0 > 1e-6 and ((ry1 - m._y_max) / cs) % 1.0 > 1e-6
    assert nx == (int(np.ceil((m._x_max - rx0) / cs)) - 1) - int(np.floor((m._x_min - rx0) / cs)) + 1
    assert ny == (int(np.ceil((ry1 - m._y_min) / cs)) - 1) - int(np.floor((ry1 - m._y_max) / cs)) + 1
    assert nx >= int(np.ceil((m._x_max - m._x_min) / cs)) and ny >= int(np.ceil((m._y_max - m._y_min) / cs))
    gx = m._x_min + cs * (np.arange(nx) + 0.5)
    gy = m._y_max - cs * (np.arange(ny) + 0.5)
    fx, fy = np.meshgrid(gx, gy)
    mask = orc.points_in_polygons(fx.ravel(), fy.ravel(), rings, 3000.0)
    assert np.array_equal(m._cntn_idxs, mask) and 0.1 < mask.mean() < 0.9
    # the reference addresses the raster through the grid's row / column window
    # (interp/prepare.py:150-172, interp/drift.py:175-188), not through cell centres
    kc, kr = np.meshgrid(np.arange(nx), np.arange(ny))
    col = (int(np.floor((m._x_min - rx0) / cs)) + kc.ravel())[mask]
    row = (int(np.floor((ry1 - m._y_max) / cs)) + kr.ravel())[mask]
    assert np.array_equal(m._drft_arrs[0], orc.sample_raster(elev, row, col, -9999.0), equal_nan=True)
    assert np.isnan(m._drft_arrs).sum() > 0          # the no-data patch lies inside a polygon
    m.interpolate()
    exp, _ = orc.interp_chunk(
        m._data_df.values, m._crds_df['X'].values, m._crds_df['Y'].values,
        m._interp_x_crds_msh, m._interp_y_crds_msh, m._interp_crds_orig_shape, m._interp_args,
        vgs=[vg] * T, cntn_idxs=m._cntn_idxs, drft_arrs=m._drft_arrs,
        stns_drft=m._stns_drft_df.values, min_var_cut=0.0, intrp_dtype=np.float32)
    h = ncwriter.open_for_read(m._nc_file_path)
    for lab in ('OK', 'EDK'):
        ref = np.round(exp[lab], 2).reshape(T, ny, nx)
        for t in range(T):
            got = h.read(lab, t)
            assert np.array_equal(np.isnan(got), np.isnan(ref[t]))
            assert np.nanmax(np.abs(got - ref[t])) <= 0.0101, (lab, t)
    h.close()
