"""Grid preparation against REFERENCE output: the fixtures
``tests/golden/p*_prep_*.npz`` were produced by the reference's own preparation methods
(interp/prepare.py:92-288, interp/bdpolys.py:26-182, interp/drift.py:25-226,
misc.py:221-540) running unmodified over in-memory stand-ins for OGR / GDAL
(``tests/golden/make_golden_prep.py``).  CPU tests pin the oracle's statement to them, GPU
tests compare the CUDA kernels and ``SpInterpMain``'s preparation with them bit for bit."""
from pathlib import Path

import numpy as np
import pytest

from oracle import spinterp_oracle as orc

GOLD = Path(__file__).parent / 'golden'
CASES = ['p1_prep_polys_edk', 'p2_prep_plain', 'p3_prep_polys_stations',
         'p4_prep_polys_nobuf_aligned', 'p5_prep_align_polys_edk', 'p6_prep_align_only']


def _load(name):
    d = dict(np.load(GOLD / f'{name}.npz'))
    d['rings'] = [d[f'ring{i}'] for i in range(int(d['n_rings']))] or None
    d['rasters'] = [d[f'raster{i}'] for i in range(int(d['n_rasters']))]
    return d


def _raster_geo(d):
    if not d['rasters']:
        return None
    x_min, y_max = d['raster_geo'][:2]
    return (x_min, y_max) + d['rasters'][0].shape


def _align(d):
    """(x_min, y_max, cell_size, n_rows, n_cols) of the alignment raster or None."""
    a = d['align']
    return None if np.isnan(a[0]) else (float(a[0]), float(a[1]), float(a[2]), int(a[3]), int(a[4]))


@pytest.mark.skipif(not Path('/root/reference/interp/prepare.py').exists(),
                    reason='needs the reference sources (build container only)')
def test_fixtures_regenerate_from_the_reference(tmp_path):
    """The committed fixtures ARE what the reference's preparation code produces: the
    generating script, run again in a fresh process, writes byte-identical files."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(GOLD / 'make_golden_prep.py'), '--out', str(tmp_path)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in CASES:
        assert (tmp_path / f'{name}.npz').read_bytes() == (GOLD / f'{name}.npz').read_bytes(), name


@pytest.mark.parametrize('name', CASES)
def test_oracle_preparation_matches_the_reference(name):
    d = _load(name)
    cs = float(d['cell_size'])
    if _align(d) is not None:
        assert cs == _align(d)[2]               # prepare.py:553-555: the alignment raster's
    elif d['rasters']:
        assert cs == d['raster_geo'][2] and np.isnan(d['cell_size_in'])   # prepare.py:549-550
    bounds, window, xs, ys = orc.prepare_grid(
        d['stn_xs'], d['stn_ys'], cs, float(d['cell_bdist']), d['rings'], _raster_geo(d),
        _align(d))
    if _align(d) is not None:
        # on the alignment lattice, and (with polygons) not aligned before the adjustment
        ax0, ay1 = _align(d)[:2]
        for v, o in ((bounds[0], ax0), (bounds[1], ax0), (bounds[2], ay1), (bounds[3], ay1)):
            r = ((v - o) / cs) % 1.0
            assert min(r, 1.0 - r) < 1e-6
    assert np.array_equal(bounds, d['bounds'])
    assert np.array_equal(window, d['window'])
    assert np.array_equal(xs, d['nc_x_crds']) and np.array_equal(ys, d['nc_y_crds'])
    assert (ys.size, xs.size) == tuple(d['grid_shape'])
    mx, my = np.meshgrid(xs, ys)
    mx, my = mx.ravel(), my.ravel()
    mask = None
    if d['rings'] is not None:
        keep = orc.points_in_polygons(d['stn_xs'], d['stn_ys'], d['rings'], float(d['stn_bdist']))
        assert np.array_equal(np.flatnonzero(keep), d['sel_stations'])
        assert 0 < keep.sum() < keep.size
        if bool(d['ipoly']):
            mask = orc.points_in_polygons(mx, my, d['rings'], float(d['cell_bdist']))
            assert np.array_equal(mask, d['cntn_idxs']) and 0.1 < mask.mean() < 0.9
            mx, my = mx[mask], my[mask]
        else:
            assert 'cntn_idxs' not in d
    else:
        assert np.array_equal(d['sel_stations'], np.arange(d['stn_xs'].size))
    assert np.array_equal(mx, d['cell_xs']) and np.array_equal(my, d['cell_ys'])
    if d['rasters']:
        rx_min, ry_max, _, ndv = d['raster_geo']
        ndv = None if np.isnan(ndv) else ndv
        rows, cols = orc.drift_window_indices(window, mask)
        sel = d['sel_stations']
        srows, scols = orc.drift_station_indices(
            d['stn_xs'][sel], d['stn_ys'][sel], d['drft_bounds'][0], d['drft_bounds'][3], cs)
        for i, ras in enumerate(d['rasters']):
            assert np.array_equal(orc.sample_raster(ras, rows, cols, ndv), d['drft_arrs'][i],
                                  equal_nan=True)
            assert np.array_equal(orc.sample_raster(ras, srows, scols, ndv), d['stns_drft'][:, i])
        if name == 'p1_prep_polys_edk':
            # the no-data patches lie inside the polygons; the second one differs from the
            # no-data value by 1e-9 (np.isclose, interp/drift.py:193)
            assert np.isnan(d['drft_arrs'][0]).sum() > 0 and np.isnan(d['drft_arrs'][1]).sum() > 0
            # grid origin not aligned to the raster: one more column / row than the extent
            assert ((d['bounds'][0] - d['drft_bounds'][0]) / cs) % 1.0 > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_kernels_match_the_reference_preparation(name):
    from spinterps_b200 import prep
    d = _load(name)
    mx, my = np.meshgrid(d['nc_x_crds'], d['nc_y_crds'])
    mx, my = mx.ravel(), my.ravel()
    mask = None
    if d['rings'] is not None:
        keep = prep.points_in_polygons(d['stn_xs'], d['stn_ys'], d['rings'], float(d['stn_bdist']))
        assert np.array_equal(np.flatnonzero(keep), d['sel_stations'])
        if bool(d['ipoly']):
            mask = prep.points_in_polygons(mx, my, d['rings'], float(d['cell_bdist']))
            assert np.array_equal(mask, d['cntn_idxs'])
    if d['rasters']:
        ndv = None if np.isnan(d['raster_geo'][3]) else d['raster_geo'][3]
        mr0, mr1, mc0, mc1 = (int(v) for v in d['window'])
        rows, cols = prep.drift_cell_indices(mr0, mr1, mc0, mc1, mask)
        sel = d['sel_stations']
        srows, scols = prep.drift_point_indices(
            d['stn_xs'][sel], d['stn_ys'][sel], d['drft_bounds'][0], d['drft_bounds'][3],
            float(d['cell_size']))
        for i, ras in enumerate(d['rasters']):
            assert np.array_equal(prep.sample_raster(ras, rows, cols, ndv), d['drft_arrs'][i],
                                  equal_nan=True)
            assert np.array_equal(prep.sample_raster(ras, srows, scols, ndv), d['stns_drft'][:, i])


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_main_preparation_matches_the_reference(name, tmp_path):
    """``SpInterpMain.verify()`` (our ``_prepare``) leaves the attributes the reference's
    ``_prepare`` leaves."""
    import pandas as pd
    from spinterps_b200 import SpInterpMain
    d = _load(name)
    n_stn, T = d['stn_xs'].size, 3
    labels = [f'S{i:05d}' for i in range(n_stn)]
    tidx = pd.date_range('2001-03-01', periods=T)
    rng = np.random.default_rng(5)
    data = pd.DataFrame(rng.gamma(1.0, 5.0, size=(T, n_stn)), index=tidx, columns=labels)
    crds = pd.DataFrame({'X': d['stn_xs'], 'Y': d['stn_ys']}, index=labels)
    m = SpInterpMain(verbose=False)
    m.set_data(data, crds)
    m.set_vgs_ser(pd.Series(['0.1 Nug(0.0) + 0.9 Sph(30000)'] * T, index=tidx))
    m.set_out_dir(tmp_path / 'run')
    m.set_netcdf4_parameters('precip.nc', 'mm', 'precipitation', 'days since 1900-01-01',
                             'gregorian', 2, 1)
    m.set_interp_time_parameters('2001-03-01', '2001-03-03', 'D', '%Y-%m-%d')
    m.set_neighbor_selection_method('all')
    if np.isnan(d['cell_size_in']):
        m.set_misc_settings(min_cutoff_value=0.0)         # cell size comes from the raster
    else:
        m.set_misc_settings(cell_size=float(d['cell_size_in']), min_cutoff_value=0.0)
    if d['rings'] is not None:
        m.set_cell_selection_polygons(d['rings'], float(d['stn_bdist']), bool(d['ipoly']),
                                      float(d['cell_bdist']))
    if _align(d) is not None:
        ax0, ay1, acs, anr, anc = _align(d)
        m.set_alignment_raster(dict(x_min=ax0, y_max=ay1, cell_size=acs, n_rows=anr, n_cols=anc))
    if d['rasters']:
        rx_min, ry_max, rcs, ndv = d['raster_geo']
        m.turn_external_drift_kriging_on([
            dict(values=r, x_min=rx_min, y_max=ry_max, cell_size=rcs,
                 ndv=None if np.isnan(ndv) else ndv) for r in d['rasters']])
    m.turn_ordinary_kriging_on()
    m.verify()
    assert m._cell_size == float(d['cell_size'])
    assert np.array_equal([m._x_min, m._x_max, m._y_min, m._y_max], d['bounds'])
    assert np.array_equal([m._min_row, m._max_row, m._min_col, m._max_col], d['window'])
    assert tuple(m._interp_crds_orig_shape) == tuple(d['grid_shape'])
    assert np.array_equal(m._nc_x_crds, d['nc_x_crds'])
    assert np.array_equal(m._nc_y_crds, d['nc_y_crds'])
    assert np.array_equal(m._interp_x_crds_msh, d['cell_xs'])
    assert np.array_equal(m._interp_y_crds_msh, d['cell_ys'])
    # the reference keeps the selected stations in a hash order; compare the sets
    assert sorted(m._crds_df.index) == [labels[i] for i in d['sel_stations']]
    assert sorted(m._data_df.columns) == sorted(m._crds_df.index)
    if 'cntn_idxs' in d:
        assert m._cntn_idxs.dtype == bool and np.array_equal(m._cntn_idxs, d['cntn_idxs'])
    else:
        assert m._cntn_idxs is None
    if d['rasters']:
        assert np.array_equal([m._drft_x_min, m._drft_x_max, m._drft_y_min, m._drft_y_max],
                              d['drft_bounds'])
        assert np.array_equal(m._drft_arrs, d['drft_arrs'], equal_nan=True)
        got = m._stns_drft_df.loc[sorted(m._crds_df.index)].values
        assert np.array_equal(got, d['stns_drft'])
