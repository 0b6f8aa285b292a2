"""Golden fixtures for the grid preparation from the
UNMODIFIED reference.  Build container only (needs /root/reference):

    python tests/golden/make_golden_prep.py [--out DIR]

The reference prepares the grid with OGR / GDAL, which are absent here.  This script runs
the reference's own methods --

    SpInterpPrepare._cmpt_corner_coordinates      interp/prepare.py:92-146
    SpInterpPrepare._cmpt_aligned_coordinates     interp/prepare.py:45-90
        -> misc.get_aligned_shp_bds_and_cell_size (misc.py:743-885)
    SpInterpBoundaryPolygons._select_nearest_stations   interp/bdpolys.py:26-182
        -> misc.get_all_polys_in_shp / linearize_sub_polys / chk_pt_cntmnt_in_polys_mp
           (misc.py:221-286, :372-404, :407-540)
    KrigingDrift._assemble_drift_data             interp/drift.py:25-163
    SpInterpPrepare._prepare_crds                 interp/prepare.py:148-242
    SpInterpPrepare._select_nearby_cells          interp/prepare.py:244-288
    KrigingDrift._prepare_stns_drift              interp/drift.py:165-226

-- in the order of ``_prepare`` (interp/prepare.py:549-594) on an instance of the
reference's ``SpInterpPrepare`` whose attributes are set directly, with ``ogr`` / ``gdal``
replaced by the in-memory stand-ins of ``gis_standin.py`` (polygons and rasters from
arrays).  Everything the reference computes itself is therefore reference output: grid
bounds, raster-relative row / column window, cell-centre coordinates, the envelope
pre-filter / chunking / union of the containment search, the boolean cell mask, the
selected stations, drift values of cells and stations with no-data -> NaN.  The two
geometric predicates (point in ring, distance buffer) are the stand-in's; the script
refuses inputs on which GEOS could answer differently (a point closer to a ring or to a
buffer outline than the sagitta of a 30-segments-per-quadrant arc, or within the 1e-6 of
the reference's ``"POINT (%f %f)"`` round trip).
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import gis_standin                      # noqa: E402
from make_golden import import_reference  # noqa: E402


def star(cx, cy, r_out, r_in, n=7, rot=0.3):
    ang = rot + np.pi * np.arange(2 * n) / n
    rad = np.where(np.arange(2 * n) % 2 == 0, r_out, r_in)
    return np.column_stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)])


def check_margins(gis, xs, ys, rings, buf, what):
    """No point where GEOS and the stand-in could disagree."""
    polys = [gis_standin.polygon_from_ring(r) for r in rings]
    sag = buf * (1.0 - np.cos(np.pi / 120.0)) + 1e-3       # 30 segments per quadrant
    worst = np.inf
    for x, y in zip(xs, ys):
        for p in polys:
            d = p.edge_distance(x, y)
            worst = min(worst, d, abs(d - buf) - sag if buf > 0 else np.inf)
    assert worst > 1e-3, (what, worst)
    return worst


OUT = HERE          # --out DIR writes the fixtures elsewhere (the regeneration test)


def run_case(mods, gis, name, *, stn_xs, stn_ys, cell_size, rings=None, stn_bdist=0.0,
             cell_bdist=0.0, ipoly=True, rasters=None, align=None):
    prep_mod, misc = mods
    n_stn = stn_xs.size
    labels = [f'S{i:05d}' for i in range(n_stn)]
    crds_df = pd.DataFrame({'X': stn_xs, 'Y': stn_ys}, index=labels)
    data_df = pd.DataFrame(np.zeros((2, n_stn)), columns=labels,
                           index=pd.date_range('2000-01-01', periods=2))

    p = prep_mod.SpInterpPrepare()
    p._vb = False
    p._n_cpus = 1
    p._crds_df = crds_df.copy()
    p._data_df = data_df.copy()
    p._cell_size = cell_size
    p._cell_bdist = float(cell_bdist)
    p._stn_bdist = float(stn_bdist)
    p._ipoly_flag = bool(ipoly)
    p._cell_sel_prms_set = rings is not None
    p._algn_ras_set_flag = align is not None
    p._poly_shp = None
    if align is not None:       # (x_min, y_max, cell_size, n_rows, n_cols): geometry only
        p._algn_ras = Path(f'/standin/{name}_align.tif')
        gis.add_raster(p._algn_ras, np.zeros((align[3], align[4])), align[0], align[1],
                       align[2], None)
    p._poly_simplify_tol_ratio = 0.0
    p._plot_figs_flag = False
    p._edk_flag = rasters is not None
    if rings is not None:
        p._poly_shp = Path(f'/standin/{name}.shp')
        gis.add_polygons(p._poly_shp, rings)
    if rasters is not None:
        p._drft_rass = []
        for i, r in enumerate(rasters):
            path = Path(f'/standin/{name}_drift{i}.tif')
            gis.add_raster(path, r['values'], r['x_min'], r['y_max'], r['cell_size'], r['ndv'])
            p._drft_rass.append(path)

    # interp/prepare.py:549-594, the GIS-dependent half of _prepare in its order
    if p._edk_flag and (not p._algn_ras_set_flag):
        p._cell_size = misc.get_ras_props(str(p._drft_rass[0]))[6]
    if p._algn_ras_set_flag:
        p._cmpt_aligned_coordinates()       # updates the cell size to the raster's
    else:
        p._cmpt_corner_coordinates()
    assert p._cell_size is not None
    if p._cell_sel_prms_set:
        p._select_nearest_stations()
    if p._edk_flag:
        p._assemble_drift_data()
    p._prepare_crds()
    full_x, full_y = p._interp_x_crds_msh.copy(), p._interp_y_crds_msh.copy()
    if p._cell_sel_prms_set and p._ipoly_flag:
        p._select_nearby_cells()
    if p._edk_flag:
        p._prepare_stns_drift()

    if rings is not None:
        check_margins(gis, stn_xs, stn_ys, rings, stn_bdist, name + ' stations')
        if ipoly:
            check_margins(gis, full_x, full_y, rings, cell_bdist, name + ' cells')

    # the reference keeps the selected stations in set order (misc.py:533-539: a hash
    # order, different from run to run); store them sorted
    sel = sorted(p._crds_df.index)
    out = dict(
        stn_xs=stn_xs, stn_ys=stn_ys, cell_size_in=np.float64(np.nan if cell_size is None else cell_size),
        stn_bdist=np.float64(stn_bdist), cell_bdist=np.float64(cell_bdist), ipoly=np.bool_(ipoly),
        n_rings=np.int64(0 if rings is None else len(rings)),
        n_rasters=np.int64(0 if rasters is None else len(rasters)),
        align=np.array([np.nan] * 5 if align is None else align, dtype=np.float64),
        cell_size=np.float64(p._cell_size),
        bounds=np.array([p._x_min, p._x_max, p._y_min, p._y_max]),
        window=np.array([p._min_row, p._max_row, p._min_col, p._max_col], dtype=np.int64),
        grid_shape=np.array(p._interp_crds_orig_shape, dtype=np.int64),
        nc_x_crds=p._nc_x_crds, nc_y_crds=p._nc_y_crds,
        cell_xs=p._interp_x_crds_msh, cell_ys=p._interp_y_crds_msh,
        sel_stations=np.array([labels.index(s) for s in sel], dtype=np.int64))
    if rings is not None:
        for i, r in enumerate(rings):
            out[f'ring{i}'] = np.asarray(r, dtype=np.float64)
    if p._cntn_idxs is not None:
        out['cntn_idxs'] = p._cntn_idxs
    if rasters is not None:
        for i, r in enumerate(rasters):
            out[f'raster{i}'] = np.asarray(r['values'], dtype=np.float64)
        r0 = rasters[0]
        out['raster_geo'] = np.array([r0['x_min'], r0['y_max'], r0['cell_size'],
                                      np.nan if r0['ndv'] is None else r0['ndv']])
        out['drft_bounds'] = np.array([p._drft_x_min, p._drft_x_max, p._drft_y_min, p._drft_y_max])
        out['drft_arrs'] = p._drft_arrs
        out['stns_drft'] = p._stns_drft_df.loc[sel].values.astype(np.float64)
    np.savez_compressed(OUT / f'{name}.npz', **out)
    print(name, 'grid', tuple(out['grid_shape']), 'cells', out['cell_xs'].size,
          'stations', len(sel), 'of', n_stn,
          'NaN drift cells', int(np.isnan(out['drft_arrs']).sum()) if rasters else '-')


def main():
    import_reference()
    gis = gis_standin.Registry()
    misc = importlib.import_module('spinterps.misc')
    prep_mod = importlib.import_module('spinterps.interp.prepare')
    bd_mod = importlib.import_module('spinterps.interp.bdpolys')
    dr_mod = importlib.import_module('spinterps.interp.drift')
    misc.ogr = misc.gdal = gis
    prep_mod.ogr = gis
    bd_mod.ogr = gis
    dr_mod.gdal = gis
    mods = (prep_mod, misc)

    # p1: polygons (stations by buffer, cells by buffer), two drift rasters with a no-data
    # patch, grid origin NOT aligned to the raster, cell size taken from the raster
    rng = np.random.default_rng(21)
    n_stn = 60
    xs = rng.uniform(0, 9e4, n_stn)
    ys = rng.uniform(0, 7e4, n_stn)
    rings = [star(3.0e4, 3.0e4, 2.0e4, 1.0e4, n=6), star(6.5e4, 4.5e4, 1.5e4, 1.2e4, n=8),
             np.array([[3.1e4, 5.02e4], [4.63e4, 5.02e4], [4.63e4, 5.817e4], [3.1e4, 5.817e4]])]
    cs = 2000.0
    rx0, ry1 = -2.03e4, 1.007e5
    rr, cc = np.meshgrid(np.arange(70), np.arange(80), indexing='ij')
    elev = 300.0 + 0.004 * (rx0 + (cc + 0.5) * cs) + 0.002 * (ry1 - (rr + 0.5) * cs)
    elev = elev + rng.normal(0, 3.0, elev.shape)
    slope = rng.uniform(0.0, 30.0, elev.shape)
    elev[27:29, 42:45] = -9999.0
    slope[40:42, 20:22] = -9999.0 + 1e-9        # np.isclose to the no-data value
    run_case(mods, gis, 'p1_prep_polys_edk', stn_xs=xs, stn_ys=ys, cell_size=None, rings=rings,
             stn_bdist=1.5e4, cell_bdist=3000.0, ipoly=True,
             rasters=[dict(values=elev, x_min=rx0, y_max=ry1, cell_size=cs, ndv=-9999.0),
                      dict(values=slope, x_min=rx0, y_max=ry1, cell_size=cs, ndv=-9999.0)])

    # p2: no polygons, no drift: bounds from the station extent
    rng = np.random.default_rng(22)
    xs = rng.uniform(3.1234e5, 3.9e5, 25)
    ys = rng.uniform(5.2e6, 5.27e6, 25)
    run_case(mods, gis, 'p2_prep_plain', stn_xs=xs, stn_ys=ys, cell_size=1500.0)

    # p3: polygons select stations only (interp_around_polys_flag False): full grid over
    # the polygons' extent + cell buffer, no mask
    rng = np.random.default_rng(23)
    xs = rng.uniform(0, 1.0e5, 80)
    ys = rng.uniform(0, 8e4, 80)
    rings = [star(5.0e4, 4.0e4, 2.5e4, 1.4e4, n=5, rot=0.7)]
    run_case(mods, gis, 'p3_prep_polys_stations', stn_xs=xs, stn_ys=ys, cell_size=2500.0,
             rings=rings, stn_bdist=8000.0, cell_bdist=5000.0, ipoly=False)

    # p4: polygons without any buffer (pure containment), drift raster aligned to the grid
    rng = np.random.default_rng(24)
    xs = rng.uniform(1.0e4, 7.0e4, 70)
    ys = rng.uniform(1.0e4, 6.0e4, 70)
    rings = [star(4.0e4, 3.5e4, 2.8e4, 1.7e4, n=9, rot=0.2)]
    cs = 1000.0
    allv = np.concatenate(rings)
    rx0 = allv[:, 0].min() - 7 * cs
    ry1 = allv[:, 1].max() + 5 * cs
    ras = rng.normal(800.0, 100.0, (75, 80))
    run_case(mods, gis, 'p4_prep_polys_nobuf_aligned', stn_xs=xs, stn_ys=ys, cell_size=None,
             rings=rings, stn_bdist=0.0, cell_bdist=0.0, ipoly=True,
             rasters=[dict(values=ras, x_min=rx0, y_max=ry1, cell_size=cs, ndv=None)])


    # p5: alignment raster + polygons (stations and cells by buffer) + a drift raster on
    # the same lattice: bounds snapped outwards to the alignment lattice, cell size from it
    rng = np.random.default_rng(25)
    xs = rng.uniform(0, 9e4, 60)
    ys = rng.uniform(0, 7e4, 60)
    rings = [star(3.3e4, 3.1e4, 2.0e4, 1.1e4, n=6, rot=0.4), star(6.4e4, 4.4e4, 1.4e4, 1.1e4, n=7)]
    cs = 1500.0
    ax0, ay1 = -3.0e4 + 137.0, 1.2e5 + 61.0          # lattice origin unrelated to the polygons
    ras = rng.normal(400.0, 50.0, (110, 120))
    ras[50:52, 60:63] = -32768.0
    run_case(mods, gis, 'p5_prep_align_polys_edk', stn_xs=xs, stn_ys=ys, cell_size=None,
             rings=rings, stn_bdist=1.2e4, cell_bdist=2450.0, ipoly=True,
             rasters=[dict(values=ras, x_min=ax0 - 3 * cs, y_max=ay1 + 2 * cs, cell_size=cs,
                           ndv=-32768.0)],
             align=(ax0, ay1, cs, 100, 110))

    # p6: alignment raster alone: the grid is the raster's extent
    rng = np.random.default_rng(26)
    xs = rng.uniform(1.0e4, 5.0e4, 30)
    ys = rng.uniform(1.0e4, 4.0e4, 30)
    run_case(mods, gis, 'p6_prep_align_only', stn_xs=xs, stn_ys=ys, cell_size=999.0,
             align=(5.0e3 + 0.25, 4.5e4 - 0.75, 1250.0, 30, 38))


if __name__ == '__main__':
    if '--out' in sys.argv:
        OUT = Path(sys.argv[sys.argv.index('--out') + 1])
        OUT.mkdir(parents=True, exist_ok=True)
    main()
