"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; ~1 min for the Cython
build):

    python tests/golden/make_golden.py [--out DIR]

It imports the reference package from /root/reference exactly as SURVEY.md
section 8c describes (pyximport with a writable build dir, MagicMock stubs for
the absent I/O libraries, pandas string inference off), builds a fake "main"
object with the 19 attributes read at interp/steps.py:33-53, calls
``SpInterpSteps(fake)._get_all_interp_outputs(args)`` and stores inputs and FP64
outputs as ``<case>.npz``.  The GPU box never runs this; tests read the .npz.
"""
import importlib
import os
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
OUT = HERE          # --out DIR writes the fixtures elsewhere (the regeneration test)
REF = Path('/root/reference')


def import_reference(build_dir='/tmp/spinterps_refbuild'):
    pd.set_option('future.infer_string', False)
    for m in ['netCDF4', 'osgeo', 'osgeo.ogr', 'osgeo.gdal', 'pathos',
              'pathos.multiprocessing', 'shapefile', 'cftime', 'matplotlib',
              'matplotlib.pyplot', 'descartes']:
        sys.modules.setdefault(m, MagicMock())
    pkg = types.ModuleType('spinterps')
    pkg.__path__ = [str(REF)]
    sys.modules['spinterps'] = pkg
    import pyximport
    os.makedirs(build_dir, exist_ok=True)
    pyximport.install(build_dir=build_dir, language_level=3)
    cy = types.ModuleType('spinterps.cyth')
    cy.__path__ = [str(REF / 'cyth')]
    sys.modules['spinterps.cyth'] = cy
    im = importlib.import_module('spinterps.cyth.interpmthds')
    for k in dir(im):
        if not k.startswith('_'):
            setattr(cy, k, getattr(im, k))
    from spinterps.interp.steps import SpInterpSteps
    import spinterps.misc as misc
    return im, SpInterpSteps, misc


class FakeLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_reference(SpInterpSteps, case):
    """case: dict of plain arrays / settings (see make_case)."""
    n_stn = case['stn_xs'].size
    labels = [f'S{i:05d}' for i in range(n_stn)]
    T = case['data'].shape[0]
    tidx = pd.date_range('2000-01-01', periods=T)
    data_df = pd.DataFrame(case['data'].copy(), index=tidx, columns=labels)
    crds_df = pd.DataFrame({'X': case['stn_xs'], 'Y': case['stn_ys']}, index=labels)

    main = types.SimpleNamespace(
        _vb=False, _n_cpus=1, _mp_flag=False, _crds_df=crds_df,
        _min_var_thr=case['min_var_thr'], _min_var_cut=case['min_var_cut'],
        _max_var_cut=case['max_var_cut'], _cntn_idxs=case['cntn_idxs'],
        _interp_crds_orig_shape=tuple(case['grid_shape']),
        _interp_x_crds_msh=case['cell_xs'].copy(),
        _interp_y_crds_msh=case['cell_ys'].copy(),
        _nc_file_path=None, _nc_nmrl_prcn=2,
        _neb_sel_mthd=case['neb_sel_mthd'], _n_nebs=case['n_nebs'],
        _n_pies=case.get('n_pies'),
        _min_vg_val=case['min_vg_val'],
        _interp_flag_est_vars=case['est_var_flag'], _intrp_dtype=np.float64)

    vgs_ser = None
    rord = None
    if case['vgs'] is not None:
        vgs_ser = pd.Series(list(case['vgs']), index=tidx, dtype=object)
        rord = pd.Series(np.arange(T), index=tidx)
    stns_drft_df = None
    if case['stns_drft'] is not None:
        stns_drft_df = pd.DataFrame(case['stns_drft'], index=labels)

    args = (data_df, 0, T, 1, case['interp_args'], FakeLock(), case['drft_arrs'],
            stns_drft_df, vgs_ser, rord, case['fld_beg_row'], case['fld_end_row'])
    out = SpInterpSteps(main)._get_all_interp_outputs(args)
    return out[7]


def make_inputs(seed, n_stn, T, ny, nx, cell=5000.0, miss=0.0):
    rng = np.random.default_rng(seed)
    side_x, side_y = nx * cell, ny * cell
    stn_xs = rng.uniform(0, side_x, n_stn)
    stn_ys = rng.uniform(0, side_y, n_stn)
    data = rng.gamma(1.0, 5.0, size=(T, n_stn))
    if miss > 0:
        data[rng.random((T, n_stn)) < miss] = np.nan
    xs = 0.5 * cell + cell * np.arange(nx)
    ys = side_y - 0.5 * cell - cell * np.arange(ny)
    mx, my = np.meshgrid(xs, ys)
    return rng, stn_xs, stn_ys, data, mx.ravel(), my.ravel()


def base_case(**kw):
    case = dict(
        vgs=None, cntn_idxs=None, drft_arrs=None, stns_drft=None,
        fld_beg_row=0, fld_end_row=None, neb_sel_mthd='all', n_nebs=None, n_pies=None,
        min_var_thr=-np.inf, min_var_cut=None, max_var_cut=None, min_vg_val=0.0,
        est_var_flag=False)
    case.update(kw)
    if case['fld_end_row'] is None:
        case['fld_end_row'] = case['grid_shape'][0]
    return case


def cases():
    VG = '0.1 Nug(0.0) + 0.9 Sph(20000)'
    out = {}

    # A: the seeded probe of SURVEY.md section 8c (OK + IDW p=2 + NNB, 'all')
    rng, sx, sy, data, cx, cy = make_inputs(0, 20, 5, 20, 20)
    out['a_ok_idw_nnb'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(20, 20),
        vgs=[VG] * 5,
        interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0), ('NNB', None, 'NNB')])

    # B: missing data -> several availability groups, 3 vg strings (one nugget
    # only), low-value steps (min_var_thr), cut-offs, est. variance
    rng, sx, sy, data, cx, cy = make_inputs(1, 30, 12, 15, 15, miss=0.2)
    data[3, :] = np.where(np.isnan(data[3, :]), np.nan, 0.01)
    data[7, :] = np.where(np.isnan(data[7, :]), np.nan, 0.05)
    data[9, 1:] = np.nan  # a single-station step (n_refs == 1)
    data[9, 0] = 3.25
    data[10, :] = np.nan  # a step without any station
    vgs = [VG] * 12
    for i in (1, 5, 8):
        vgs[i] = '0.25 Nug(0.0) + 1.10 Exp(35000.0)'
    vgs[2] = '0.0 Nug(0.0)'
    vgs[6] = '0.3 Sph(12000) + 0.7 Sph(45000)'
    out['b_ok_groups_flags'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(15, 15),
        vgs=vgs, min_var_thr=0.1, min_var_cut=0.0, max_var_cut=14.0, est_var_flag=True,
        interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 1.0),
                     ('IDW', None, 'IDW_001', 3.5), ('NNB', None, 'NNB'),
                     ('EST_VARS_OK', None, 'EST_VARS_OK')])

    # C: EDK, two drifts, per-step variograms, NaN drift at some cells, missing data
    rng, sx, sy, data, cx, cy = make_inputs(2, 25, 6, 12, 14, miss=0.15)

    def elev(x, y):
        return 300 + 0.004 * x + 0.002 * y + 80 * np.sin(x / 9000.0) * np.cos(y / 7000.0)

    def slope(x, y):
        return 5 + 2 * np.cos(x / 15000.0) + 1e-4 * y
    drft = np.vstack([elev(cx, cy), slope(cx, cy)])
    drft[0, [5, 77]] = np.nan
    sdrft = np.column_stack([elev(sx, sy), slope(sx, sy)])
    vgs = ['%0.5f Nug(0.0) + %0.5f Sph(%0.5f)' % (rng.uniform(0, 0.2), rng.uniform(0.5, 1.5),
                                                    rng.uniform(1e4, 5e4)) for _ in range(6)]
    out['c_edk_drift'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(12, 14),
        vgs=vgs, drft_arrs=drft, stns_drft=sdrft,
        interp_args=[('OK', None, 'OK'), ('EDK', None, 'EDK')])

    # D: SK + OK with a polygon-like mask and a grid-row chunk
    rng, sx, sy, data, cx, cy = make_inputs(3, 22, 4, 16, 18, miss=0.1)
    mask = (((cx - 45000) / 40000) ** 2 + ((cy - 40000) / 30000) ** 2) <= 1.0
    out['d_sk_ok_mask_rows'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx[mask], cell_ys=cy[mask],
        grid_shape=(16, 18), cntn_idxs=mask, fld_beg_row=3, fld_end_row=11,
        vgs=[VG] * 4,
        interp_args=[('OK', None, 'OK'), ('SK', None, 'SK'), ('IDW', None, 'IDW_000', 2.0)])

    # E: nearest-neighbour selection ('nrst'), many cell groups
    rng, sx, sy, data, cx, cy = make_inputs(4, 40, 4, 10, 10, miss=0.1)
    out['e_nrst'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(10, 10),
        vgs=[VG] * 4, neb_sel_mthd='nrst', n_nebs=8,
        interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0), ('NNB', None, 'NNB')])

    # H: 'pie' neighbour selection (sector round-robin).  The reference's caller passes
    # uint32 work arrays to a Cython signature that wants `unsigned long` (grps.py:173,176
    # vs pyx:822-823), which only matches on LLP64 (Windows); main() installs a shim that
    # casts those two arrays at the call boundary, the reference code itself is untouched.
    rng, sx, sy, data, cx, cy = make_inputs(8, 36, 4, 9, 11, miss=0.1)
    out['h_pie'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(9, 11),
        vgs=[VG] * 4, neb_sel_mthd='pie', n_nebs=10, n_pies=4,
        interp_args=[('OK', None, 'OK'), ('IDW', None, 'IDW_000', 2.0), ('NNB', None, 'NNB')])

    # F: every variogram family, min_vg_val > 0
    rng, sx, sy, data, cx, cy = make_inputs(5, 18, 7, 9, 11)
    vgs = ['0.2 Nug(0.0) + 0.5 Exp(30000) + 0.3 Gau(15000)',
           '0.15 Nug(0.0) + 0.85 Lin(40000)',
           '0.1 Nug(0.0) + 0.002 Pow(0.5)',
           '0.3 Nug(0.0) + 0.7 Hol(25000)',
           '0.05 Nug(0.0) + 0.95 Gau(20000)',
           '1.0 Exp(20000)',
           '0.00001 Nug(0.0) + 0.00002 Sph(20000)']
    out['f_vg_families'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(9, 11),
        vgs=vgs, min_vg_val=1e-4,
        interp_args=[('OK', None, 'OK')])

    # G: IDW only, no variograms (vgs_ser None route), integer + fractional exponents
    rng, sx, sy, data, cx, cy = make_inputs(6, 35, 6, 8, 9, miss=0.25)
    out['g_idw_only'] = base_case(
        stn_xs=sx, stn_ys=sy, data=data, cell_xs=cx, cell_ys=cy, grid_shape=(8, 9),
        interp_args=[('IDW', None, 'IDW_000', 1.0), ('IDW', None, 'IDW_001', 2.0),
                     ('IDW', None, 'IDW_002', 3.0), ('IDW', None, 'IDW_003', 5.0),
                     ('IDW', None, 'IDW_004', 2.5)])
    return out


def kats(im, misc):
    """Known-answer vectors of the Cython free functions."""
    rng = np.random.default_rng(99)
    k = {}
    x1, y1 = rng.uniform(0, 1e5, 7), rng.uniform(0, 1e5, 7)
    x2, y2 = rng.uniform(0, 1e5, 5), rng.uniform(0, 1e5, 5)
    d = np.full((7, 5), np.nan)
    im.fill_dists_2d_mat(x1, y1, x2, y2, d)
    k.update(d_x1=x1, d_y1=y1, d_x2=x2, d_y2=y2, d_out=d)
    dd = np.full((7, 7), np.nan)
    im.fill_dists_2d_mat(x1, y1, x1, y1, dd)
    vg_list = ['0.1 Nug(0.0) + 0.9 Sph(20000)', '0.2 Nug(0.0) + 0.5 Exp(30000) + 0.3 Gau(15000)',
               '0.15 Nug(0.0) + 0.85 Lin(40000)', '0.1 Nug(0.0) + 0.002 Pow(0.5)',
               '0.3 Nug(0.0) + 0.7 Hol(25000)', '2.5 Rng(1.0)', '0.5 Sph(0.0)']
    k['vg_list'] = np.array(vg_list)
    for vi, vg in enumerate(vg_list):
        for cov in (0, 1):
            for mv in (0.0, 0.3):
                a = np.full_like(d, np.nan)
                im.fill_vg_var_arr(d, a, cov, 0, vg, mv)
                k[f'vg{vi}_c{cov}_m{int(mv > 0)}_rect'] = a
                b = np.full_like(dd, np.nan)
                im.fill_vg_var_arr(dd, b, cov, 1, vg, mv)
                k[f'vg{vi}_c{cov}_m{int(mv > 0)}_diag'] = b
    h = np.linspace(0, 1e6, 10)
    k['theo_h'] = h
    k['theo_out'] = misc.get_theo_vg_vals('100 Sph(10000) + 10 Exp(1000000)', h)
    dist = np.array([0.2, 0.5, 1.0])
    w = np.full(3, np.nan)
    k['idw_sum'] = np.array(im.fill_wts_and_sum(dist, w, 2.0))
    k['idw_w'] = w
    k['idw_ms'] = np.array(im.get_mults_sum(w, np.array([1.0, 2.0, 4.0])))
    arr = rng.normal(size=(6, 8))
    ri = np.array([4, 0, 5], dtype=np.int64)
    ci = np.array([7, 7, 1, 2], dtype=np.int64)
    sub = np.full((4, 6), np.nan)
    im.copy_2d_arr_at_idxs(arr, ri, ci, sub)
    k.update(cp_arr=arr, cp_ri=ri, cp_ci=ci, cp_out=sub)
    for s in ['0.0 Nug(0.0)', '0.1 Nug(0.0) + 0.9 Sph(20000)', '0.00001 Nug(0.0) + 0.00002 Sph(20000)']:
        k['nugget__' + s] = np.array(misc.check_full_nuggetness(s, 1e-4))

    # stand-alone kriging classes (pyx:251-765)
    n_in, n_out = 14, 9
    xi, yi = rng.uniform(0, 100, n_in), rng.uniform(0, 100, n_in)
    zi = rng.gamma(1.0, 5.0, n_in)
    xk, yk = rng.uniform(0, 100, n_out), rng.uniform(0, 100, n_out)
    xk[0], yk[0] = xi[3], yi[3]                       # a target on top of a station
    si = np.vstack([50 + 0.5 * xi + 0.1 * yi, np.cos(xi / 30.0)])
    sk = np.vstack([50 + 0.5 * xk + 0.1 * yk, np.cos(xk / 30.0)])
    model = '0.1 Nug(0.0) + 0.9 Sph(60)'
    k.update(kc_xi=xi, kc_yi=yi, kc_zi=zi, kc_xk=xk, kc_yk=yk, kc_si=si, kc_sk=sk,
             kc_model=np.array(model))
    c = im.OrdinaryKriging(xi, yi, zi, xk, yk, model); c.krige()
    k.update(ok_zk=c.zk, ok_lambdas=c.lambdas, ok_mus=c.mus, ok_est_vars=c.est_vars,
             ok_rhss=c.rhss, ok_in_vars=c.in_vars)
    c = im.SimpleKriging(xi, yi, zi, xk, yk, model); c.krige()
    k.update(sk_zk=c.zk, sk_lambdas=c.lambdas, sk_est_covars=c.est_covars, sk_rhss=c.rhss,
             sk_in_covars=c.in_covars)
    c = im.ExternalDriftKriging(xi, yi, zi, si[0], xk, yk, sk[0], model); c.krige()
    k.update(edk_zk=c.zk, edk_lambdas=c.lambdas, edk_mus_1=c.mus_1, edk_mus_2=c.mus_2)
    c = im.ExternalDriftKriging_MD(xi, yi, zi, si, xk, yk, sk, model); c.krige()
    k.update(md_zk=c.zk, md_lambdas=c.lambdas, md_mus_arr=c.mus_arr)
    c = im.OrdinaryIndicatorKriging(xi, yi, zi, xk, yk, 4.0, model); c.ikrige()
    k.update(oik_ik=c.ik, oik_est_vars=c.est_vars)
    c = im.SimpleIndicatorKriging(xi, yi, zi, xk, yk, 4.0, model); c.ikrige()
    k.update(sik_ik=c.ik, sik_est_covars=c.est_covars)
    # pie helper (pyx:811-890; DT_UL = unsigned long, 64-bit here) and get_nd_dists (:893)
    n_ref = 23
    rx, ry = rng.uniform(0, 1e5, n_ref), rng.uniform(0, 1e5, n_ref)
    rx[5], ry[6] = 4.0e4, 5.5e4          # stations due north/south and east/west of a target
    pts = np.array([[4.0e4, 5.5e4], [1.2e4, 9.0e4], [7.7e4, 3.1e4], [-5.0e3, 2.0e4], [5.0e4, 5.0e4]])
    k.update(pie_rx=rx, pie_ry=ry, pie_pts=pts)
    for n_pies in (3, 4, 8):
        sel_all, pidx_all, cts_all, d_all = [], [], [], []
        for px, py in pts:
            dists = np.zeros(n_ref)
            tem = np.zeros(n_ref)
            sel = np.zeros(n_ref, dtype=np.int64)
            pidx = np.zeros(n_ref, dtype=np.uint64)
            cts = np.zeros(n_pies, dtype=np.uint64)
            im.sel_equidist_refs(px, py, rx, ry, n_pies, -1.0, -1, dists, tem, sel, pidx, cts)
            sel_all.append(sel); pidx_all.append(pidx); cts_all.append(cts); d_all.append(dists)
        k[f'pie{n_pies}_sel'] = np.array(sel_all)
        k[f'pie{n_pies}_pidx'] = np.array(pidx_all).astype(np.int64)
        k[f'pie{n_pies}_cts'] = np.array(cts_all).astype(np.int64)
        k[f'pie{n_pies}_dists'] = np.array(d_all)
    nd = rng.normal(size=(9, 3))
    k.update(nd_pts=nd, nd_out=np.asarray(im.get_nd_dists(nd)))
    # the survey's known answer
    c = im.OrdinaryKriging(np.array([0., 10., 0.]), np.array([0., 0., 10.]), np.array([1., 2., 4.]),
                           np.array([5., 2.]), np.array([5., 1.]), '0.1 Nug(0.0) + 0.9 Sph(20)')
    c.krige()
    k.update(svy_zk=c.zk, svy_lambdas=c.lambdas, svy_mus=c.mus, svy_est_vars=c.est_vars)
    return k


def save_case(name, case, flds):
    d = {}
    for key, val in case.items():
        if key == 'interp_args':
            d['ia_types'] = np.array([a[0] for a in val])
            d['ia_labels'] = np.array([a[2] for a in val])
            d['ia_exps'] = np.array([a[3] if len(a) > 3 else np.nan for a in val])
        elif val is None:
            continue
        elif key == 'vgs':
            d['vgs'] = np.array(list(val))
        else:
            d[key] = np.asarray(val)
    for lab, arr in flds.items():
        d['out__' + lab] = arr
    np.savez_compressed(OUT / f'{name}.npz', **d)


def install_pie_shim(im):
    """grps.py calls sel_equidist_refs with uint32 work arrays; the compiled signature
    takes `unsigned long` buffers (64-bit on LP64).  Cast the two arrays going in and copy
    them back, nothing else changes."""
    import spinterps.interp.grps as grps
    real = im.sel_equidist_refs

    def shim(dst_x, dst_y, ref_xs, ref_ys, n_pies, thr, flag, dists, tem, sel, pidx, cts):
        p64, c64 = pidx.astype(np.uint64), cts.astype(np.uint64)
        real(dst_x, dst_y, ref_xs, ref_ys, n_pies, thr, flag, dists, tem, sel, p64, c64)
        pidx[:] = p64
        cts[:] = c64
    grps.sel_equidist_refs = shim


def main():
    im, SpInterpSteps, misc = import_reference()
    install_pie_shim(im)
    np.savez_compressed(OUT / 'kats.npz', **kats(im, misc))
    if '--kats-only' in sys.argv:
        return
    only = [a.split('=', 1)[1] for a in sys.argv if a.startswith('--only=')]
    for name, case in cases().items():
        if only and name not in only:
            continue
        flds = run_reference(SpInterpSteps, case)
        save_case(name, case, flds)
        msg = ', '.join(f'{lab}[{np.nanmin(a):.6g},{np.nanmax(a):.6g}] nan={np.isnan(a).sum()}'
                        for lab, a in flds.items())
        print(name, msg)


if __name__ == '__main__':
    if '--out' in sys.argv:
        OUT = Path(sys.argv[sys.argv.index('--out') + 1])
        OUT.mkdir(parents=True, exist_ok=True)
    main()
