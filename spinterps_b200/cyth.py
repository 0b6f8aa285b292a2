"""Drop-in mirror of the reference's native module ``spinterps.cyth``
(cyth/interpmthds.pyx free functions) on top of the sm_100a kernels.

Same names, argument order and in-place semantics as the Cython ``cpdef``
functions, NumPy arrays in host memory; every call goes through the C-ABI
(include/spx_b200.h, group 1) and therefore through the GPU.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _c64(a, name, ndim):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == ndim
            and a.flags.c_contiguous):
        raise ValueError(f'Buffer dtype mismatch / not C-contiguous float64 [{ndim}D]: {name}')
    return a


def fill_dists_2d_mat(x1s, y1s, x2s, y2s, dists):
    """cyth/interpmthds.pyx:123-143."""
    for n, a in (('x1s', x1s), ('y1s', y1s), ('x2s', x2s), ('y2s', y2s)):
        _c64(a, n, 1)
    _c64(dists, 'dists', 2)
    assert x1s.size == y1s.size and x2s.size == y2s.size
    assert dists.shape == (x1s.size, x2s.size)
    lib = _lib.load()
    _lib.check(lib.spx_fill_dists_2d_mat(
        _lib.f64p(x1s), _lib.f64p(y1s), x1s.size, _lib.f64p(x2s), _lib.f64p(y2s), x2s.size,
        _lib.f64p(dists)), 'fill_dists_2d_mat')


def fill_vg_var_arr(dists, in_vars, covar_flag, diag_mat_flag, vg_models_str, min_vg_val):
    """cyth/interpmthds.pyx:146-226."""
    _c64(dists, 'dists', 2)
    _c64(in_vars, 'in_vars', 2)
    assert dists.shape == in_vars.shape
    lib = _lib.load()
    _lib.check(lib.spx_fill_vg_var_arr(
        _lib.f64p(dists), _lib.f64p(in_vars), dists.shape[0], dists.shape[1],
        int(covar_flag), int(diag_mat_flag), str(vg_models_str).encode(), float(min_vg_val)),
        'fill_vg_var_arr')


def copy_2d_arr_at_idxs(arr, row_idxs, col_idxs, subset_arr):
    """cyth/interpmthds.pyx:229-248 (index arrays are ``long long``)."""
    _c64(arr, 'arr', 2)
    _c64(subset_arr, 'subset_arr', 2)
    for n, a in (('row_idxs', row_idxs), ('col_idxs', col_idxs)):
        if not (isinstance(a, np.ndarray) and a.dtype == np.int64 and a.ndim == 1
                and a.flags.c_contiguous):
            raise ValueError(f'Buffer dtype mismatch, expected int64 1D: {n}')
    lib = _lib.load()
    _lib.check(lib.spx_copy_2d_arr_at_idxs(
        _lib.f64p(arr), arr.shape[0], arr.shape[1], _lib.i64p(row_idxs), row_idxs.size,
        _lib.i64p(col_idxs), col_idxs.size, _lib.f64p(subset_arr), subset_arr.shape[0],
        subset_arr.shape[1]), 'copy_2d_arr_at_idxs')


def fill_theo_vg_vals(vg_str, h_arr, r, s, vg_arr):
    """cyth/interpmthds.pyx:98-120 (accumulates into vg_arr)."""
    h = np.ascontiguousarray(h_arr, dtype=np.float64)
    out = np.ascontiguousarray(vg_arr, dtype=np.float64)
    assert h.shape[0] and h.shape[0] == out.shape[0]
    assert s >= 0 and r >= 0
    lib = _lib.load()
    _lib.check(lib.spx_fill_theo_vg_vals(
        str(vg_str).encode(), _lib.f64p(h), h.size, float(r), float(s), _lib.f64p(out)),
        'fill_theo_vg_vals')
    if out is not vg_arr:
        vg_arr[...] = out


def fill_dists_one_pt(x, y, xs, ys, dists):
    """cyth/interpmthds.pyx:768-781."""
    _c64(xs, 'xs', 1), _c64(ys, 'ys', 1), _c64(dists, 'dists', 1)
    lib = _lib.load()
    _lib.check(lib.spx_fill_dists_one_pt(
        float(x), float(y), _lib.f64p(xs), _lib.f64p(ys), xs.size, _lib.f64p(dists)),
        'fill_dists_one_pt')


def fill_wts_and_sum(dists, wts, idw_exp):
    """cyth/interpmthds.pyx:784-795 -> float."""
    _c64(dists, 'dists', 1), _c64(wts, 'wts', 1)
    out = C.c_double(0.0)
    lib = _lib.load()
    _lib.check(lib.spx_fill_wts_and_sum(
        _lib.f64p(dists), _lib.f64p(wts), dists.size, float(idw_exp), C.byref(out)),
        'fill_wts_and_sum')
    return out.value


def get_mults_sum(wts, data):
    """cyth/interpmthds.pyx:798-808 -> float."""
    _c64(wts, 'wts', 1), _c64(data, 'data', 1)
    out = C.c_double(0.0)
    lib = _lib.load()
    _lib.check(lib.spx_get_mults_sum(_lib.f64p(wts), _lib.f64p(data), wts.size, C.byref(out)),
               'get_mults_sum')
    return out.value


def get_theo_vg_vals(in_model, h_arr):
    """misc.py:1027-1047 on top of fill_theo_vg_vals (no range clamp there)."""
    h = np.ascontiguousarray(h_arr, dtype=np.float64)
    vals = np.zeros_like(h)
    for t, s, r in _lib.parse_vg_str(in_model, clamp_range=False):
        fill_theo_vg_vals(_lib.VG_NAMES[t], h, r, s, vals)
    return vals


def sel_equidist_refs(dst_x, dst_y, ref_xs, ref_ys, n_pies, min_dist_thresh, not_neb_flag,
                      dists, tem_ref_sel_dists, ref_sel_pie_idxs, ref_pie_idxs, ref_pie_cts):
    """cyth/interpmthds.pyx:811-890.  ``ref_pie_idxs`` / ``ref_pie_cts`` are the
    reference's ``unsigned long`` buffers: uint64 on LP64."""
    _c64(ref_xs, 'ref_xs', 1), _c64(ref_ys, 'ref_ys', 1), _c64(dists, 'dists', 1)
    _c64(tem_ref_sel_dists, 'tem_ref_sel_dists', 1)
    for name, a, dt in (('ref_sel_pie_idxs', ref_sel_pie_idxs, np.int64),
                        ('ref_pie_idxs', ref_pie_idxs, np.uint64),
                        ('ref_pie_cts', ref_pie_cts, np.uint64)):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.ndim == 1
                and a.flags.c_contiguous):
            raise ValueError(f'Buffer dtype mismatch, expected {np.dtype(dt).name}: {name}')
    n = ref_xs.size
    assert ref_ys.size == dists.size == ref_sel_pie_idxs.size == ref_pie_idxs.size == n
    assert ref_pie_cts.size >= n_pies
    lib = _lib.load()
    _lib.check(lib.spx_sel_equidist_refs(
        float(dst_x), float(dst_y), ref_xs.ctypes.data, ref_ys.ctypes.data, n, int(n_pies),
        float(min_dist_thresh), int(not_neb_flag), dists.ctypes.data,
        tem_ref_sel_dists.ctypes.data, ref_sel_pie_idxs.ctypes.data, ref_pie_idxs.ctypes.data,
        ref_pie_cts.ctypes.data), 'sel_equidist_refs')


def get_nd_dists(pts):
    """cyth/interpmthds.pyx:893-925 -> ndarray[n (n - 1) / 2]."""
    _c64(pts, 'pts', 2)
    n = pts.shape[0]
    out = np.full((n * (n - 1)) // 2, np.nan, dtype=np.float64)
    lib = _lib.load()
    _lib.check(lib.spx_get_nd_dists(pts.ctypes.data, n, pts.shape[1], out.ctypes.data),
               'get_nd_dists')
    return out
