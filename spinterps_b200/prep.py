"""Grid preparation on the GPU.

* ``points_in_polygons``: which cells / stations lie inside or within a buffer distance of
  the selection polygons -- the reference buffers the polygons with OGR and calls
  ``Contains`` once per point from Python (misc.py:407-540 ``chk_pt_cntmnt_in_polys_mp``,
  used by interp/bdpolys.py:148 for stations and interp/prepare.py:266 for cells).
* ``sample_raster`` / ``drift_at_cells`` / ``drift_at_points``: drift values at cells and
  stations (interp/drift.py:165-226).

Polygons are passed as arrays (rings) and rasters as arrays + geometry; gisio.py reads them
from ESRI shapefiles / ASCII grids, other formats need GDAL, which is outside this path.  torch is used for device memory only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _to_dev(a, dev):
    if not a.flags.writeable:          # e.g. a read-only pandas view
        a = a.copy()
    return torch.from_numpy(a).to(dev)


def rings_to_edges(rings):
    """[(n_i, 2) outer rings] -> edge arrays x1, y1, x2, y2 (float64), ring id per edge
    (int32, non-decreasing).  A ring is closed implicitly (last -> first vertex); a
    repeated closing vertex just adds a zero-length edge, which is skipped."""
    x1, y1, x2, y2, rid = [], [], [], [], []
    for k, ring in enumerate(rings):
        r = np.asarray(ring, dtype=np.float64)
        assert r.ndim == 2 and r.shape[1] == 2, 'a ring is an (n, 2) array of x, y'
        if r.shape[0] >= 2 and np.array_equal(r[0], r[-1]):
            r = r[:-1]
        assert r.shape[0] >= 3, f'Polygon not having enough points ({r.shape[0]})!'  # misc.py:415
        nxt = np.roll(r, -1, axis=0)
        x1.append(r[:, 0]); y1.append(r[:, 1]); x2.append(nxt[:, 0]); y2.append(nxt[:, 1])
        rid.append(np.full(r.shape[0], k, dtype=np.int32))
    cat = lambda v, dt: np.ascontiguousarray(np.concatenate(v), dtype=dt)  # noqa: E731
    return (cat(x1, np.float64), cat(y1, np.float64), cat(x2, np.float64), cat(y2, np.float64),
            cat(rid, np.int32))


def points_in_polygons(xs, ys, rings, buffer_dist=0.0, device=None):
    """bool [n]: point inside any ring (even-odd rule) or, with buffer_dist > 0, closer than
    buffer_dist to any ring edge.  One kernel launch over all points
    (spx_points_in_polygons_dev)."""
    _lib.require_gpu()
    lib = _lib.load()
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    assert xs.shape == ys.shape
    assert np.isfinite(buffer_dist) and buffer_dist >= 0
    ex1, ey1, ex2, ey2, rid = rings_to_edges(rings)
    ch = lib.spx_points_in_polygons_chunk()
    n_e = ex1.size
    starts = np.arange(0, n_e, ch)
    cymin = np.minimum(np.minimum.reduceat(ey1, starts), np.minimum.reduceat(ey2, starts))
    cymax = np.maximum(np.maximum.reduceat(ey1, starts), np.maximum.reduceat(ey2, starts))
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
    with torch.cuda.device(dev):
        t = [_to_dev(a, dev) for a in (xs, ys, ex1, ey1, ex2, ey2, rid, cymin, cymax)]
        out = torch.empty(xs.size, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.spx_points_in_polygons_dev(
            t[0].data_ptr(), t[1].data_ptr(), xs.size, t[2].data_ptr(), t[3].data_ptr(),
            t[4].data_ptr(), t[5].data_ptr(), t[6].data_ptr(), n_e, t[7].data_ptr(),
            t[8].data_ptr(), float(buffer_dist), out.data_ptr(), st), 'points_in_polygons')
        return out.cpu().numpy().astype(bool)


def sample_raster(ras, rows, cols, ndv=None, device=None):
    """ras[rows, cols] with no-data (np.isclose to ndv) and out-of-raster -> NaN."""
    _lib.require_gpu()
    lib = _lib.load()
    ras = np.ascontiguousarray(ras, dtype=np.float64)
    assert ras.ndim == 2
    rows = np.ascontiguousarray(rows, dtype=np.int64).ravel()
    cols = np.ascontiguousarray(cols, dtype=np.int64).ravel()
    assert rows.shape == cols.shape
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
    with torch.cuda.device(dev):
        d_ras, d_r, d_c = (_to_dev(a, dev) for a in (ras, rows, cols))
        out = torch.empty(rows.size, dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.spx_sample_raster_dev(
            d_ras.data_ptr(), ras.shape[0], ras.shape[1], d_r.data_ptr(), d_c.data_ptr(),
            rows.size, 0.0 if ndv is None else float(ndv), int(ndv is not None),
            out.data_ptr(), st), 'sample_raster')
        return out.cpu().numpy()


def drift_cell_indices(min_row, max_row, min_col, max_col, cntn_idxs=None):
    """Raster (row, col) of every interpolation cell, interp/drift.py:175-188."""
    cols = np.arange(min_col, max_col + 1, dtype=np.int64)
    rows = np.arange(min_row, max_row + 1, dtype=np.int64)
    cm, rm = np.meshgrid(cols, rows)
    cm, rm = cm.ravel(), rm.ravel()
    if cntn_idxs is not None:
        cm, rm = cm[cntn_idxs], rm[cntn_idxs]
    return rm, cm


def drift_point_indices(xs, ys, ras_x_min, ras_y_max, cell_size):
    """Raster (row, col) of points, interp/drift.py:209-210 (``int()`` truncates)."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    cols = np.trunc((xs - ras_x_min) / cell_size).astype(np.int64)
    rows = np.trunc((ras_y_max - ys) / cell_size).astype(np.int64)
    return rows, cols


def aligned_bounds(ras_x_min, ras_y_max, ras_cell_size, ras_n_rows, ras_n_cols,
                   extent=None, cell_bdist=0.0):
    """Grid bounds snapped outwards to the cell lattice of an alignment raster
    (misc.py:743-885 ``get_aligned_shp_bds_and_cell_size``; host arithmetic).

    extent: (x_min, x_max, y_min, y_max) of the selection polygons, or None -- then the
    grid is the raster's own extent (the reference's ``bounds_shp_file == 'None'``).
    Returns ((x_min, x_max, y_min, y_max), cell_size)."""
    rel_cell_err = 1e-5
    cs = float(ras_cell_size)
    abs_cell_err = abs(cs * rel_cell_err)
    ras_min_x, ras_max_y = float(ras_x_min), float(ras_y_max)
    ras_max_x = ras_min_x + (ras_n_cols * cs)                  # misc.py:657-658
    ras_min_y = ras_max_y - (ras_n_rows * cs)
    if extent is not None:
        x0, x1, y0, y1 = (float(v) for v in extent)
        if cell_bdist:
            x0 -= cell_bdist
            x1 += cell_bdist
            y0 -= cell_bdist
            y1 += cell_bdist
    else:
        x0, x1, y0, y1 = ras_min_x, ras_max_x, ras_min_y, ras_max_y
    assert not (x0 < ras_min_x) or abs(x0 - ras_min_x) <= abs_cell_err, (
        f'bounds_shp x_min ({x0}) < align_raster x_min ({ras_min_x})!')
    assert not (x1 > ras_max_x) or abs(x1 - ras_max_x) <= abs_cell_err, (
        f'bounds_shp x_max ({x1}) < align_raster x_max ({ras_max_x})!')
    assert not (y0 < ras_min_y) or abs(y0 - ras_min_y) <= abs_cell_err, (
        f'bounds_shp y_min ({y0}) < align_raster y_min ({ras_min_y})!')
    assert not (y1 > ras_max_y) or abs(y1 - ras_max_y) <= abs_cell_err, (
        f'bounds_shp y_max ({y1}) < align_raster y_max ({ras_max_y})!')

    def near(a, b):
        return bool(np.isclose(a, b, rtol=0, atol=rel_cell_err))

    # west / north edges move out to the lattice line at or before them, east / south edges
    # to the line after them (one whole cell further when they sit on a line already)
    ax0 = ras_min_x if near(x0, ras_min_x) else x0 - (((x0 - ras_min_x) / cs) % 1) * cs
    ay1 = ras_max_y if near(y1, ras_max_y) else y1 + (((ras_max_y - y1) / cs) % 1) * cs
    ax1 = ras_max_x if near(x1, ras_max_x) else x1 + (cs - (((x1 - ras_min_x) / cs) % 1) * cs)
    ay0 = ras_min_y if near(y0, ras_min_y) else y0 - (cs - (((ras_max_y - y0) / cs) % 1) * cs)
    for rem in ((ax1 - ax0) % cs, (ay1 - ay0) % cs):
        assert near(rem, 0.0) or near(rem, cs), 'adjusted bounds off the alignment lattice'
    assert ax0 >= ras_min_x, f'Adjusted bounds_shp x_min ({ax0}) < align_raster x_min ({ras_min_x})!'
    assert ax1 <= ras_max_x, f'Adjusted bounds_shp x_max ({ax1}) < align_raster x_max ({ras_max_x})!'
    assert ay0 >= ras_min_y, f'Adjusted bounds_shp y_min ({ay0}) < align_raster y_min ({ras_min_y})!'
    assert ay1 <= ras_max_y, f'Adjusted bounds_shp y_max ({ay1}) < align_raster y_max ({ras_max_y})!'
    return (ax0, ax1, ay0, ay1), cs

