"""``SpInterpMain``: the reference's configuration / verify() / interpolate()
surface (interp/main.py:36-236, interp/data.py, variograms/vgsinput.py:28-185,
interp/prepare.py:510-729) in front of the GPU engine.

Same setter names, argument meaning, ``assert``-style validation and call-order
flags.  Differences, all forced by what exists in this image:

* GIS inputs: the reference reads them with GDAL/OGR, which is not installed.  The
  path-taking setters read the simple file formats themselves (gisio.py: ESRI
  shapefile polygons for ``set_cell_selection_parameters``, ESRI ASCII grids and
  single-band GeoTIFFs for ``set_alignment_raster`` / ``turn_external_drift_kriging_on``),
  and array-level forms take data read by anything else:
  ``set_cell_selection_polygons`` (rings), ``set_cell_selection_mask`` (callable or bool
  array -> ``_cntn_idxs``), dict rasters (values + geometry) and callables ``f(x, y)`` as
  drift "rasters".  Containment, buffers and raster sampling run on the GPU (prep.py).
* The scheduler (``_get_thread_steps_idxs``: RAM-driven time / grid-row chunks
  mapped over a process pool) is replaced by time chunks sized for HBM, pipelined
  on one GPU, and by time-sharding across ranks when ``torch.distributed`` is
  initialised (spinterps_b200/dist.py).
* ``turn_simple_kriging_on`` works (the reference raises "SK is deprecated!",
  interp/main.py:263-265, while its compute path still supports ``_spk_flag``).
"""
from __future__ import annotations

import timeit
from math import ceil, floor
from pathlib import Path
from threading import Lock

import numpy as np
import pandas as pd

from . import ncwriter
from .steps import InterpFields, SpInterpSteps


def _print_sl():
    print(2 * '\n', 50 * '#', sep='')


def _print_el():
    print(50 * '#', 2 * '\n', sep='')


class SpInterpMain:

    _neb_sel_mthds = ('all', 'nrst', 'pie')

    def __init__(self, verbose=True):
        assert isinstance(verbose, bool), 'verbose can only be a boolean!'
        self._vb = verbose

        # data (variograms/vgsinput.py:14-26)
        self._data_df = None
        self._crds_df = None
        self._index_type = None
        self._stns_min_dist_thrsh = 0.0
        self._data_set_flag = False

        # interp/data.py:18-74
        self._vgs_ser = None
        self._out_dir = None
        self._nc_out = None
        self._nc_vunits = self._nc_vlab = None
        self._nc_tunits = self._nc_tcldr = None
        self._nc_nmrl_prcn = None
        self._nc_cprm_levl = None
        self._tbeg = self._tend = self._tfreq = None
        self._poly_shp = None
        self._ipoly_flag = False
        self._stn_bdist = None
        self._cell_bdist = 0.0
        self._poly_simplify_tol_ratio = 0.0
        self._algn_ras = None
        self._neb_sel_mthd = 'all'
        self._n_nebs = None
        self._n_pies = None
        self._n_cpus = 1
        self._mp_flag = False
        self._plot_figs_flag = False
        self._cell_size = None
        self._min_var_thr = -np.inf
        self._min_var_cut = None
        self._max_var_cut = None
        self._max_steps_per_chunk = None
        self._min_vg_val = 0.0
        self._cell_mask_src = None
        self._poly_rings = None

        self._vg_ser_set_flag = False
        self._out_dir_set_flag = False
        self._nc_set_flag = False
        self._time_prms_set_flag = False
        self._cell_sel_prms_set = False
        self._algn_ras_set_flag = False
        self._neb_sel_mthd_set_flag = False
        self._misc_settings_set_flag = False
        self._data_vrfd_flag = False

        # interp/prepare.py:26-43
        self._cntn_idxs = None
        self._drft_arrs = None
        self._stns_drft_df = None
        self._prpd_flag = False
        self._intrp_dtype = np.float32
        self._nc_xlab, self._nc_ylab, self._nc_tlab = 'X', 'Y', 'time'

        # interp/main.py:36-61
        self._drft_rass = None
        self._n_drft_rass = None
        self._idw_exps = None
        self._n_idw_exps = None
        self._ork_flag = False
        self._spk_flag = False
        self._edk_flag = False
        self._idw_flag = False
        self._nnb_flag = False
        self._interp_flag_est_vars = False
        self._main_vrfd_flag = False
        self._engine = None

    # ------------------------------------------------------------------ data
    def set_data(self, stns_time_ser_df, stns_crds_df, index_type='date',
                 stns_min_dist_thresh=0):
        """variograms/vgsinput.py:28-185."""
        assert isinstance(stns_time_ser_df, pd.DataFrame), (
            'stns_time_ser_df has to be a pd.DataFrame object!')
        assert isinstance(stns_crds_df, pd.DataFrame), (
            'stns_crds_df has to be a pd.DataFrame object!')
        assert all(stns_time_ser_df.shape), 'Empty stns_time_ser_df!'
        assert all(stns_crds_df.shape), 'Empty stns_crds_df!'
        assert np.issubdtype(stns_time_ser_df.values.dtype, np.floating), (
            'dtype of stns_time_ser_df should be a subtype of np.floating!')
        assert np.issubdtype(stns_crds_df.values.dtype, np.number), (
            'dtype of stns_crds_df should be a subtype of np.number!')
        if self._index_type is not None:
            assert index_type == self._index_type, (
                'Given and previously set index_type do not match!')
        if index_type == 'date':
            assert isinstance(stns_time_ser_df.index, pd.DatetimeIndex), (
                'Data type of index of stns_time_ser_df does not match index_type!')
        elif index_type == 'obj':
            pass
        else:
            raise AssertionError('index_type can only be \'obj\' or \'date\'!')
        assert all(c in stns_crds_df.columns for c in ('X', 'Y')), (
            'stns_crds_df has a missing \'X\' or \'Y\' column!')
        assert isinstance(stns_min_dist_thresh, (float, int)), (
            'stns_min_dist_thresh can only be a float or an int!')
        assert 0 <= stns_min_dist_thresh < np.inf

        data_df = stns_time_ser_df.copy()
        data_df.columns = [str(c) for c in data_df.columns]
        crds_df = stns_crds_df.loc[:, ['X', 'Y']].astype(float).copy()
        crds_df.index = [str(i) for i in crds_df.index]
        crds_df = crds_df[~crds_df.index.duplicated(keep='last')]
        crds_df = crds_df.dropna(axis=0, how='any')
        data_df = data_df.dropna(axis=1, how='all')

        if stns_min_dist_thresh > 0:  # drop the later station of every too-close pair
            xs, ys = crds_df['X'].values, crds_df['Y'].values
            dists = np.hypot(xs[:, None] - xs[None, :], ys[:, None] - ys[None, :])
            keep = np.ones(xs.size, dtype=bool)
            for i in range(xs.size):
                if keep[i]:
                    close = (dists[i] < stns_min_dist_thresh) & (np.arange(xs.size) > i)
                    keep[close] = False
            crds_df = crds_df.loc[keep]

        cmn = data_df.columns.intersection(crds_df.index)
        assert cmn.shape[0] > 1, 'Less than 2 common stations in data and coordinates!'
        self._data_df = data_df.loc[:, cmn]
        self._crds_df = crds_df.loc[cmn]
        self._index_type = index_type
        self._stns_min_dist_thrsh = float(stns_min_dist_thresh)
        if self._vb:
            _print_sl()
            print('Set data with', self._data_df.shape[0], 'steps and', cmn.shape[0], 'stations.')
            _print_el()
        self._data_set_flag = True

    def set_vgs_ser(self, vgs_ser, index_type='date'):
        """interp/data.py:77-130."""
        assert isinstance(vgs_ser, pd.Series), 'vgs_ser has to be a pd.Series object!'
        if self._index_type is not None:
            assert index_type == self._index_type, (
                'Given and previously set index_type do not match!')
        if index_type == 'date':
            assert isinstance(vgs_ser.index, pd.DatetimeIndex), (
                'Data type of index of vgs_ser does not match index_type!')
        elif index_type != 'obj':
            raise AssertionError('index_type can only be \'obj\' or \'date\'!')
        assert all(vgs_ser.shape), 'Empty vgs_ser!'
        self._vgs_ser = vgs_ser.astype(object)
        self._index_type = index_type
        self._vg_ser_set_flag = True

    def set_out_dir(self, out_dir):
        """interp/data.py:132-152."""
        assert isinstance(out_dir, (str, Path)), 'out_dir can only be a string or pathlib.Path!'
        out_dir = Path(out_dir).absolute()
        assert out_dir.parents[0].exists(), 'Parent directory of the out_dir does not exist!'
        self._out_dir = out_dir
        self._out_dir_set_flag = True

    def set_netcdf4_parameters(self, out_file_name, var_units, var_label, time_units,
                               time_calendar, nmrl_prcn, cprm_levl):
        """interp/data.py:154-267."""
        assert isinstance(out_file_name, str), 'out_file_name not a string!'
        assert out_file_name, 'Empty out_file_name!'
        assert isinstance(var_units, str), 'var_units not a string!'
        assert isinstance(var_label, str), 'var_label not a string!'
        if self._index_type != 'obj':
            assert isinstance(time_units, str), 'time_units not a string!'
            assert time_units, 'Empty time_units string!'
            assert isinstance(time_calendar, str), 'time_calendar not a string!'
            assert time_calendar, 'Empty time_calendar string!'
        assert isinstance(nmrl_prcn, int), 'nmrl_prcn must be an integer!'
        assert nmrl_prcn >= 0, 'nmrl_prcn must be greater than zero!'
        assert isinstance(cprm_levl, int), 'cprm_levl must be an integer!'
        assert 0 <= cprm_levl <= 9, 'cprm_levl must be between 0 and 9!'
        self._nc_out = out_file_name
        self._nc_vunits = var_units
        self._nc_vlab = var_label
        self._nc_tunits = time_units
        self._nc_tcldr = time_calendar
        self._nc_nmrl_prcn = nmrl_prcn
        self._nc_cprm_levl = cprm_levl
        self._nc_set_flag = True

    def set_interp_time_parameters(self, beg_time, end_time, time_freq, time_fmt):
        """interp/data.py:269-347."""
        if time_fmt is not None:
            assert isinstance(time_fmt, str), 'time_fmt not a string!'
            assert isinstance(time_freq, str), 'time-_freq not a string!'
            self._tfreq = time_freq
            vals = []
            for t in (beg_time, end_time):
                if isinstance(t, str):
                    vals.append(pd.to_datetime(t, format=time_fmt))
                elif isinstance(t, pd.Timestamp):
                    vals.append(t)
                else:
                    raise AssertionError(
                        'beg_time / end_time can only be an str or a pd.Timestamp object!')
            self._tbeg, self._tend = vals
            assert self._tend >= self._tbeg, (
                'Begining time of interpolation cannot be less than the ending time!')
        self._time_prms_set_flag = True

    def set_cell_selection_parameters(self, polygons_shapefile, station_select_buffer_distance,
                                      interp_around_polys_flag=True,
                                      polygon_cell_buffer_distance=None,
                                      simplify_tolerance_ratio=0.0):
        """interp/data.py:349-461.  The polygons are read from the ``.shp`` file without OGR
        (gisio.read_shp_polygons: every ring becomes a polygon, like misc.py:221-286) and
        handed to ``set_cell_selection_polygons``.  ``simplify_tolerance_ratio`` > 0 would
        need GEOS (``SimplifyPreserveTopology``)."""
        assert isinstance(polygons_shapefile, (str, Path)), (
            'polygons_shapefile has to be a string or a pathlib.Path object!')
        polygons_shapefile = Path(polygons_shapefile).absolute()
        assert polygons_shapefile.exists(), 'polygons_shapefile does not exist!'
        assert polygons_shapefile.is_file(), 'polygons_shapefile is not a file!'
        assert isinstance(simplify_tolerance_ratio, float), 'simplify_tolerance_ratio not a float!'
        assert simplify_tolerance_ratio >= 0, 'Invalid simplify_tolerance_ratio!'
        if simplify_tolerance_ratio:
            raise NotImplementedError('simplifying the polygons needs GEOS; pass 0.0')
        assert isinstance(polygon_cell_buffer_distance, (float, int)), (
            'polygon_cell_buffer_distance should be a float or an int '
            'if interp_around_polys_flag is True!')
        from . import gisio
        self.set_cell_selection_polygons(gisio.read_shp_polygons(polygons_shapefile),
                                         station_select_buffer_distance,
                                         interp_around_polys_flag, polygon_cell_buffer_distance)
        self._poly_shp = polygons_shapefile

    def set_cell_selection_polygons(self, polygons, station_select_buffer_distance,
                                    interp_around_polys_flag=True,
                                    polygon_cell_buffer_distance=None):
        """``set_cell_selection_parameters`` (interp/data.py:349-461) with the polygons
        given as arrays -- a list of (n, 2) outer rings -- instead of a shapefile.  Same
        meaning of the other arguments: stations within ``station_select_buffer_distance``
        of the polygons are kept (interp/bdpolys.py:80-170); with
        ``interp_around_polys_flag`` only cells inside or within
        ``polygon_cell_buffer_distance`` of a polygon are interpolated
        (interp/prepare.py:244-288); the grid spans the polygons' extent +- that distance
        (interp/prepare.py:107-137).  Containment runs on the GPU (prep.points_in_polygons).
        The reference buffers with OGR (arcs of 30 segments per quadrant); here the exact
        distance is used."""
        rings = [np.asarray(r, dtype=np.float64) for r in polygons]
        assert rings, 'Zero polygons in the polygons_shapefile!'
        for r in rings:
            assert r.ndim == 2 and r.shape[1] == 2 and r.shape[0] >= 3, (
                f'Polygon not having enough points ({r.shape[0]})!')
        assert isinstance(station_select_buffer_distance, (float, int)), (
            'station_select_buffer_distance not a float or an int!')
        assert 0 <= station_select_buffer_distance < np.inf, (
            'station_select_buffer_distance not in between zero and infinity!')
        assert isinstance(interp_around_polys_flag, bool), (
            'interp_around_polys_flag not a boolean!')
        if interp_around_polys_flag or polygon_cell_buffer_distance is not None:
            assert isinstance(polygon_cell_buffer_distance, (float, int)), (
                'polygon_cell_buffer_distance not a float or an int!')
            assert 0 <= polygon_cell_buffer_distance < np.inf, (
                'polygon_cell_buffer_distance not in between zero and infinity!')
        self._poly_rings = rings
        self._stn_bdist = float(station_select_buffer_distance)
        self._ipoly_flag = interp_around_polys_flag
        self._cell_bdist = float(polygon_cell_buffer_distance or 0.0)
        self._cell_sel_prms_set = True

    def set_cell_selection_mask(self, mask, polygon_cell_buffer_distance=0.0):
        """Array-level replacement of the polygon cell selection
        (interp/prepare.py:244-288 produces exactly such a boolean ``_cntn_idxs``).

        mask : bool array over the raveled grid, or callable(xs, ys) -> bool array.
        polygon_cell_buffer_distance : grid bounds = station extent +- this value
        (``_cell_bdist``, interp/prepare.py:126-130)."""
        assert callable(mask) or isinstance(mask, np.ndarray)
        assert isinstance(polygon_cell_buffer_distance, (float, int))
        assert 0 <= polygon_cell_buffer_distance < np.inf
        self._cell_mask_src = mask
        self._cell_bdist = float(polygon_cell_buffer_distance)
        self._ipoly_flag = True

    def set_alignment_raster(self, align_raster):
        """interp/data.py:463-494.  The grid bounds are snapped to the cell lattice of this
        raster and its cell size is used (interp/prepare.py:45-90).  Array level: a dict
        with the raster's geometry -- ``x_min``, ``y_max``, ``cell_size`` and ``n_rows`` /
        ``n_cols`` (or ``values``, whose shape gives them); a path to an ESRI ASCII grid is read
        by gisio.read_raster, any other raster file needs GDAL."""
        if isinstance(align_raster, dict):
            assert {'x_min', 'y_max', 'cell_size'} <= set(align_raster), (
                'array alignment raster needs x_min, y_max, cell_size and n_rows / n_cols')
            if 'values' in align_raster:
                n_rows, n_cols = np.asarray(align_raster['values']).shape
            else:
                n_rows, n_cols = int(align_raster['n_rows']), int(align_raster['n_cols'])
            cs = float(align_raster['cell_size'])
            assert 0 < cs < np.inf and n_rows > 0 and n_cols > 0
            self._algn_ras = dict(x_min=float(align_raster['x_min']),
                                  y_max=float(align_raster['y_max']), cell_size=cs,
                                  n_rows=int(n_rows), n_cols=int(n_cols))
            self._algn_ras_set_flag = True
            return
        assert isinstance(align_raster, (str, Path)), (
            'align_raster has to be a string or a pathlib.Path object!')
        align_raster = Path(align_raster).absolute()
        assert align_raster.exists(), 'align_raster does not exist!'
        assert align_raster.is_file(), 'align_raster is not a file!'
        from . import gisio
        self.set_alignment_raster(gisio.read_raster(align_raster))   # ESRI ASCII grid, else GDAL

    def set_neighbor_selection_method(self, selection_method, n_neighbors=None, n_pies=None):
        """interp/data.py:496-588."""
        assert isinstance(selection_method, str), 'selection_method not a string!'
        assert selection_method in self._neb_sel_mthds, (
            f'selection_method can only be one of {self._neb_sel_mthds}!')
        if selection_method in ('nrst', 'pie'):
            assert isinstance(n_neighbors, int), 'n_neighbors not an integer!'
            assert n_neighbors > 0, 'n_neighbors less than or equal to zero!'
            self._n_nebs = n_neighbors
        if selection_method == 'pie':
            assert isinstance(n_pies, int), 'n_pies not an integer!'
            assert 0 < n_pies <= n_neighbors, 'n_pies should be > 0 and <= n_neighbors!'
            self._n_pies = n_pies
        self._neb_sel_mthd = selection_method
        self._neb_sel_mthd_set_flag = True

    def set_misc_settings(self, n_cpus=1, plot_figs_flag=False, cell_size=None,
                          min_value_to_interp_thresh=-np.inf, min_cutoff_value=None,
                          max_cutoff_value=None, max_steps_per_chunk=None, min_vg_val=0.0):
        """interp/data.py:590-745.  ``n_cpus`` is accepted for compatibility; the
        GPU path does not use a process pool."""
        if isinstance(n_cpus, str):
            assert n_cpus == 'auto', 'Invalid n_cpus!'
            n_cpus = 1
        else:
            assert isinstance(n_cpus, int), 'n_cpus is not an integer!'
            assert n_cpus > 0, 'Invalid n_cpus!'
        assert isinstance(plot_figs_flag, bool), 'plot_figs_flag not a boolean!'
        if cell_size is not None:
            assert isinstance(cell_size, (int, float)), 'cell_size not a float or an int!'
            assert 0 < cell_size < np.inf, 'Invalid cell_size!'
            self._cell_size = float(cell_size)
        assert isinstance(min_value_to_interp_thresh, (int, float))
        assert -np.inf <= min_value_to_interp_thresh < np.inf
        self._min_var_thr = float(min_value_to_interp_thresh)
        if min_cutoff_value is not None:
            assert isinstance(min_cutoff_value, (int, float))
            self._min_var_cut = float(min_cutoff_value)
        if max_cutoff_value is not None:
            assert isinstance(max_cutoff_value, (int, float))
            self._max_var_cut = float(max_cutoff_value)
        if max_steps_per_chunk is not None:
            assert isinstance(max_steps_per_chunk, int), 'max_steps_per_chunk not an integer!'
            assert max_steps_per_chunk > 0, 'Invalid max_steps_per_chunk!'
            self._max_steps_per_chunk = max_steps_per_chunk
        if self._min_var_cut is not None:
            assert self._min_var_thr >= self._min_var_cut or self._min_var_thr == -np.inf, (
                'min_value_to_interp_thresh cannot be less than min_cutoff_value!')
        if (self._min_var_cut is not None) and (self._max_var_cut is not None):
            assert self._min_var_cut < self._max_var_cut, (
                'min_cutoff_value cannot be greater than or equal to max_cutoff_value!')
        if self._max_var_cut is not None:
            assert self._min_var_thr < self._max_var_cut
        assert isinstance(min_vg_val, float), 'min_vg_val must be a float!'
        assert 0 <= min_vg_val < np.inf, 'Invalid value of min_vg_vals!'
        self._n_cpus = n_cpus
        self._plot_figs_flag = plot_figs_flag
        self._min_vg_val = min_vg_val
        self._misc_settings_set_flag = True

    # ------------------------------------------------------------------ toggles
    def turn_ordinary_kriging_on(self):
        self._ork_flag = True

    def turn_ordinary_kriging_off(self):
        assert self._ork_flag
        self._ork_flag = False

    def turn_simple_kriging_on(self):
        self._spk_flag = True

    def turn_simple_kriging_off(self):
        assert self._spk_flag
        self._spk_flag = False

    def turn_external_drift_kriging_on(self, drift_rasters):
        """interp/main.py:284-336.  Elements may be callables ``f(x, y)``, array rasters
        (dicts) or raster paths (ESRI ASCII grids are read here, gisio.read_ascii_grid;
        other formats need GDAL)."""
        assert hasattr(drift_rasters, '__iter__')
        rass = []
        for dr in drift_rasters:
            if callable(dr):
                rass.append(dr)
                continue
            if isinstance(dr, dict):
                # array raster: values [rows, cols] (row 0 = north), corner and cell size,
                # optional no-data value -- what interp/drift.py:25-163 reads from a GeoTIFF
                assert {'values', 'x_min', 'y_max', 'cell_size'} <= set(dr), (
                    'array drift raster needs values, x_min, y_max, cell_size')
                vals = np.ascontiguousarray(dr['values'], dtype=np.float64)
                assert vals.ndim == 2
                rass.append(dict(values=vals, x_min=float(dr['x_min']), y_max=float(dr['y_max']),
                                 cell_size=float(dr['cell_size']), ndv=dr.get('ndv')))
                continue
            assert isinstance(dr, (str, Path)), (
                'Supplied drift raster path is not a string or a pathlib.Path object!')
            dr = Path(dr).absolute()
            assert dr.exists(), 'Supplied drift raster path does not point to a file!'
            assert dr.is_file(), 'Supplied drift raster path does not point to a file!'
            from . import gisio
            ras = gisio.read_raster(dr)                 # ESRI ASCII grid, else needs GDAL
            ras['values'] = np.ascontiguousarray(ras['values'], dtype=np.float64)
            rass.append(ras)
        self._drft_rass = tuple(rass)
        self._n_drft_rass = len(rass)
        assert self._n_drft_rass, 'Zero drift rasters were supplied!'
        self._edk_flag = True

    def turn_external_drift_kriging_off(self):
        assert self._edk_flag
        self._drft_rass = None
        self._n_drft_rass = None
        self._edk_flag = False

    def turn_inverse_distance_weighting_on(self, idw_exps):
        assert hasattr(idw_exps, '__iter__')
        exps = []
        for e in idw_exps:
            assert isinstance(e, (int, float)), 'IDW exponent not a float or an int!'
            exps.append(float(e))
        assert exps, 'Zero IDW exponents given!'
        self._idw_exps = tuple(exps)
        self._n_idw_exps = len(exps)
        self._idw_flag = True

    def turn_inverse_distance_weighting_off(self):
        assert self._idw_flag
        self._idw_exps = None
        self._n_idw_exps = None
        self._idw_flag = False

    def turn_nearest_neighbor_on(self):
        self._nnb_flag = True

    def turn_nearest_neighbor_off(self):
        assert self._nnb_flag
        self._nnb_flag = False

    def turn_ordinary_kriging_est_var_on(self):
        assert self._ork_flag
        self._interp_flag_est_vars = True

    def turn_ordinary_kriging_est_var_off(self):
        assert self._interp_flag_est_vars
        self._interp_flag_est_vars = False

    # ------------------------------------------------------------------ verify
    def _verify_data(self):
        """interp/data.py:747-801."""
        assert self._data_set_flag, 'Call the set_data method first!'
        assert self._out_dir_set_flag, 'Call the set_out_dir method first!'
        assert self._nc_set_flag, 'Call the set_netcdf4_parameters method first!'
        assert self._time_prms_set_flag, 'Call the set_interp_time_parameters method first!'
        assert self._neb_sel_mthd_set_flag, 'Call set_neighbor_selection_method method first!'
        if self._index_type == 'obj' and self._vg_ser_set_flag:
            assert not self._data_df.index.difference(self._vgs_ser.index).size, (
                'For object type index, data and variograms must have the same index entries!')
        self._data_vrfd_flag = True

    def _prepare(self):
        """interp/prepare.py:510-729 without the GIS parts."""
        if not any([self._ork_flag, self._spk_flag, self._edk_flag]):
            self._vg_ser_set_flag = False
            self._vgs_ser = None
        assert any([self._ork_flag, self._spk_flag, self._edk_flag, self._idw_flag,
                    self._nnb_flag])
        if any([self._ork_flag, self._spk_flag, self._edk_flag]):
            assert self._vg_ser_set_flag, 'Kriging needs set_vgs_ser!'

        if self._index_type == 'date':
            assert all(v is not None for v in (self._tbeg, self._tend, self._tfreq)), (
                'beg_time, end_time and time_freq are not set!')
            self._time_rng = pd.date_range(self._tbeg, self._tend, freq=self._tfreq)
        else:
            self._time_rng = self._data_df.index

        # grid: interp/prepare.py:92-242
        arr_ras = [f for f in (self._drft_rass or []) if not callable(f)] if self._edk_flag else []
        if arr_ras and not self._algn_ras_set_flag:
            # the drift rasters decide the cell size (interp/prepare.py:549-550)
            self._cell_size = arr_ras[0]['cell_size']
        if self._algn_ras_set_flag:
            # ... unless an alignment raster is set (interp/prepare.py:553-555, :45-90)
            self._cell_size = self._algn_ras['cell_size']
        assert self._cell_size is not None, 'Cell size unspecified!'
        cs = self._cell_size
        if self._poly_rings is not None:
            # stations near the polygons (interp/bdpolys.py:80-170), grid bounds from the
            # polygons' extent (interp/prepare.py:107-137)
            from . import prep
            keep = prep.points_in_polygons(self._crds_df['X'].values, self._crds_df['Y'].values,
                                           self._poly_rings, self._stn_bdist)
            assert keep.any(), 'Found zero stations that are close enough to the polygons!'
            fin_stns = self._crds_df.index[keep]
            self._crds_df = self._crds_df.loc[fin_stns]
            self._data_df = self._data_df.loc[:, self._data_df.columns.intersection(fin_stns)]
            allv = np.concatenate(self._poly_rings, axis=0)
            x_min, x_max = allv[:, 0].min(), allv[:, 0].max()
            y_min, y_max = allv[:, 1].min(), allv[:, 1].max()
        else:
            x_min, x_max = self._crds_df['X'].min(), self._crds_df['X'].max()
            y_min, y_max = self._crds_df['Y'].min(), self._crds_df['Y'].max()
        if self._algn_ras_set_flag:
            # bounds snapped outwards to the alignment raster's lattice; without polygons
            # the grid is the raster's extent (interp/prepare.py:45-90, misc.py:743-885)
            from . import prep
            a = self._algn_ras
            (x_min, x_max, y_min, y_max), _ = prep.aligned_bounds(
                a['x_min'], a['y_max'], a['cell_size'], a['n_rows'], a['n_cols'],
                (x_min, x_max, y_min, y_max) if self._poly_rings is not None else None,
                self._cell_bdist)
        else:
            x_min -= self._cell_bdist
            x_max += self._cell_bdist
            y_min -= self._cell_bdist
            y_max += self._cell_bdist
        self._x_min, self._x_max, self._y_min, self._y_max = x_min, x_max, y_min, y_max
        if arr_ras:
            # with drift rasters the row / column window is RASTER-relative
            # (interp/prepare.py:150-173): a grid origin that is not aligned to the raster
            # spans one more column / row than ceil((x_max - x_min) / cell).  Raster bounds
            # rounded to 6 decimals like interp/drift.py:134-152
            f0 = arr_ras[0]
            nr0, nc0 = f0['values'].shape
            for f in arr_ras:
                assert np.isclose(f['cell_size'], cs), (
                    f"Drift raster's cell width {f['cell_size']} unequal to the one used {cs}!")
                assert f['values'].shape == (nr0, nc0) and np.isclose(
                    f['x_min'], f0['x_min']) and np.isclose(f['y_max'], f0['y_max']), (
                        'Drift rasters have dissimilar spatial properties!')
                assert (f['ndv'] is None) == (f0['ndv'] is None) and (
                    f['ndv'] is None or np.isclose(f['ndv'], f0['ndv'])), (
                        'Drift rasters have dissimilar spatial properties!')
            dx_min, dx_max, dy_min, dy_max = (float(v) for v in np.round(
                (f0['x_min'], f0['x_min'] + nc0 * cs, f0['y_max'] - nr0 * cs, f0['y_max']), 6))
            self._drft_x_min, self._drft_x_max = dx_min, dx_max
            self._drft_y_min, self._drft_y_max = dy_min, dy_max
            assert x_min >= dx_min, 'Grid x_min outside of the drift rasters!'
            assert x_max <= dx_max, 'Grid x_max outside of drift rasters!'
            assert y_min >= dy_min, 'Grid y_min outside of the drift rasters!'
            assert y_max <= dy_max, 'Grid y_max outside of drift rasters!'
            min_col = int(floor((x_min - dx_min) / cs))
            max_col = int(ceil((x_max - dx_min) / cs)) - 1
            min_row = int(floor((dy_max - y_max) / cs))
            max_row = int(ceil((dy_max - y_min) / cs)) - 1
        else:
            min_col = min_row = 0
            max_col = int(ceil((x_max - x_min) / cs)) - 1
            max_row = int(ceil((y_max - y_min) / cs)) - 1
        assert 0 <= min_col <= max_col, (min_col, max_col)
        assert 0 <= min_row <= max_row, (min_row, max_row)
        self._min_row, self._max_row = min_row, max_row
        self._min_col, self._max_col = min_col, max_col
        n_cols, n_rows = max_col - min_col + 1, max_row - min_row + 1
        xs = np.linspace(x_min + 0.5 * cs, x_min + 0.5 * cs + (n_cols - 1) * cs, n_cols)
        ys = np.linspace(y_max - 0.5 * cs, y_max - 0.5 * cs - (n_rows - 1) * cs, n_rows)
        mx, my = np.meshgrid(xs, ys)
        self._interp_crds_orig_shape = mx.shape
        self._nc_x_crds, self._nc_y_crds = xs, ys
        self._interp_x_crds_msh = mx.ravel()
        self._interp_y_crds_msh = my.ravel()
        full_x, full_y = self._interp_x_crds_msh, self._interp_y_crds_msh

        # cell mask: interp/prepare.py:244-288
        self._cntn_idxs = None
        if self._poly_rings is not None and self._ipoly_flag:
            from . import prep
            m = prep.points_in_polygons(full_x, full_y, self._poly_rings, self._cell_bdist)
            assert m.sum(), 'No cells selected for interpolation!'
            self._interp_x_crds_msh = full_x[m]
            self._interp_y_crds_msh = full_y[m]
            self._cntn_idxs = m
        elif self._cell_mask_src is not None:
            m = self._cell_mask_src
            m = np.asarray(m(full_x, full_y) if callable(m) else m, dtype=bool).ravel()
            assert m.shape == full_x.shape, 'cell mask does not match the grid!'
            assert m.sum(), 'No cells selected for interpolation!'
            self._interp_x_crds_msh = full_x[m]
            self._interp_y_crds_msh = full_y[m]
            self._cntn_idxs = m

        # neighbours: interp/prepare.py:434-463
        if self._neb_sel_mthd in ('nrst', 'pie') and self._n_nebs >= self._crds_df.shape[0]:
            self._neb_sel_mthd = 'all'

        # drift: interp/drift.py:25-226 reduced to sampling callables
        if self._edk_flag:
            sx, sy = self._crds_df['X'].values, self._crds_df['Y'].values
            cell_rows, stn_cols = [], []
            for f in self._drft_rass:
                if callable(f):
                    cell_rows.append(np.asarray(
                        f(self._interp_x_crds_msh, self._interp_y_crds_msh), dtype=np.float64))
                    stn_cols.append(np.asarray(f(sx, sy), dtype=np.float64))
                    continue
                # array raster sampled on the GPU: interp/drift.py:165-226 (cells through
                # the row / column window of interp/prepare.py:150-172, stations through
                # int((x - x_min) / cell), int((y_max - y) / cell))
                from . import prep
                ndv = arr_ras[0]['ndv']         # interp/drift.py:154: the first raster's
                rr, cc = prep.drift_cell_indices(min_row, max_row, min_col, max_col,
                                                 self._cntn_idxs)
                cell_rows.append(prep.sample_raster(f['values'], rr, cc, ndv))
                rr, cc = prep.drift_point_indices(sx, sy, self._drft_x_min, self._drft_y_max, cs)
                stn_cols.append(prep.sample_raster(f['values'], rr, cc, ndv))
            self._drft_arrs = np.vstack(cell_rows)
            self._stns_drft_df = pd.DataFrame(np.column_stack(stn_cols), index=self._crds_df.index)
            assert np.all(np.isfinite(self._stns_drft_df.values)), (   # interp/drift.py:221-222
                'Invalid value(s) of drift(s) for stations in drift rasters!')

        self._out_dir.mkdir(exist_ok=True)

        # interp_args: interp/prepare.py:638-695 (order OK, SK, EDK, IDW.., NNB, EST_VARS_OK)
        self._interp_args = []
        if self._ork_flag:
            self._interp_args.append(('OK', None, 'OK'))
        if self._spk_flag:
            self._interp_args.append(('SK', None, 'SK'))
        if self._edk_flag:
            self._interp_args.append(('EDK', None, 'EDK'))
        if self._idw_flag:
            for i, e in enumerate(self._idw_exps):
                self._interp_args.append(('IDW', None, f'IDW_{i:03d}', e))
        if self._nnb_flag:
            self._interp_args.append(('NNB', None, 'NNB'))
        if self._interp_flag_est_vars:
            self._interp_args.append(('EST_VARS_OK', None, 'EST_VARS_OK'))

        # align stations and time: interp/prepare.py:701-720
        all_stns = self._data_df.columns.intersection(self._crds_df.index)
        if self._edk_flag:
            all_stns = all_stns.intersection(self._stns_drft_df.index)
            self._stns_drft_df = self._stns_drft_df.loc[all_stns]
        assert all_stns.shape[0] > 1, 'Less than 2 common stations!'
        self._data_df = self._data_df.loc[:, all_stns].reindex(self._time_rng)
        self._crds_df = self._crds_df.loc[all_stns]
        if self._vg_ser_set_flag:
            self._vgs_ser = self._vgs_ser.reindex(self._time_rng).astype(str)
            # variogram clustering (interp/prepare.py:465-508) only reorders the time
            # axis for the CPU scheduler; the engine groups by variogram itself
            self._vgs_rord_tidxs_ser = pd.Series(
                np.arange(self._time_rng.shape[0]), index=self._time_rng)
        else:
            self._vgs_rord_tidxs_ser = None

        import torch.distributed as tdist
        if (not tdist.is_initialized()) or tdist.get_rank() == 0:
            self._initiate_nc()
        else:
            self._nc_file_path = self._out_dir / (self._nc_out.split('.', 1)[0] + '.nc')
        if tdist.is_initialized():
            tdist.barrier()
        self._prpd_flag = True

    def _initiate_nc(self):
        """interp/prepare.py:290-432."""
        if self._index_type == 'date':
            tvals = ncwriter.time_numbers(self._time_rng, self._nc_tunits, self._nc_tcldr,
                                          self._tfreq)
            tunits, tcal = self._nc_tunits, self._nc_tcldr
        else:
            tvals = np.arange(self._time_rng.shape[0], dtype=np.int64)
            tunits = tcal = None
        sett = {
            'sett_index_type': self._index_type,
            'sett_stns_min_dist_thrsh': self._stns_min_dist_thrsh,
            'sett_drft_rass': self._drft_rass, 'sett_idw_exps': self._idw_exps,
            'sett_ork_flag': self._ork_flag, 'sett_spk_flag': self._spk_flag,
            'sett_edk_flag': self._edk_flag, 'sett_idw_flag': self._idw_flag,
            'sett_nnb_flag': self._nnb_flag,
            'sett_interp_flag_est_vars': self._interp_flag_est_vars,
            'sett_out_dir': self._out_dir, 'sett_cell_size': self._cell_size,
            'sett_tbeg': self._tbeg, 'sett_tend': self._tend, 'sett_tfreq': self._tfreq,
            'sett_algn_ras': self._algn_ras, 'sett_poly_shp': self._poly_shp,
            'sett_ipoly_flag': self._ipoly_flag, 'sett_stn_bdist': self._stn_bdist,
            'sett_cell_bdist': self._cell_bdist,
            'sett_poly_simplify_tol_ratio': self._poly_simplify_tol_ratio,
            'sett_min_var_thr': self._min_var_thr, 'sett_min_var_cut': self._min_var_cut,
            'sett_max_var_cut': self._max_var_cut,
            'sett_max_steps_per_chunk': self._max_steps_per_chunk,
            'sett_min_vg_val': self._min_vg_val, 'sett_neb_sel_mthd': self._neb_sel_mthd,
            'sett_n_nebs': self._n_nebs, 'sett_n_pies': self._n_pies}
        self._nc_file_path = ncwriter.create(
            self._out_dir / (self._nc_out.split('.', 1)[0] + '.nc'), self._nc_x_crds,
            self._nc_y_crds, tvals, self._interp_args, self._intrp_dtype, self._nc_vunits,
            self._nc_vlab, tunits, tcal, self._nc_cprm_levl, sett,
            self._nc_xlab, self._nc_ylab, self._nc_tlab)

    def verify(self):
        """interp/main.py:63-72."""
        self._verify_data()
        self._prepare()
        assert self._prpd_flag, 'Preparing data for interpolation failed!'
        self._main_vrfd_flag = True

    # ------------------------------------------------------------------ interpolate
    def _grid_row_chunks(self, world, n_steps):
        """Grid-row chunks, the reference's second task axis (interp/main.py:733-734,
        :805-811).  One chunk unless (a) there are fewer time steps than 4 x ranks -- few
        steps on a huge grid shard by rows instead of by time -- or (b) a single step's
        fields do not fit an eighth of the free device memory."""
        import torch
        ny = int(self._interp_crds_orig_shape[0])
        chunks = 1
        if world > 1 and n_steps < 4 * world:
            chunks = min(world, ny)
        fld = int(np.prod(self._interp_crds_orig_shape)) * np.dtype(self._intrp_dtype).itemsize
        per_step = fld * max(1, len(self._interp_args))
        free, _ = torch.cuda.mem_get_info()
        while chunks < ny and per_step / chunks > 0.125 * free:
            chunks += 1
        forced = getattr(self, '_grid_row_chunks_forced', None)   # test hook
        if forced is not None:
            chunks = int(forced)
        return np.unique(np.linspace(0, ny, chunks + 1, dtype=np.int64))

    def _time_chunks(self, n_row_chunks=1):
        """Time chunks sized for HBM instead of the reference's RAM model
        (interp/main.py:652-859): the fields of one task may take a quarter of the
        free device memory (two tasks are in flight)."""
        import torch
        n_steps = self._data_df.shape[0]
        fld = int(np.prod(self._interp_crds_orig_shape)) * np.dtype(self._intrp_dtype).itemsize
        free, _ = torch.cuda.mem_get_info()
        per_step = fld * max(1, len(self._interp_args)) / max(1, int(n_row_chunks))
        max_steps = max(1, int(0.25 * free // per_step))
        if self._max_steps_per_chunk is not None:
            max_steps = min(max_steps, self._max_steps_per_chunk)
        n_chunks = int(ceil(n_steps / max_steps))
        return np.unique(np.linspace(0, n_steps, n_chunks + 1, dtype=np.int64))

    def _chunk_args(self, beg, end, n_chunks, lock, row_beg=0, row_end=None):
        """interp/main.py:604-650."""
        data_df = self._data_df.iloc[beg:end]
        krg = any(a[0] in ('OK', 'SK', 'EDK') for a in self._interp_args)
        vgs_ser = self._vgs_ser.loc[data_df.index] if krg else None
        rord = self._vgs_rord_tidxs_ser.loc[data_df.index] if krg else None
        if krg:
            assert np.all(vgs_ser.values != 'nan'), (
                'NaN VGs not allowed! Use Nugget or any other appropriate one!')
        edk = self._edk_flag
        if row_end is None:
            row_end = int(self._interp_crds_orig_shape[0])
        return (data_df, int(beg), int(end), n_chunks, self._interp_args, lock,
                self._drft_arrs if edk else None, self._stns_drft_df if edk else None,
                vgs_ser, rord, int(row_beg), int(row_end))

    def interpolate(self):
        """interp/main.py:74-236.  The job is a list of (time chunk x grid-row chunk)
        tasks (dist.plan_tasks).  One rank: task i+1 is prepared and queued on the GPU
        while the fields of task i are downloaded, reduced to statistics and written.
        Several ranks (torch.distributed initialised): every rank interpolates its own
        tasks; round by round the finished, rounded slabs travel over NCCL (NVLink) to
        the writer rank, which holds at most two foreign slabs at a time, downloads and
        writes them -- only that rank touches the file (HDF5 is single-writer).  Memory
        on every GPU is bounded by the task size, not by the job size."""
        assert self._main_vrfd_flag, 'Call the verify method first!'
        import torch
        import torch.distributed as tdist
        from . import dist as sdist

        t0 = timeit.default_timer()
        lock = Lock()
        n_steps = self._data_df.shape[0]
        multi = tdist.is_initialized() and tdist.get_world_size() > 1
        rank = tdist.get_rank() if multi else 0
        world = tdist.get_world_size() if multi else 1
        writer = 0
        row_b = self._grid_row_chunks(world, n_steps)
        time_b = self._time_chunks(row_b.size - 1)
        if multi and row_b.size == 2 and time_b.size - 1 < world:
            # fewer time chunks than ranks: one (or more) per rank
            time_b = np.unique(np.linspace(0, n_steps, min(n_steps, world) + 1, dtype=np.int64))
        labels = [a[2] for a in self._interp_args]
        tasks = sdist.plan_tasks(time_b, row_b, world)
        sg = sdist.StreamedGather(tasks, labels, writer=writer)
        mine = sg.my_tasks()
        n_cols = int(self._interp_crds_orig_shape[1])
        steps_cls = SpInterpSteps(self)
        stats_acc = {}
        dev = torch.device('cuda', torch.cuda.current_device())
        tdt = torch.float32 if np.dtype(self._intrp_dtype) == np.float32 else torch.float64
        eng = steps_cls._get_engine()
        self.gather_stats = dict(tasks=len(tasks), mine=len(mine), rounds=sg.n_rounds)

        def consume(task, lab, buf, st):
            """Writer: one foreign slab (still on the GPU) -> host -> statistics + file."""
            tb, te, rb, re = task[:4]
            flds = InterpFields()
            flds.rounded = True
            flds.stats = {lab: st.cpu().numpy()}
            dl = eng._packed_downloader(buf, int(self._nc_nmrl_prcn))
            flds[lab] = (dl.download(buf, int(self._nc_nmrl_prcn)) if dl is not None
                         else buf.cpu().numpy())
            args = self._chunk_args(tb, te, 1, lock, rb, re)
            out = (lock, tb, te, args[0], args[8], 1, [lab], flds, rb, re, args[9],
                   args[0].index, timeit.default_timer())
            self._collect_stats(out, stats_acc)
            steps_cls._write_to_disk(out)

        prev = None
        for j in range(sg.n_rounds + 1):
            cur = None
            if j < len(mine):
                tb, te, rb, re = mine[j][:4]
                args = self._chunk_args(tb, te, 1, lock, rb, re)
                cur = (steps_cls._submit_interp(args, output_stage=True), args,
                       timeit.default_timer())
            if prev is not None:
                if rank == writer:
                    out = steps_cls._finish_interp(*prev)
                    self._collect_stats(out, stats_acc)
                    steps_cls._write_to_disk(out)
                else:
                    out = steps_cls._finish_interp(*prev, to_host=False)
                    flds = out[7]
                    sg.send({lab: flds[lab] for lab in labels},
                            {lab: torch.from_numpy(np.ascontiguousarray(flds.stats[lab])).to(dev)
                             for lab in labels})
            if multi and rank == writer and j >= 1:
                sg.receive_round(j - 1, n_cols, tdt, dev, consume)
            prev = cur
        sg.flush()
        self.gather_stats['bytes_received'] = int(sg.bytes_received)
        if multi:
            tdist.barrier()
        if rank == writer:
            from . import ncwriter
            ncwriter.finalize(self._nc_file_path)      # chunk index written, file closed
            self._save_stats_sers(self._finish_stats(stats_acc))
        if self._vb:
            print(f'Done with the interpolation in {timeit.default_timer() - t0:0.1f} seconds.')

    # ------------------------------------------------------------------ stats
    def _collect_stats(self, out, stats_acc):
        """Per-step min / mean / max / std / count of every field
        (interp/main.py:474-525 computes them by re-reading the netCDF; here they
        are taken from the rounded field before it is written).  A step may arrive in
        several grid-row chunks: the parts are merged (Chan et al. for mean / M2)."""
        labels, flds, time_steps = out[6], out[7], out[11]
        dev_stats = getattr(flds, 'stats', None)
        for lab in labels:
            if lab == 'EST_VARS_OK':
                continue
            if dev_stats is not None and lab in dev_stats:
                st = np.asarray(dev_stats[lab], dtype=np.float64)   # spx_round_stats_dev
                mn, mean, mx, std, cnt = st[0], st[1], st[2], st[3], st[4]
            else:
                f = np.round(flds[lab], self._nc_nmrl_prcn)
                with np.errstate(invalid='ignore'), np.testing.suppress_warnings() as sup:
                    sup.filter(RuntimeWarning)
                    mn, mean, mx = (np.nanmin(f, axis=1), np.nanmean(f, axis=1),
                                    np.nanmax(f, axis=1))
                    std = np.nanstd(f, axis=1)
                    cnt = np.isfinite(f).sum(axis=1).astype(np.float64)
            acc = stats_acc.setdefault(lab, {})
            for k, t in enumerate(time_steps):
                n_b = float(cnt[k])
                part = (n_b, float(mean[k]), float(std[k]) ** 2 * n_b, float(mn[k]), float(mx[k]))
                a = acc.get(t)
                if a is None or a[0] == 0.0:
                    acc[t] = part if (a is None or n_b > 0.0) else a
                elif n_b > 0.0:
                    n = a[0] + n_b
                    dlt = part[1] - a[1]
                    acc[t] = (n, a[1] + dlt * (n_b / n), a[2] + part[2] + dlt * dlt * (a[0] * n_b / n),
                              min(a[3], part[3]), max(a[4], part[4]))

    @staticmethod
    def _finish_stats(stats_acc):
        stats_rows = {}
        for lab, acc in stats_acc.items():
            for t, (n, mean, m2, mn, mx) in acc.items():
                vals = dict(min=mn, mean=mean, max=mx,
                            std=np.sqrt(m2 / n) if n > 0 else np.nan, count=n)
                if n == 0:
                    vals.update(min=np.nan, mean=np.nan, max=np.nan)
                for stat, v in vals.items():
                    stats_rows.setdefault(f'{lab}_{stat}', {})[t] = np.float32(v)
        return stats_rows

    def _save_stats_sers(self, stats_rows):
        data_df = self._data_df.sort_index()
        stats = ['min', 'mean', 'max', 'std', 'count']
        df = pd.DataFrame(index=data_df.index, dtype=np.float32)
        for stat in stats:
            df[f'data_{stat}'] = getattr(data_df, stat)(axis=1).astype(np.float32)
        for col, vals in stats_rows.items():
            df[col] = pd.Series(vals, dtype=np.float32)
        df.to_csv(self._out_dir / 'stats.csv', sep=';', float_format='%0.6f')
