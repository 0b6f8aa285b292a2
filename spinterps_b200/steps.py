"""Operator seam: ``SpInterpSteps`` with the reference's call signature
(interp/steps.py:29-71, compute half :478-877) on top of the GPU engine.

``SpInterpSteps(main).interpolate_subset(args)`` takes the same 12-tuple the
reference's scheduler builds (interp/main.py:604-650): pandas frames for the
data / variograms / station drifts, chunk offsets, ``interp_args`` and the
grid-row range.  ``_get_all_interp_outputs`` returns the same 13-tuple whose
element 7 is ``{label: ndarray[T_chunk, rows*cols] of _intrp_dtype}``.
"""
from __future__ import annotations

import timeit

import numpy as np

from .engine import ChunkEngine

_READ_LABS = (  # interp/steps.py:33-53
    '_vb', '_n_cpus', '_mp_flag', '_crds_df', '_min_var_thr', '_min_var_cut',
    '_max_var_cut', '_cntn_idxs', '_interp_crds_orig_shape', '_interp_x_crds_msh',
    '_interp_y_crds_msh', '_nc_file_path', '_nc_nmrl_prcn', '_neb_sel_mthd', '_n_nebs',
    '_n_pies', '_min_vg_val', '_interp_flag_est_vars', '_intrp_dtype')


class InterpFields(dict):
    """``{label: field}`` (element 7 of the reference's 13-tuple) plus what the GPU
    output stage already did: ``rounded`` (fields are rounded to ``nmrl_prcn`` decimals)
    and ``stats`` ({label: ndarray[5, T]} per-step min / mean / max / std / count)."""
    rounded = False
    stats = None


class SpInterpSteps:

    def __init__(self, spinterp_main_cls, engine=None):
        for lab in _READ_LABS:
            setattr(self, lab, getattr(spinterp_main_cls, lab))
        self._engine = engine

    def _get_engine(self):
        if self._engine is None:
            self._engine = ChunkEngine()
        return self._engine

    def interpolate_subset(self, args_for_interp):
        """interp/steps.py:61-71.  Unlike the reference's traceback_wrapper
        (misc.py:95-117) errors are raised, not printed and swallowed."""
        pend = self._submit_interp(args_for_interp, output_stage=True)
        self._write_to_disk(self._finish_interp(pend, args_for_interp))
        return

    # -- compute half ------------------------------------------------------
    def _chunk_kwargs(self, args):
        (data_df, beg_idx, end_idx, max_rng, interp_args, lock, drft_arrs, stns_drft_df,
         vgs_ser, vgs_rord_tidxs_ser, fld_beg_row, fld_end_row) = args

        interp_types = [a[0] for a in interp_args]
        krg_flag = any(t in interp_types for t in ('OK', 'SK', 'EDK'))
        if krg_flag:
            assert np.all(vgs_ser != 'nan'), (   # steps.py:504-507
                'NaN VGs not allowed! Use Nugget or any other appropriate one!')
        assert np.all(data_df.columns == self._crds_df.index)   # steps.py:571
        assert np.unique(data_df.columns.values).size == data_df.shape[1]
        if vgs_ser is not None:
            assert not data_df.index.difference(vgs_ser.index).shape[0], (
                'Data and variogram series have non-intersecting indices!')
            vgs = [str(v) for v in vgs_ser.loc[data_df.index].values]
        else:
            vgs = None

        return dict(
            data=np.ascontiguousarray(data_df.values, dtype=np.float64),
            stn_xs=np.ascontiguousarray(self._crds_df.loc[:, 'X'].values, dtype=np.float64),
            stn_ys=np.ascontiguousarray(self._crds_df.loc[:, 'Y'].values, dtype=np.float64),
            cell_xs=self._interp_x_crds_msh, cell_ys=self._interp_y_crds_msh,
            grid_shape=tuple(self._interp_crds_orig_shape), interp_args=list(interp_args),
            vgs=vgs, cntn_idxs=self._cntn_idxs, drft_arrs=drft_arrs,
            stns_drft=None if stns_drft_df is None else np.ascontiguousarray(
                stns_drft_df.loc[self._crds_df.index].values, dtype=np.float64),
            fld_beg_row=int(fld_beg_row), fld_end_row=int(fld_end_row),
            neb_sel_mthd=self._neb_sel_mthd, n_nebs=self._n_nebs, n_pies=self._n_pies,
            min_var_thr=self._min_var_thr, min_var_cut=self._min_var_cut,
            max_var_cut=self._max_var_cut, min_vg_val=self._min_vg_val,
            est_var_flag=bool(self._interp_flag_est_vars), intrp_dtype=self._intrp_dtype)

    def _submit_interp(self, args, output_stage=False):
        """Queue the chunk on the GPU (engine.submit_chunk); returns the pending chunk.
        output_stage: also run the rounding of interp/steps.py:907-912 and the per-step
        statistics of interp/main.py:474-525 on the device (the fields then come back
        rounded, which ``_get_all_interp_outputs`` of the reference does not do)."""
        return self._get_engine().submit_chunk(
            round_decimals=int(self._nc_nmrl_prcn) if output_stage else None,
            field_stats=bool(output_stage), **self._chunk_kwargs(args))

    def _finish_interp(self, pend, args, interp_beg_time=None, to_host=True):
        (data_df, beg_idx, end_idx, max_rng, interp_args, lock, drft_arrs, stns_drft_df,
         vgs_ser, vgs_rord_tidxs_ser, fld_beg_row, fld_end_row) = args
        if interp_beg_time is None:
            interp_beg_time = timeit.default_timer()
        interp_labels = [a[2] for a in interp_args]
        # to the host in the 2-byte transport form when the output stage rounded the fields:
        # the writer decodes step by step (transfer.PackedField)
        raw, prblm = pend.result(to_host=('packed' if (to_host and pend.round_decimals is not None)
                                          else to_host))
        flds = InterpFields(raw)
        flds.rounded = pend.round_decimals is not None
        flds.stats = pend.field_stats

        time_steps = data_df.index
        if prblm and self._vb:   # steps.py:847-860
            with lock:
                print('WARNING: There were problems while interpolating at the following steps:')
                print([time_steps[i] for i in prblm])

        return (lock, beg_idx, end_idx, data_df, vgs_ser, max_rng, interp_labels, flds,
                fld_beg_row, fld_end_row, vgs_rord_tidxs_ser, time_steps, interp_beg_time)

    def _get_all_interp_outputs(self, args):
        interp_beg_time = timeit.default_timer()
        return self._finish_interp(self._submit_interp(args), args, interp_beg_time)

    # -- output half -------------------------------------------------------
    def _write_to_disk(self, args):
        """interp/steps.py:879-969: round to ``nmrl_prcn`` decimals in the field
        dtype and write under the lock -- consecutive slabs when there are no
        variograms, one step at a time to ``vgs_rord_tidxs_ser`` otherwise."""
        from .ncwriter import open_for_update

        (lock, beg_idx, end_idx, data_df, vgs_ser, max_rng, interp_labels, interp_flds_dict,
         fld_beg_row, fld_end_row, vgs_rord_tidxs_ser, time_steps, interp_beg_time) = args

        with lock:
            nc_hdl = open_for_update(self._nc_file_path)
            try:
                for label in interp_labels:
                    flds = interp_flds_dict[label]
                    if hasattr(flds, 'row'):              # transfer.PackedField (rounded)
                        ny_all = self._interp_crds_orig_shape[0]
                        if (hasattr(nc_hdl, 'write_packed') and fld_beg_row == 0
                                and fld_end_row == ny_all):
                            if vgs_ser is None:
                                t_index = np.arange(beg_idx, end_idx)
                            else:
                                t_index = [int(vgs_rord_tidxs_ser.loc[t]) for t in time_steps]
                            nc_hdl.write_packed(label, t_index, flds)
                            flds.release()
                            interp_flds_dict[label] = None
                            nc_hdl.sync()
                            continue
                        dec = flds.decode()
                        flds.release()
                        flds = dec
                    if (np.issubdtype(flds.dtype, np.floating)
                            and not getattr(interp_flds_dict, 'rounded', False)):
                        np.round(flds, self._nc_nmrl_prcn, flds)
                    nrows = fld_end_row - fld_beg_row
                    flds3 = flds.reshape(flds.shape[0], nrows, -1)
                    if vgs_ser is None:
                        nc_is = np.linspace(beg_idx, end_idx, max_rng + 1, dtype=int)
                        ar_is = nc_is - beg_idx
                        for i in range(max_rng):
                            nc_hdl.write(label, slice(nc_is[i], nc_is[i + 1]),
                                         fld_beg_row, fld_end_row, flds3[ar_is[i]:ar_is[i + 1]])
                    else:
                        for i in range(len(time_steps)):
                            nc_idx = int(vgs_rord_tidxs_ser.loc[time_steps[i]])
                            nc_hdl.write(label, nc_idx, fld_beg_row, fld_end_row, flds3[i])
                    interp_flds_dict[label] = None
                    nc_hdl.sync()
            finally:
                nc_hdl.close()
        return
