"""ctypes binding of libspx_b200.so (the C-ABI declared in include/spx_b200.h).

The product path has NO CPU fallback: if the library is missing, or a compute
entry point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / 'lib' / 'libspx_b200.so'

SPX_VG_MAX_TERMS = 8
SPX_BM = 256
VG_NAMES = ('Rng', 'Nug', 'Sph', 'Exp', 'Lin', 'Gau', 'Pow', 'Hol')
KRG_KINDS = {'OK': 0, 'SK': 1, 'EDK': 2}
GEN_VG, GEN_IDW = 0, 1
EPI_FIELD, EPI_AUX, EPI_FIELD_DIV, EPI_QUADFORM = 0, 1, 2, 3

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class SpxError(RuntimeError):
    pass


class spx_vg(C.Structure):
    _fields_ = [('n_terms', C.c_int32),
                ('types', C.c_int32 * SPX_VG_MAX_TERMS),
                ('sills', C.c_double * SPX_VG_MAX_TERMS),
                ('ranges', C.c_double * SPX_VG_MAX_TERMS)]


# numpy mirror of spx_vg for device upload (same layout as the C struct)
VG_DTYPE = np.dtype([('n_terms', np.int32), ('types', np.int32, SPX_VG_MAX_TERMS),
                     ('sills', np.float64, SPX_VG_MAX_TERMS),
                     ('ranges', np.float64, SPX_VG_MAX_TERMS)], align=True)
assert VG_DTYPE.itemsize == C.sizeof(spx_vg), (VG_DTYPE.itemsize, C.sizeof(spx_vg))


class spx_systems(C.Structure):
    _fields_ = [('n_sys', C.c_int32), ('n_drifts', C.c_int32),
                ('sys_n', C.c_void_p), ('sys_kind', C.c_void_p), ('sys_vg', C.c_void_p),
                ('sys_stn_off', C.c_void_p), ('sys_w_off', C.c_void_p),
                ('sys_piv_off', C.c_void_p), ('stn_list', C.c_void_p),
                ('stn_x', C.c_void_p), ('stn_y', C.c_void_p), ('stn_drift', C.c_void_p),
                ('work', C.c_void_p), ('piv', C.c_void_p), ('info', C.c_void_p),
                ('max_m', C.c_int32)]


class spx_rhs(C.Structure):
    _fields_ = [('n_rhs', C.c_int32),
                ('rhs_sys', C.c_void_p), ('rhs_kind', C.c_void_p), ('rhs_arg', C.c_void_p),
                ('rhs_row', C.c_void_p), ('data', C.c_void_p),
                ('n_stn', C.c_int32), ('kpad', C.c_int32),
                ('coef', C.c_void_p), ('resid', C.c_void_p),
                ('dense', C.c_void_p), ('dense_ld', C.c_int64), ('coef_row_major', C.c_int32)]


class spx_downdate(C.Structure):
    _fields_ = [('n_sys', C.c_int32), ('n_stn', C.c_int32), ('n_border', C.c_int32),
                ('max_r', C.c_int32), ('ginv', C.c_void_p),
                ('sys_r', C.c_void_p), ('sys_miss_off', C.c_void_p), ('miss_list', C.c_void_p),
                ('sys_n', C.c_void_p), ('sys_stn_off', C.c_void_p), ('stn_list', C.c_void_p),
                ('sys_rhs_off', C.c_void_p), ('sys_rhs_cnt', C.c_void_p),
                ('rhs_urow', C.c_void_p), ('rhs_row', C.c_void_p), ('rhs_kind', C.c_void_p),
                ('ut', C.c_void_p), ('kpad', C.c_int32), ('coef', C.c_void_p),
                ('resid', C.c_void_p), ('info', C.c_void_p), ('coef_row_major', C.c_int32),
                ('sys_order', C.c_void_p),
                ('coef_t', C.c_void_p), ('coef_t_ld', C.c_int64), ('base', C.c_void_p),
                ('base_f', C.c_double)]


class spx_dd_plan(C.Structure):
    _fields_ = [('n_sys', C.c_int32), ('n_data', C.c_int32), ('n_rhs', C.c_int32),
                ('max_r', C.c_int32), ('total_r', C.c_int64), ('total_n', C.c_int64),
                ('off_sys_r', C.c_int64), ('off_sys_miss_off', C.c_int64),
                ('off_miss_list', C.c_int64), ('off_sys_n', C.c_int64),
                ('off_sys_stn_off', C.c_int64), ('off_stn_list', C.c_int64),
                ('off_sys_rhs_off', C.c_int64), ('off_sys_rhs_cnt', C.c_int64),
                ('off_rhs_urow', C.c_int64), ('off_rhs_row', C.c_int64),
                ('off_rhs_kind', C.c_int64), ('off_sys_order', C.c_int64),
                ('off_bt_step', C.c_int64), ('off_sys_grp', C.c_int64),
                ('off_pos_ones', C.c_int64), ('n_upload_bytes', C.c_int64),
                ('n_bytes', C.c_int64)]


class spx_multivg(C.Structure):
    _fields_ = [('coef', C.c_void_p), ('n_rows', C.c_int64),
                ('kpad', C.c_int32), ('n_stn', C.c_int32), ('n_border', C.c_int32),
                ('stn_x', C.c_void_p), ('stn_y', C.c_void_p),
                ('cell_x', C.c_void_p), ('cell_y', C.c_void_p), ('n_cells', C.c_int64),
                ('cell_drift', C.c_void_p), ('vgs', C.c_void_p), ('row_vg', C.c_void_p),
                ('covar_flag', C.c_int32), ('min_vg_val', C.c_double),
                ('row_dst', C.c_void_p), ('out', C.c_void_p), ('out_ld', C.c_int64),
                ('out_f64', C.c_int32), ('cell_pos', C.c_void_p),
                ('has_lo', C.c_int32), ('has_hi', C.c_int32),
                ('lo', C.c_double), ('hi', C.c_double), ('all_fast', C.c_int32)]


class spx_local(C.Structure):
    _fields_ = [('stn_x', C.c_void_p), ('stn_y', C.c_void_p),
                ('bin_start', C.c_void_p), ('bin_stn', C.c_void_p),
                ('x0', C.c_double), ('y0', C.c_double), ('inv_bin', C.c_double),
                ('nbx', C.c_int32), ('nby', C.c_int32), ('R', C.c_double), ('F', C.c_double),
                ('cell_x', C.c_void_p), ('cell_y', C.c_void_p), ('n_cells', C.c_int64),
                ('cap', C.c_int32), ('cnt', C.c_void_p), ('idx', C.c_void_p), ('val', C.c_void_p),
                ('vg', spx_vg), ('covar_flag', C.c_int32), ('min_vg_val', C.c_double),
                ('coef', C.c_void_p), ('base', C.c_void_p), ('n_rows', C.c_int64),
                ('kpad', C.c_int32), ('n_stn', C.c_int32), ('n_drifts', C.c_int32),
                ('cell_drift', C.c_void_p), ('row_dst', C.c_void_p), ('out', C.c_void_p),
                ('out_ld', C.c_int64), ('out_f64', C.c_int32), ('cell_pos', C.c_void_p),
                ('has_lo', C.c_int32), ('has_hi', C.c_int32),
                ('lo', C.c_double), ('hi', C.c_double), ('rows_all_valid', C.c_int32),
                ('coef_t', C.c_void_p), ('coef_t_ld', C.c_int64),
                ('tile_cnt', C.c_void_p), ('tile_stn', C.c_void_p), ('slot', C.c_void_p)]


SPX_LOCAL_TILE = 256
SPX_LOCAL_TILE_CAP = 32


class spx_nrst(C.Structure):
    _fields_ = [('n_grp', C.c_int32), ('n_cells', C.c_int64),
                ('k', C.c_int32), ('n_border', C.c_int32), ('n_drifts', C.c_int32),
                ('kind', C.c_int32), ('n_stn', C.c_int32),
                ('nbu', C.c_void_p), ('cell_grp', C.c_void_p),
                ('stn_x', C.c_void_p), ('stn_y', C.c_void_p), ('stn_drift', C.c_void_p),
                ('cell_x', C.c_void_p), ('cell_y', C.c_void_p), ('cell_drift', C.c_void_p),
                ('vg', spx_vg), ('min_vg_val', C.c_double),
                ('data', C.c_void_p), ('steps', C.c_void_p), ('n_t', C.c_int32),
                ('min_var_thr', C.c_double), ('step_bypass', C.c_void_p),
                ('coef', C.c_void_p), ('ovr', C.c_void_p), ('info', C.c_void_p),
                ('cell_pos', C.c_void_p), ('out', C.c_void_p), ('out_ld', C.c_int64),
                ('out_f64', C.c_int32), ('has_lo', C.c_int32), ('has_hi', C.c_int32),
                ('lo', C.c_double), ('hi', C.c_double), ('idw_exp', C.c_double),
                ('inv', C.c_void_p), ('ev_out', C.c_void_p), ('u_beg', C.c_int32),
                ('u_end', C.c_int32)]


class spx_gemm(C.Structure):
    _fields_ = [('coef', C.c_void_p), ('n_rows', C.c_int64),
                ('kpad', C.c_int32), ('n_stn', C.c_int32), ('n_border', C.c_int32),
                ('stn_x', C.c_void_p), ('stn_y', C.c_void_p),
                ('cell_x', C.c_void_p), ('cell_y', C.c_void_p), ('n_cells', C.c_int64),
                ('cell_drift', C.c_void_p),
                ('gen', C.c_int32), ('covar_flag', C.c_int32),
                ('vg', spx_vg),
                ('min_vg_val', C.c_double), ('idw_exp', C.c_double), ('dist_scale', C.c_double),
                ('epi', C.c_int32),
                ('row_dst', C.c_void_p), ('row_aux', C.c_void_p),
                ('out', C.c_void_p), ('out_ld', C.c_int64), ('out_f64', C.c_int32),
                ('cell_pos', C.c_void_p), ('aux', C.c_void_p),
                ('has_lo', C.c_int32), ('has_hi', C.c_int32),
                ('lo', C.c_double), ('hi', C.c_double), ('quad_slot', C.c_int32)]


class spx_pack_row(C.Structure):
    _fields_ = [('mode', C.c_int32), ('qmin', C.c_int32), ('qmax', C.c_int32),
                ('n_nan', C.c_int32)]


PACK_ROW_DTYPE = np.dtype([('mode', np.int32), ('qmin', np.int32), ('qmax', np.int32),
                           ('n_nan', np.int32)])
SPX_PACK_U16, SPX_PACK_RAW = 0, 2
SPX_DPACK_ROUND, SPX_DPACK_WRITE_BACK = 1, 2


class spx_sparse_cov(C.Structure):
    _fields_ = [('n_comp', C.c_int32), ('max_size', C.c_int32), ('n_single', C.c_int32),
                ('reserved', C.c_int32), ('comp_off', C.c_void_p),
                ('comp_stn', C.c_void_p), ('blk_off', C.c_void_p), ('blk', C.c_void_p)]


SPX_SPARSE_MAX_COMP = 8


class spx_fast_cfg(C.Structure):
    _fields_ = [('n_stn', C.c_int32), ('n_border', C.c_int32), ('kpad', C.c_int32),
                ('max_steps', C.c_int32), ('n_slots', C.c_int32), ('min_systems', C.c_int32),
                ('min_var_thr', C.c_double), ('ginv', C.c_void_p),
                ('lambda_bound', C.c_double), ('lambda_tol', C.c_double),
                ('estimator', C.c_int32), ('want_coef_t', C.c_int32), ('base_f', C.c_double),
                ('local', spx_local), ('gemm', spx_gemm), ('profile', C.c_int32),
                ('solve_stream', C.c_int32), ('sparse', spx_sparse_cov)]


class spx_fast_result(C.Structure):
    _fields_ = [('status', C.c_int32), ('slot', C.c_int32), ('n_grps', C.c_int32),
                ('n_krige', C.c_int32), ('n_sys', C.c_int32), ('max_r', C.c_int32),
                ('n_none', C.c_int32), ('n_single', C.c_int32), ('n_mean', C.c_int32),
                ('launches', C.c_int32), ('h2d_bytes', C.c_int64), ('d_data', C.c_void_p),
                ('host_ms', C.c_double * 6)]


_SIGS = {
    'spx_version': (C.c_int, []),
    'spx_last_error': (C.c_char_p, []),
    'spx_device_count': (C.c_int, []),
    'spx_parse_vg_str': (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_i32p, c_i32p, c_f64p, c_f64p]),
    'spx_fill_dists_2d_mat': (C.c_int, [c_f64p, c_f64p, C.c_int64, c_f64p, c_f64p, C.c_int64, c_f64p]),
    'spx_fill_vg_var_arr': (C.c_int, [c_f64p, c_f64p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      C.c_char_p, C.c_double]),
    'spx_copy_2d_arr_at_idxs': (C.c_int, [c_f64p, C.c_int64, C.c_int64, c_i64p, C.c_int64, c_i64p,
                                          C.c_int64, c_f64p, C.c_int64, C.c_int64]),
    'spx_fill_theo_vg_vals': (C.c_int, [C.c_char_p, c_f64p, C.c_int64, C.c_double, C.c_double, c_f64p]),
    'spx_fill_dists_one_pt': (C.c_int, [C.c_double, C.c_double, c_f64p, c_f64p, C.c_int64, c_f64p]),
    'spx_fill_wts_and_sum': (C.c_int, [c_f64p, c_f64p, C.c_int64, C.c_double, c_f64p]),
    'spx_get_mults_sum': (C.c_int, [c_f64p, c_f64p, C.c_int64, c_f64p]),
    'spx_fill_dists_2d_mat_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'spx_fill_vg_var_arr_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                          C.c_int, C.c_int, c_i32p, c_f64p, c_f64p, C.c_double,
                                          C.c_void_p]),
    'spx_copy_2d_arr_at_idxs_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p]),
    'spx_coef_offset': (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    'spx_krige_assemble_dev': (C.c_int, [C.POINTER(spx_systems), C.c_void_p, C.c_int, C.c_double,
                                         C.c_void_p]),
    'spx_krige_factor_dev': (C.c_int, [C.POINTER(spx_systems), C.c_void_p]),
    'spx_krige_solve_dev': (C.c_int, [C.POINTER(spx_systems), C.POINTER(spx_rhs), C.c_void_p]),
    'spx_krige_downdate_dev': (C.c_int, [C.POINTER(spx_downdate), C.c_void_p]),
    'spx_krige_downdate_max_r': (C.c_int, []),
    'spx_krige_downdate_reg_max_r': (C.c_int, []),
    'spx_avail_groups_host': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i32p]),
    'spx_downdate_plan_bytes': (C.c_int64, [C.c_int64, C.c_int32]),
    'spx_downdate_plan_host_bytes': (C.c_int64, [C.c_int64]),
    'spx_downdate_plan_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.POINTER(spx_dd_plan)]),
    'spx_avail_lists_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_build_bt_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    'spx_estimate_gemm_dev': (C.c_int, [C.POINTER(spx_gemm), C.c_void_p]),
    'spx_estimate_gemm_config': (C.c_int, [C.POINTER(spx_gemm), c_i32p, c_i32p, c_i32p, c_i32p]),
    'spx_pack_rows_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    'spx_nnb_index_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'spx_nnb_candidates_width': (C.c_int, []),
    'spx_nnb_candidates_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]),
    'spx_nnb_index_cand_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    'spx_nnb_gather_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_double, C.c_double, C.c_void_p]),
    'spx_fill_rows_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_double, C.c_double, C.c_void_p]),
    'spx_estimate_multivg_dev': (C.c_int, [C.POINTER(spx_multivg), C.c_void_p]),
    'spx_local_build_dev': (C.c_int, [C.POINTER(spx_local), C.c_void_p]),
    'spx_estimate_local_dev': (C.c_int, [C.POINTER(spx_local), C.c_void_p]),
    'spx_local_set_bulk': (C.c_int, [C.c_int]),
    'spx_gemm_set_ksplit': (C.c_int, [C.c_int]),
    'spx_local_tiles_dev': (C.c_int, [C.POINTER(spx_local), C.c_void_p]),
    'spx_nrst_max_neighbors': (C.c_int, []),
    'spx_nrst_set_topk_warp': (C.c_int, [C.c_int]),
    'spx_nrst_set_thread_rhs': (C.c_int, [C.c_int]),
    'spx_nrst_topk_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    'spx_pie_select_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    'spx_sel_equidist_refs': (C.c_int, [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_uint64, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_get_nd_dists': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    'spx_nrst_solve_dev': (C.c_int, [C.POINTER(spx_nrst), C.c_void_p]),
    'spx_nrst_krige_dev': (C.c_int, [C.POINTER(spx_nrst), C.c_void_p]),
    'spx_nrst_idw_dev': (C.c_int, [C.POINTER(spx_nrst), C.c_void_p, C.c_void_p]),
    'spx_mask_lists_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_void_p]),
    'spx_bcast_rows_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int32, C.c_void_p]),
    'spx_copy_to_mapped_host_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'spx_points_in_polygons_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_void_p, C.c_void_p]),
    'spx_points_in_polygons_chunk': (C.c_int, []),
    'spx_sample_raster_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_double, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    'spx_upload_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'spx_round_stats_workspace': (C.c_int64, [C.c_int64, C.c_int64]),
    'spx_round_stats_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_fast_slot_bytes': (C.c_int64, [C.POINTER(spx_fast_cfg), C.c_int32]),
    'spx_fast_create': (C.c_int, [C.POINTER(spx_fast_cfg), C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_void_p)]),
    'spx_fast_destroy': (C.c_int, [C.c_void_p]),
    'spx_fast_submit': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.POINTER(spx_fast_result)]),
    'spx_fast_check': (C.c_int, [C.c_void_p, C.c_int32, c_i32p]),
    'spx_fast_times': (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)]),
    'spx_fast_timeline': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    'spx_ut_gemm_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_pack_stride': (C.c_int64, [C.c_int64]),
    'spx_pack_field_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_unpack_field_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                        C.c_void_p, C.c_int64, C.c_int32]),
    'spx_lambda_check_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    'spx_sparse_cov_blocks_dev': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(spx_sparse_cov),
                                            C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    'spx_krige_sparse_ok_dev': (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                          C.POINTER(spx_sparse_cov), C.c_double, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    'spx_dpack_segments': (C.c_int64, [C.c_int64]),
    'spx_dpack_capacity': (C.c_int64, [C.c_int64, C.c_int64]),
    'spx_dpack_stats_workspace': (C.c_int64, [C.c_int64, C.c_int64]),
    'spx_dpack_field_dev': (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p, C.c_void_p]),
    'spx_dunpack_rows_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int32, C.c_void_p, C.c_int64, C.c_int32]),
}

EXPORTED = tuple(_SIGS)

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SpxError(
            f'{LIB_PATH} is missing: build it with `python -m spinterps_b200.build` '
            '(there is no CPU fallback)')
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().spx_last_error().decode(errors='replace')
        raise SpxError(f'{what or "spx call"} failed (code {rc}): {msg}')


def unpack_group_bits(bits, n_stn):
    """[n_grps, W] uint64 availability words -> bool [n_grps, n_stn]."""
    b8 = np.ascontiguousarray(bits).view(np.uint8)
    return np.unpackbits(b8, axis=1, count=n_stn, bitorder='little').view(np.bool_)


def avail_groups(data, min_var_thr=-np.inf, data_copy=None, want_mask=True):
    """Availability groups of a [T, N] float64 block on the host (spx_avail_groups_host;
    interp/grps.py:57-101).  Returns grp_of_step [T] int32, grp_mask [n_grps, N] bool
    (want_mask=False: the packed words [n_grps, ceil(N / 64)] uint64 instead, see
    unpack_group_bits), grp_n [n_grps] int64, grp_first [n_grps] int32, n_avail [T]
    int64, step_flag [T] bool.  data_copy: optional address of a writable float64 buffer
    of T*N elements that receives a dense copy of the data in the same pass."""
    data = np.asarray(data)
    assert data.dtype == np.float64 and data.ndim == 2
    T, N = data.shape
    assert T == 0 or (data.strides[1] == 8 and data.strides[0] % 8 == 0 and data.strides[0] >= 8 * N)
    W = (N + 63) // 64
    ints = np.empty((4, max(T, 1)), dtype=np.int32)     # grp_of_step, grp_first, grp_n, n_avail
    step_flag = np.empty(max(T, 1), dtype=np.uint8)
    grp_mask = np.empty((T, N), dtype=np.uint8) if want_mask else None
    grp_bits = None if want_mask else np.empty((max(T, 1), W), dtype=np.uint64)
    n_grps = C.c_int32(0)
    if T:
        check(load().spx_avail_groups_host(
            data.ctypes.data, T, N, data.strides[0] // 8, float(min_var_thr), ints[0].ctypes.data,
            ints[1].ctypes.data, ints[2].ctypes.data,
            grp_mask.ctypes.data if want_mask else None, ints[3].ctypes.data,
            step_flag.ctypes.data, None if data_copy is None else int(data_copy),
            None if want_mask else grp_bits.ctypes.data, C.byref(n_grps)), 'avail_groups_host')
    g = n_grps.value
    masks = grp_mask[:g].view(np.bool_) if want_mask else grp_bits[:g]
    return (ints[0, :T], masks, ints[2, :g].astype(np.int64), ints[1, :g],
            ints[3, :T].astype(np.int64), step_flag[:T].view(np.bool_))


def require_gpu():
    if load().spx_device_count() < 1:
        raise SpxError('no CUDA device visible: the spinterps_b200 compute path has no CPU fallback')


def f64p(a):
    return a.ctypes.data_as(c_f64p)


def i64p(a):
    return a.ctypes.data_as(c_i64p)


def parse_vg_str(vg_str, clamp_range=True):
    """-> list of (type_code, sill, range); raises SpxError on a malformed string."""
    lib = load()
    n = C.c_int32(0)
    types = (C.c_int32 * SPX_VG_MAX_TERMS)()
    sills = (C.c_double * SPX_VG_MAX_TERMS)()
    ranges = (C.c_double * SPX_VG_MAX_TERMS)()
    rc = lib.spx_parse_vg_str(str(vg_str).encode(), int(bool(clamp_range)), SPX_VG_MAX_TERMS,
                              C.byref(n), types, sills, ranges)
    check(rc, f'parse_vg_str({vg_str!r})')
    return [(int(types[i]), float(sills[i]), float(ranges[i])) for i in range(n.value)]


def make_vg(vg_str):
    v = spx_vg()
    terms = parse_vg_str(vg_str)
    v.n_terms = len(terms)
    for i, (t, s, r) in enumerate(terms):
        v.types[i] = t
        v.sills[i] = s
        v.ranges[i] = r
    return v


def vgs_to_numpy(vg_strs):
    arr = np.zeros(len(vg_strs), dtype=VG_DTYPE)
    for k, s in enumerate(vg_strs):
        terms = parse_vg_str(s)
        arr[k]['n_terms'] = len(terms)
        for i, (t, sill, rng) in enumerate(terms):
            arr[k]['types'][i] = t
            arr[k]['sills'][i] = sill
            arr[k]['ranges'][i] = rng
    return arr
