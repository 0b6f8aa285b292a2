/*
 * spx_b200.h -- C-ABI of the B200-native spinterps interpolation hot path.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature
 * (streams are passed as void* and may be NULL = default stream).
 *
 * Three groups of entry points:
 *
 *  (1) DROP-IN functions, HOST pointers.  One per `cpdef` free function of the
 *      reference's only native module, cyth/interpmthds.pyx (the functions
 *      interp/steps.py:14-19 and interp/grps.py:9 import).  Same argument
 *      meaning, caller-allocated outputs filled in place.  Each call copies its
 *      operands to the GPU, runs the sm_100a kernel and copies the result back.
 *
 *  (2) The same kernels on DEVICE pointers (suffix _dev) for callers that keep
 *      data resident in HBM.
 *
 *  (3) The ENGINE entry points (device pointers) that together implement the
 *      compute half of SpInterpSteps.interpolate_subset
 *      (interp/steps.py:478-877): kriging-system assembly, batched LU factor and
 *      solve, the fused variogram-fill + FP64 tensor-core (DMMA) estimate
 *      contraction, fused IDW, nearest-neighbour index and the field epilogues.
 *      spinterps_b200/engine.py drives them.
 *
 * Every function returns 0 on success, a negative SPX_E* code on failure;
 * spx_last_error() returns a thread-local message for the last failure.
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with SPX_ECUDA.
 */
#ifndef SPX_B200_H
#define SPX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_OK 0
#define SPX_EINVAL (-1)  /* bad argument                                   */
#define SPX_ECUDA (-2)   /* CUDA runtime error / no device                 */
#define SPX_EPARSE (-3)  /* malformed variogram string                     */
#define SPX_ENOMEM (-4)  /* shared-memory or workspace limit exceeded      */

/* Variogram families, cyth/interpmthds.pyx:88-95 */
#define SPX_VG_RNG 0
#define SPX_VG_NUG 1
#define SPX_VG_SPH 2
#define SPX_VG_EXP 3
#define SPX_VG_LIN 4
#define SPX_VG_GAU 5
#define SPX_VG_POW 6
#define SPX_VG_HOL 7
#define SPX_VG_MAX_TERMS 8

/* Kriging system kinds, interp/steps.py:181-191 */
#define SPX_KRG_OK 0
#define SPX_KRG_SK 1
#define SPX_KRG_EDK 2

int spx_version(void);
const char* spx_last_error(void);
/* Number of visible CUDA devices (0 if none); never fails. */
int spx_device_count(void);

/* Parse "<sill> <Name>(<range>) + ..." exactly like cyth/interpmthds.pyx:174-184
 * (split on '+', strip, one space between sill and name; range clamped to
 * >= 1e-5 when clamp_range != 0).  Outputs hold up to max_terms entries. */
int spx_parse_vg_str(const char* vg_models_str, int clamp_range, int max_terms,
                     int* n_terms, int* types, double* sills, double* ranges);

/* ---------------- (1) drop-in, host pointers ---------------------------- */

/* cyth/interpmthds.pyx:123-143  fill_dists_2d_mat(x1s, y1s, x2s, y2s, dists) */
int spx_fill_dists_2d_mat(const double* x1s, const double* y1s, int64_t n1,
                          const double* x2s, const double* y2s, int64_t n2,
                          double* dists /* [n1, n2] C-order */);

/* cyth/interpmthds.pyx:146-226  fill_vg_var_arr(dists, in_vars, covar_flag,
 * diag_mat_flag, vg_models_str, min_vg_val) */
int spx_fill_vg_var_arr(const double* dists, double* in_vars, int64_t rows, int64_t cols,
                        int covar_flag, int diag_mat_flag, const char* vg_models_str,
                        double min_vg_val);

/* cyth/interpmthds.pyx:229-248  copy_2d_arr_at_idxs(arr, row_idxs, col_idxs,
 * subset_arr); subset_arr may be wider/taller than the index lists
 * (subset_cols = its row pitch); untouched entries keep their value. */
int spx_copy_2d_arr_at_idxs(const double* arr, int64_t arr_rows, int64_t arr_cols,
                            const int64_t* row_idxs, int64_t n_row_idxs,
                            const int64_t* col_idxs, int64_t n_col_idxs,
                            double* subset_arr, int64_t subset_rows, int64_t subset_cols);

/* cyth/interpmthds.pyx:98-120  fill_theo_vg_vals(vg_str, h_arr, r, s, vg_arr)
 * -- ACCUMULATES into vg_arr. */
int spx_fill_theo_vg_vals(const char* vg_name, const double* h_arr, int64_t n,
                          double r, double s, double* vg_arr);

/* cyth/interpmthds.pyx:768-781  fill_dists_one_pt(x, y, xs, ys, dists) */
int spx_fill_dists_one_pt(double x, double y, const double* xs, const double* ys,
                          int64_t n, double* dists);

/* cyth/interpmthds.pyx:811-890  sel_equidist_refs: angular sector (ref_pie_idxs) of every
 * reference point around the destination, members per sector (ref_pie_cts) and the distance
 * rank of every point INSIDE its sector (ref_sel_pie_idxs); a point within min_dist_thresh
 * short-cuts to "only the nearest point, rank 0, everything else not_neb_flag" (sector
 * outputs are then left untouched, as in the reference).  The reference's DT_UL is
 * unsigned long: uint64_t on LP64.  tem_ref_sel_dists is the reference's scratch (unused).
 * Equal distances are ranked by index (the reference: np.argsort, unstable). */
int spx_sel_equidist_refs(double dst_x, double dst_y, const double* ref_xs, const double* ref_ys,
                          int64_t n_refs, uint64_t n_pies, double min_dist_thresh,
                          int64_t not_neb_flag, double* dists, double* tem_ref_sel_dists,
                          int64_t* ref_sel_pie_idxs, uint64_t* ref_pie_idxs,
                          uint64_t* ref_pie_cts);
/* cyth/interpmthds.pyx:893-925  get_nd_dists: Euclidean distance of every pair i > j of
 * pts[n_pts, n_dims], row by row: dists[i (i - 1) / 2 + j]. */
int spx_get_nd_dists(const double* pts, int64_t n_pts, int64_t n_dims, double* dists);

/* cyth/interpmthds.pyx:784-795  fill_wts_and_sum(dists, wts, idw_exp) -> sum */
int spx_fill_wts_and_sum(const double* dists, double* wts, int64_t n, double idw_exp,
                         double* wts_sum);

/* cyth/interpmthds.pyx:798-808  get_mults_sum(wts, data) -> sum */
int spx_get_mults_sum(const double* wts, const double* data, int64_t n, double* mults_sum);

/* ---------------- (2) same kernels, device pointers --------------------- */

int spx_fill_dists_2d_mat_dev(const double* x1s, const double* y1s, int64_t n1,
                              const double* x2s, const double* y2s, int64_t n2,
                              double* dists, void* stream);

int spx_fill_vg_var_arr_dev(const double* dists, double* in_vars, int64_t rows, int64_t cols,
                            int covar_flag, int diag_mat_flag, int n_terms, const int* types,
                            const double* sills, const double* ranges, double min_vg_val,
                            void* stream);

int spx_copy_2d_arr_at_idxs_dev(const double* arr, int64_t arr_cols,
                                const int64_t* row_idxs, int64_t n_row_idxs,
                                const int64_t* col_idxs, int64_t n_col_idxs,
                                double* subset_arr, int64_t subset_cols, void* stream);

/* ---------------- (3) engine, device pointers --------------------------- */

/* A variogram as numbers (host struct, passed by pointer). */
typedef struct spx_vg {
    int32_t n_terms;
    int32_t types[SPX_VG_MAX_TERMS];
    double sills[SPX_VG_MAX_TERMS];
    double ranges[SPX_VG_MAX_TERMS]; /* already clamped to >= 1e-5 */
} spx_vg;

/* Rows of the packed coefficient matrix are grouped in M-tiles of
 * SPX_BM rows; element (row, col) of a matrix with kpad columns lives at
 *   ((row / SPX_BM) * (kpad / 4) + col / 4) * (SPX_BM * 4)
 *     + ((row % SPX_BM) / 8) * 32 + (row % 8) * 4 + col % 4
 * i.e. 8x4 DMMA A-fragments stored lane-major so that one bulk copy brings a
 * whole (M-tile, k-chunk) stage into shared memory. */
#define SPX_BM 256
int64_t spx_coef_offset(int64_t row, int64_t col, int64_t kpad);

/* Batched kriging systems (one per availability group x variogram x kind).
 * All arrays are device pointers; per-system arrays have n_sys entries.
 *   sys_n      stations in the system          sys_kind   SPX_KRG_*
 *   sys_vg     index into vgs[]                sys_stn_off offset into stn_list
 *   sys_w_off  offset (doubles) of the column-major m x m workspace, m = n + k
 *   sys_piv_off offset (int32) of the pivot vector
 * Assembly follows interp/steps.py:170-243 with the variogram evaluated
 * directly from station coordinates (cyth/interpmthds.pyx:123-226 fused). */
typedef struct spx_systems {
    int32_t n_sys;
    int32_t n_drifts;
    const int32_t* sys_n;
    const int32_t* sys_kind;
    const int32_t* sys_vg;
    const int64_t* sys_stn_off;
    const int64_t* sys_w_off;
    const int64_t* sys_piv_off;
    const int32_t* stn_list;   /* station indices, ascending per system */
    const double* stn_x;
    const double* stn_y;
    const double* stn_drift;   /* [n_stn, n_drifts] or NULL */
    double* work;              /* matrices */
    int32_t* piv;              /* pivots */
    int32_t* info;             /* [n_sys] 0 ok, >0 = zero pivot at that column */
    int32_t max_m;             /* largest n + border in the batch (host-known; sizes the
                                  shared-memory panel of the blocked LU; 0 = unknown ->
                                  unblocked kernel) */
} spx_systems;

int spx_krige_assemble_dev(const spx_systems* s, const spx_vg* vgs_dev /* device array */, int n_vgs,
                           double min_vg_val, void* stream);

/* LU with partial pivoting, one thread block per system (replaces
 * np.linalg.pinv at interp/steps.py:351 for non-singular systems). */
int spx_krige_factor_dev(const spx_systems* s, void* stream);

/* Right-hand sides.  rhs_kind: 0 = data row (rhs_arg = step index into
 * data[n_steps, n_stn]; NaN never read because only available stations are
 * listed), 1 = ones over the stations, 2 = unit vector e_{rhs_arg}.
 * The solution x of A x = b is scattered into the packed coefficient matrix
 * row rhs_row: station entries to column stn_list[i], border entries to column
 * n_stn + i.  For kind 1 additionally resid[rhs] = ||x - e_n||_1 (the deviation
 * of sum(lambda) from 1 per unit of right-hand side). */
typedef struct spx_rhs {
    int32_t n_rhs;
    const int32_t* rhs_sys;
    const int32_t* rhs_kind;
    const int32_t* rhs_arg;
    const int64_t* rhs_row;
    const double* data;   /* [n_steps, n_stn] */
    int32_t n_stn;
    int32_t kpad;
    double* coef;         /* packed, zero-initialised by the caller */
    double* resid;        /* [n_rhs] or NULL */
    double* dense;        /* optional: x also written to dense[rhs * dense_ld + i] */
    int64_t dense_ld;
    int32_t coef_row_major; /* 0: packed fragment layout; 1: coef[row * kpad + col] */
} spx_rhs;

int spx_krige_solve_dev(const spx_systems* s, const spx_rhs* r, void* stream);

/* Downdated solves.  Every availability group's system A_g is a principal
 * sub-matrix of the full system F over all stations (+ border).  With
 * G = F^-1, kept rows K (stations of the group + border) and missing stations
 * Mi (r of them):
 *     A_g^-1 b = u_K - G[K, Mi] * (G[Mi, Mi]^-1 u_Mi),     u = G b~
 * (b~ = b zero-extended; block-inverse identity).  u for all right-hand sides
 * is one dense product Ut = B~^T G done by the caller; this kernel factors the
 * r x r matrix G[Mi, Mi] in SHARED memory (LU, partial pivoting), solves every
 * right-hand side of the system with warp-level substitutions and scatters the
 * coefficients exactly like spx_krige_solve_dev.  Cost per system O(r^3 + r n)
 * instead of O(n^3); used when r <= max_r fits in shared memory.
 * rhs_kind: 0 data, 1 ones-vector (resid[rhs] += ||x - e_n||_1; resid must be
 * zero-initialised).  Right-hand sides of one system are contiguous. */
typedef struct spx_downdate {
    int32_t n_sys;
    int32_t n_stn;
    int32_t n_border;
    int32_t max_r;
    const double* ginv;          /* [M, M], M = n_stn + n_border, symmetric */
    const int32_t* sys_r;
    const int64_t* sys_miss_off; /* into miss_list */
    const int32_t* miss_list;    /* missing station indices, ascending */
    const int32_t* sys_n;
    const int64_t* sys_stn_off;  /* into stn_list */
    const int32_t* stn_list;     /* kept station indices, ascending */
    const int64_t* sys_rhs_off;
    const int32_t* sys_rhs_cnt;
    const int32_t* rhs_urow;     /* row of ut */
    const int64_t* rhs_row;      /* packed coefficient row, < 0 = none */
    const int32_t* rhs_kind;
    const double* ut;            /* [n_urows, M] */
    int32_t kpad;
    double* coef;
    double* resid;               /* [n_rhs] */
    int32_t* info;               /* [n_sys] */
    int32_t coef_row_major;      /* 0: packed fragment layout; 1: coef[row * kpad + col] */
    const int32_t* sys_order;    /* optional [n_sys]: block b works on system sys_order[b]
                                    (largest r first shortens the tail); NULL = identity */
    /* Optional fused outputs for the local estimator (need max_r <=
     * spx_krige_downdate_reg_max_r(), n_border >= 1 and rhs_row = row index): */
    double* coef_t;              /* transposed copy coef_t[col * coef_t_ld + row]; entries of
                                    missing stations are written as zeros */
    int64_t coef_t_ld;
    double* base;                /* base[row] = base_f * sum_{stations} c + c[n_stn] */
    double base_f;
} spx_downdate;

int spx_krige_downdate_dev(const spx_downdate* d, void* stream);
/* Largest r that fits the device's shared memory. */
int spx_krige_downdate_max_r(void);
/* Largest r of the register-resident kernel (the one that writes coef_t / base). */
int spx_krige_downdate_reg_max_r(void);

/* ---- native host-side planning of a time chunk --------------------------------
 * Pure index work on HOST pointers (no GPU involved): the NumPy version of the same
 * logic cost more per chunk than the kernels it feeds. */

/* Availability groups of a [n_steps, n_stn] data block (row pitch ld, NaN = missing):
 * steps with identical sets of non-NaN stations share a group, groups numbered in
 * first-occurrence order (interp/grps.py:57-101).  Caller-allocated outputs with room
 * for n_steps groups: grp_first (first step of each group), grp_n (stations per
 * group), grp_mask [*, n_stn] (1 = available; optional) and / or grp_bits
 * [*, ceil(n_stn / 64)] (the same masks, bit j % 64 of word j / 64; optional); per step
 * n_avail and step_flag (any
 * available value >= min_var_thr, interp/steps.py:760-765).  data_copy (optional):
 * dense [n_steps, n_stn] copy of the data written in the same pass (e.g. a pinned
 * staging buffer for the upload). */
int spx_avail_groups_host(const double* data, int64_t n_steps, int32_t n_stn, int64_t ld,
                          double min_var_thr, int32_t* grp_of_step, int32_t* grp_first,
                          int32_t* grp_n, uint8_t* grp_mask, int32_t* n_avail,
                          uint8_t* step_flag, double* data_copy, uint64_t* grp_bits,
                          int32_t* n_grps_out);

/* Descriptor arrays of one spx_downdate call, packed into ONE buffer (upload it with a
 * single copy; device pointer = device base + off_*).  Systems = the groups that occur
 * among `steps` (ascending group id); right-hand sides of a system = its steps in the
 * order given (coefficient row rows[i]) followed by its ones-vector; rhs_urow indexes
 * the rows of Bt / Ut: data rows first (bt_step = their step), then one mask row per
 * system.  sys_order sorts by r descending; pos_ones[s] = rhs index of the ones-vector
 * of system s (host use: resid[pos_ones] is the sum(lambda) screening value). */
typedef struct spx_dd_plan {
    int32_t n_sys, n_data, n_rhs, max_r;
    int64_t total_r, total_n;
    int64_t off_sys_r, off_sys_miss_off, off_miss_list, off_sys_n, off_sys_stn_off,
        off_stn_list, off_sys_rhs_off, off_sys_rhs_cnt, off_rhs_urow, off_rhs_row,
        off_rhs_kind, off_sys_order, off_bt_step, off_sys_grp, off_pos_ones;
    int64_t n_upload_bytes;      /* host-written prefix (everything but the two station
                                    lists, which spx_avail_lists_dev fills on the device) */
    int64_t n_bytes;             /* bytes of the device buffer in use */
} spx_dd_plan;

/* Upper bounds for n_sel selected steps: of the whole (device) buffer, and of the
 * host-written prefix (the size `buf` of spx_downdate_plan_host must have). */
int64_t spx_downdate_plan_bytes(int64_t n_sel, int32_t n_stn);
int64_t spx_downdate_plan_host_bytes(int64_t n_sel);
int spx_downdate_plan_host(const int32_t* grp_of_step, const int32_t* grp_n, int32_t n_grps,
                           int32_t n_stn, const int32_t* steps, const int64_t* rows,
                           int64_t n_sel, void* buf, int64_t buf_bytes, spx_dd_plan* plan);
/* stn_list / miss_list of the plan: ascending available / missing station indices of
 * step src_step[s] (= the plan's bt_step + n_data) for every system s. */
int spx_avail_lists_dev(const double* data, int32_t n_stn, int64_t data_ld,
                        const int32_t* src_step, int32_t n_sys, const int64_t* stn_off,
                        int32_t* stn_list, const int64_t* miss_off, int32_t* miss_list,
                        void* stream);

/* Bt [n_rows, n_stn + n_border] of the downdate from the resident data block: row i <
 * n_data = data[src_step[i]] with NaN -> 0, row i >= n_data = availability mask (1.0 /
 * 0.0) of step src_step[i]; border columns zero.  (Ut = Bt . ginv is a plain GEMM.) */
int spx_build_bt_dev(const double* data, int32_t n_stn, int64_t data_ld, const int32_t* src_step,
                     int64_t n_rows, int64_t n_data, int32_t n_border, double* bt, void* stream);

/* The fused estimate contraction
 *     Z[row, cell] = sum_k coef[row, k] * B[k, cell]
 * B is never stored: each thread block generates its [kpad x cells] tile in
 * shared memory from coordinates (distance -> variogram / IDW weight; border
 * rows: ones, drifts), keeps it resident and sweeps all rows with DMMA
 * (mma.sync m8n8k4 f64) while coefficient stages arrive by bulk async copy.
 * Replaces interp/steps.py:403-435 (kriging) and :293-313 (IDW). */
#define SPX_GEN_VG 0   /* B[k, c] = vg(dist) (covar: sum sill - vg) */
#define SPX_GEN_IDW 1  /* B[k, c] = (dist / dist_scale) ** -idw_exp */

#define SPX_EPI_FIELD 0      /* out[row_dst, cell_pos] = clamp(Z)              */
#define SPX_EPI_AUX 1        /* aux[row_dst, cell] = Z  (f64)                   */
#define SPX_EPI_FIELD_DIV 2  /* out[row_dst, cell_pos] = clamp(Z / aux[row_aux, cell]) */
#define SPX_EPI_QUADFORM 3   /* rows = rows of ONE system's inverse, row_dst = their K index:
                                aux[quad_slot, cell] = sum_rows Z[row, cell] * (B[K(row), cell]
                                + [K(row) == n_stn])  = rhs' A^-1 rhs + lambda[n], the OK
                                estimation variance of interp/steps.py:431-434 */

typedef struct spx_gemm {
    const double* coef;      /* packed, segment start (row multiple of SPX_BM) */
    int64_t n_rows;          /* rows in the segment (last M-tile may be partial) */
    int32_t kpad;            /* multiple of 4, >= n_stn + n_border */
    int32_t n_stn;
    int32_t n_border;        /* 0 SK/IDW, 1 OK, 1 + n_drifts EDK */
    const double* stn_x;
    const double* stn_y;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    const double* cell_drift; /* [n_border - 1, n_cells] or NULL */
    int32_t gen;              /* SPX_GEN_* */
    int32_t covar_flag;
    spx_vg vg;
    double min_vg_val;
    double idw_exp;
    double dist_scale;
    int32_t epi;              /* SPX_EPI_* */
    const int32_t* row_dst;   /* [n_rows] output row (time index / aux slot), <0 skip */
    const int32_t* row_aux;   /* [n_rows] aux slot for SPX_EPI_FIELD_DIV */
    void* out;                /* float or double field [*, out_ld] */
    int64_t out_ld;
    int32_t out_f64;
    const int32_t* cell_pos;  /* [n_cells] column in the field, NULL = identity */
    double* aux;              /* [slots, n_cells] */
    int32_t has_lo, has_hi;
    double lo, hi;
    int32_t quad_slot;        /* SPX_EPI_QUADFORM: output row of aux */
} spx_gemm;

int spx_estimate_gemm_dev(const spx_gemm* g, void* stream);
/* Launch geometry actually used for a given problem (for reporting). */
int spx_estimate_gemm_config(const spx_gemm* g, int* cells_per_block, int* n_stages,
                             int* smem_bytes, int* grid);

/* Pack a dense row-major [n_rows, n_cols] matrix (NaN -> 0 if nan_to_zero;
 * mask_mode: write 1.0 where finite, 0.0 where NaN) into the packed
 * coefficient layout starting at row row0. */
int spx_pack_rows_dev(const double* src, int64_t src_ld, const int32_t* src_rows,
                      int64_t n_rows, int32_t n_cols, int32_t kpad, int mask_mode,
                      double* coef, int64_t row0, void* stream);

/* Nearest available station per (group, cell): argmin of the IEEE distance
 * sqrt(dx*dx + dy*dy), first index on ties (np.argmin, interp/steps.py:288).
 * grp_mask [n_grps, n_stn] uint8.  nnb [n_grps, n_cells] int32. */
int spx_nnb_index_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                      const uint8_t* grp_mask, int32_t n_grps,
                      const double* cell_x, const double* cell_y, int64_t n_cells,
                      int32_t* nnb, void* stream);

/* Same result through candidate lists: cand[n_cells, W] (W =
 * spx_nnb_candidates_width()) holds for every cell its W nearest stations among ALL
 * stations ordered by (distance, index); the nearest available station of a group
 * is the first available candidate (full scan only if all W are missing).  One
 * distance pass per chunk instead of one per availability group. */
int spx_nnb_candidates_width(void);
int spx_nnb_candidates_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                           const double* cell_x, const double* cell_y, int64_t n_cells,
                           int32_t* cand, void* stream);
int spx_nnb_index_cand_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                           const uint8_t* grp_mask, int32_t n_grps,
                           const double* cell_x, const double* cell_y, int64_t n_cells,
                           const int32_t* cand, int32_t* nnb, void* stream);

/* out[row_dst[r], cell_pos[c]] = clamp(data[row_step[r], nnb[row_grp[r], c]])
 * for every listed row; if fail != NULL only where fail[row_fail[r], c] != 0
 * (the kriging -> NNB fallback of interp/steps.py:418-426). */
int spx_nnb_gather_dev(const double* data, int32_t n_stn, const int32_t* nnb,
                       const int32_t* row_step, const int32_t* row_grp,
                       const int32_t* row_dst, int64_t n_rows,
                       const uint8_t* fail, const int32_t* row_fail,
                       int64_t n_cells, const int32_t* cell_pos,
                       void* out, int64_t out_ld, int32_t out_f64,
                       int32_t has_lo, int32_t has_hi, double lo, double hi, void* stream);

/* out[row_dst[r], cell_pos[:]] = clamp(vals[r])   (station-mean / single
 * station steps, interp/steps.py:282-283, :325-331, :312-313) */
int spx_fill_rows_dev(const double* vals, const int32_t* row_dst, int64_t n_rows,
                      int64_t n_cells, const int32_t* cell_pos,
                      void* out, int64_t out_ld, int32_t out_f64,
                      int32_t has_lo, int32_t has_hi, double lo, double hi, void* stream);

/* fail[slot, c] = !isclose(aux[slot, c], 1.0) (rtol 1e-5, atol 1e-8, NaN ->
 * fail) OR cell_bad[c] -- interp/steps.py:418. */
int spx_lambda_check_dev(const double* aux, int64_t n_slots, int64_t n_cells,
                         const uint8_t* cell_bad, uint8_t* fail, void* stream);

/* Ascending column indices of the set (want = 1) / clear (want = 0) entries of the
 * rows row_sel[0..n_sel) of a [*, n_cols] byte mask, written to out + off[i] (offsets
 * computed by the caller from the per-row counts): station / missing-station lists
 * of the availability groups (interp/grps.py:57-101) on the device. */
int spx_mask_lists_dev(const uint8_t* mask, int32_t n_cols, const int32_t* row_sel,
                       int32_t n_sel, const int64_t* off, int32_t want, int32_t* out,
                       void* stream);

/* out[row_dst[r], cell_pos[c]] = aux[row_slot[r], c] (0.0 if aux == NULL), restricted to
 * fail[row_fail[r], c] != 0 when fail != NULL; no clamp.  Spreads the per-system
 * estimation variance to the steps of the system (interp/steps.py:431-434, :821-831)
 * and zeroes it where kriging fell back to NNB (:425-426). */
int spx_bcast_rows_dev(const double* aux, const int32_t* row_slot, const int32_t* row_dst,
                       int64_t n_rows, const uint8_t* fail, const int32_t* row_fail,
                       int64_t n_cells, const int32_t* cell_pos, void* out, int64_t out_ld,
                       int32_t out_f64, void* stream);

/* ---- 'nrst' neighbour selection (interp/grps.py:147-166, :103-139) --------------
 * Every cell uses its k nearest available stations; cells with the same neighbour
 * set share one (k + border) system. */
typedef struct spx_nrst {
    int32_t n_grp;            /* cell groups = systems                              */
    int64_t n_cells;
    int32_t k, n_border, n_drifts, kind, n_stn;
    const int32_t* nbu;       /* [n_grp, k] neighbour station indices, ascending    */
    const int32_t* cell_grp;  /* [n_cells] system of each cell                      */
    const double* stn_x;
    const double* stn_y;
    const double* stn_drift;  /* [n_stn, n_drifts] or NULL                          */
    const double* cell_x;
    const double* cell_y;
    const double* cell_drift; /* [n_drifts, n_cells] or NULL                        */
    spx_vg vg;
    double min_vg_val;
    const double* data;       /* [*, n_stn]                                         */
    const int32_t* steps;     /* [n_t] rows of data / output rows                   */
    int32_t n_t;
    double min_var_thr;       /* interp/steps.py:760-765                            */
    const uint8_t* step_bypass; /* [n_t] 1 = nugget-only variogram (station mean)   */
    double* coef;             /* [n_grp, n_t + 1, k + n_border] dual coefficients   */
    double* ovr;              /* [n_grp, n_t] NaN = krige, else value to write      */
    int32_t* info;            /* [n_grp] LU status                                  */
    const int32_t* cell_pos;
    void* out;
    int64_t out_ld;
    int32_t out_f64, has_lo, has_hi;
    double lo, hi;
    double idw_exp;
    /* estimation variance (EST_VARS_OK, interp/steps.py:428-434): spx_nrst_solve_dev also
     * writes A^-1 of the systems u_beg .. u_end - 1 into inv, spx_nrst_krige_dev turns it
     * into sum(lambda * rhs) + lambda[n] per cell and writes ev_out next to out.  u_end = 0
     * means n_grp; cells of systems outside the range are left alone (the caller walks the
     * systems in slices that bound inv). */
    double* inv;              /* [u_end - u_beg, m, m] or NULL */
    void* ev_out;             /* field of the same layout as out, or NULL */
    int32_t u_beg, u_end;
} spx_nrst;

int spx_nrst_max_neighbors(void);
/* nb[n_cells, k]: the k nearest stations with mask[s] != 0 (all if mask == NULL) by
 * IEEE distance, indices ascending (np.sort(np.argsort(d)[:k])); hash[n_cells]: a
 * 63-bit hash of the row (cells are grouped by it, like grps.py:109-111). */
int spx_nrst_topk_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                      const uint8_t* mask, const double* cell_x, const double* cell_y,
                      int64_t n_cells, int32_t k, int32_t* nb, int64_t* hash, void* stream);
/* spx_nrst_topk_dev has two kernels with identical results: one warp per cell (distances
 * as 64-bit keys in shared memory, k-th smallest by bisection with warp-wide counting,
 * default) and one thread per cell (sorted insertion; used when the keys of four cells do
 * not fit into shared memory: n_stn > 6,400).  on: 1 / 0, -1 = environment SPX_TOPK_WARP
 * (default 1).  Returns the previous setting. */
int spx_nrst_set_topk_warp(int on);
/* spx_nrst_solve_dev runs the substitutions of the right-hand sides either with one thread
 * per right-hand side (up to 128 side by side in shared memory; systems up to ~70 unknowns,
 * default) or with one warp per right-hand side (any size); same operations in the same
 * order, identical coefficients.  on: 1 / 0, -1 = environment SPX_NRST_THREAD_RHS (default
 * 1).  Returns the previous setting. */
int spx_nrst_set_thread_rhs(int on);
/* 'pie' selection (interp/grps.py:168-247 with cyth/interpmthds.pyx:811-890): stations
 * binned into n_pies angular sectors around the cell, ranked by distance inside their
 * sector; nb = the first k stations in (rank, distance) order, indices ascending; hash
 * as above.  n_pies <= k. */
int spx_pie_select_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                       const uint8_t* mask, const double* cell_x, const double* cell_y,
                       int64_t n_cells, int32_t k, int32_t n_pies, int32_t* nb, int64_t* hash,
                       void* stream);
/* Assemble + LU + solve every step (and the ones-vector) of every cell-group system. */
int spx_nrst_solve_dev(const spx_nrst* n, void* stream);
/* Per cell: rhs from coordinates, sum(lambda) test, NNB fallback, estimate, store. */
int spx_nrst_krige_dev(const spx_nrst* n, void* stream);
/* Per cell: IDW over its own neighbour row nb[n_cells, k]. */
int spx_nrst_idw_dev(const spx_nrst* n, const int32_t* nb, void* stream);

/* Estimate with one variogram PER ROW (per-step variogram series):
 *   Z[row, cell] = sum_k coef[row, k] * vg_{row_vg[row]}(dist(station k, cell)) + border
 * coef is ROW-MAJOR [n_rows, kpad] here.  The distances of a 64-cell tile are
 * computed once into shared memory and every row re-evaluates only its variogram
 * on them (the tensor-core contraction would regenerate its whole operand tile
 * per variogram). */
typedef struct spx_multivg {
    const double* coef;
    int64_t n_rows;
    int32_t kpad, n_stn, n_border;
    const double* stn_x;
    const double* stn_y;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    const double* cell_drift;
    const spx_vg* vgs;         /* device table */
    const int32_t* row_vg;     /* [n_rows] */
    int32_t covar_flag;
    double min_vg_val;
    const int32_t* row_dst;    /* [n_rows] output row, < 0 skip */
    void* out;
    int64_t out_ld;
    int32_t out_f64;
    const int32_t* cell_pos;
    int32_t has_lo, has_hi;
    double lo, hi;
    int32_t all_fast;          /* every variogram uses only Nug/Sph/Exp/Lin/Gau terms */
} spx_multivg;
int spx_estimate_multivg_dev(const spx_multivg* g, void* stream);

/* LOCAL estimator for compactly supported variograms (only Nug / Sph / Lin terms):
 * beyond the largest range R the cell<->station variogram is the constant
 * F = sum(sills) (0 in covariance form) exactly, so
 *   Z[row, cell] = base[row] + drift terms + sum_{k: dist < R} coef[row, k] (vg(dist) - F),
 *   base[row] = F * sum_k coef[row, k] + coef[row, n_stn]   (the caller computes it).
 * spx_local_build_dev finds, through a uniform bin grid of size R over the stations,
 * the stations within R of every cell (cnt may exceed cap: rebuild with a larger
 * cap); spx_estimate_local_dev streams the field (HBM-write bound).
 * coef is ROW-MAJOR [n_rows, kpad]. */
typedef struct spx_local {
    const double* stn_x;
    const double* stn_y;
    const int32_t* bin_start;  /* [nbx * nby + 1] */
    const int32_t* bin_stn;    /* station ids ordered by bin */
    double x0, y0, inv_bin;
    int32_t nbx, nby;
    double R, F;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    int32_t cap;
    int32_t* cnt;              /* [n_cells] */
    int32_t* idx;              /* [cap, n_cells] station of the j-th near station of a cell */
    double* val;               /* [cap, n_cells] its vg(dist) - F */
    spx_vg vg;
    int32_t covar_flag;
    double min_vg_val;
    /* estimate */
    const double* coef;
    const double* base;
    int64_t n_rows;
    int32_t kpad, n_stn, n_drifts;
    const double* cell_drift;
    const int32_t* row_dst;
    void* out;
    int64_t out_ld;
    int32_t out_f64;
    const int32_t* cell_pos;
    int32_t has_lo, has_hi;
    double lo, hi;
    int32_t rows_all_valid;    /* every row_dst[r] >= 0 (enables the streamlined kernel) */
    const double* coef_t;      /* optional transposed copy [kpad, coef_t_ld] of coef (rows
                                  beyond n_rows zero); with f32 output, no drift, no cell_pos
                                  and rows_all_valid it selects the streamlined kernel */
    int64_t coef_t_ld;         /* multiple of 4, >= n_rows */
    /* Optional tile tables (spx_local_tiles_dev): per tile of SPX_LOCAL_TILE consecutive
     * cells the distinct near stations of its cells.  The streamlined kernel then stages
     * the tile's coefficient slices in shared memory and gathers from there. */
    int32_t* tile_cnt;         /* [n_tiles] stations of the tile, -1 = more than SPX_LOCAL_TILE_CAP */
    int32_t* tile_stn;         /* [n_tiles, SPX_LOCAL_TILE_CAP] ascending station ids */
    uint8_t* slot;             /* [cap, n_cells] position of idx[j, c] in its tile's list */
} spx_local;
#define SPX_LOCAL_TILE 256
#define SPX_LOCAL_TILE_CAP 32
int spx_local_build_dev(const spx_local* l, void* stream);
/* Fill tile_cnt / tile_stn / slot from cnt / idx (after spx_local_build_dev; n_stn <=
 * 65536). */
int spx_local_tiles_dev(const spx_local* l, void* stream);
int spx_estimate_local_dev(const spx_local* l, void* stream);
/* Streamlined kernel: write the field through shared-memory staged row segments and bulk
 * async (TMA) stores instead of per-lane streaming stores (needs 16-byte aligned rows:
 * out_ld % 4 == 0, n_cells % 4 == 0).  on: 1 / 0, -1 = environment SPX_LOCAL_BULK (default
 * 0: on B200 the staged variant measured 1.14 ms against 0.89 ms per 5 GB field, see
 * DESIGN.md).  Returns the previous setting.  Results are bit-identical either way. */
int spx_local_set_bulk(int on);
/* spx_estimate_gemm_dev splits a large K (kpad >~ 1000: at most 16 cells fit beside a full-K
 * tile) over up to max_passes launches whose partial sums travel through a stream-ordered f64
 * scratch in HBM; 0 / 1 = never, -1 = environment SPX_GEMM_KSPLIT (default 4).  Returns the
 * previous setting (test / measurement aid). */
int spx_gemm_set_ksplit(int max_passes);

/* ---- one native call per time chunk (the planned fast path) ------------------------
 * The common case of the compute half of SpInterpSteps.interpolate_subset
 * (interp/steps.py:673-833 for OK / EDK with neb_sel_mthd 'all'): every step of the chunk
 * shares ONE variogram whose full-station inverse `ginv` is resident, and every
 * availability group is solved by downdating it.  spx_fast_submit does, in one call,
 *   host:   availability groups + per-step flags (spx_avail_groups_host), the downdate
 *           descriptors (spx_downdate_plan_host), both written into pinned staging memory
 *   device: data block + descriptors upload, station lists, Ut = Bt . ginv (own DMMA
 *           kernel, Bt generated from the resident data), blocked LDL^T downdate, health
 *           flags copied to mapped host memory  -- all on the job's SOLVE stream --
 *           then the estimate (local estimator or DMMA contraction) on the caller's MAIN
 *           stream behind an event.
 * The solve stream is a separate high-priority stream: the latency-bound solves of chunk
 * i+1 run underneath the HBM-bound estimate of chunk i.  Buffers live in a ring of
 * n_slots slots; a slot is reused only after its estimate has finished.
 * Nothing here allocates per chunk.  Steps the path does not handle itself (no station,
 * one station, interp_steps_flag false) are only counted; the caller fills those rows. */
/* Sparse-covariance form of the ordinary-kriging systems of a compactly supported
 * variogram (Nug + Sph / Lin terms).  Beyond the largest range every variogram value is the
 * constant F, so the station matrix of interp/steps.py:199-216 is F 11' - C with C = F - vg
 * SPARSE, and the system [F 11' - C, 1; 1', 0] [x; nu] = [z; 0] reduces to
 * x = C^-1 (nu 1 - z), nu = (1' C^-1 z) / (1' C^-1 1).  C is block diagonal over the connected
 * components of the graph "stations closer than the range"; with a missing station the
 * blocks simply lose a row and a column.  When every component is small
 * (<= SPX_SPARSE_MAX_COMP stations) a time step costs O(n_stn) instead of a dense
 * factorisation: one warp per step solves the components (Cholesky in registers / local
 * memory), reduces nu and writes the coefficient row.  Same linear system, same solution
 * (to rounding) as the dense path. */
#define SPX_SPARSE_MAX_COMP 8
typedef struct spx_sparse_cov {
    int32_t n_comp;              /* 0 = not used */
    int32_t max_size;            /* largest component */
    int32_t n_single;            /* components are sorted by size: the first n_single have one
                                    station (comp_off[c] == blk_off[c] == c for them) */
    int32_t reserved;
    const int32_t* comp_off;     /* device [n_comp + 1]: members of component c are        */
    const int32_t* comp_stn;     /* device [n_stn]:       comp_stn[comp_off[c] .. comp_off[c+1]) */
    const int64_t* blk_off;      /* device [n_comp]: offset of the component's s x s block  */
    double* blk;                 /* device: C = F - vg(d_ij), row-major per component       */
} spx_sparse_cov;
/* Fill sp->blk from the station coordinates (diagonal: vg(0), i.e. the nugget -- quirk Q1). */
int spx_sparse_cov_blocks_dev(const double* stn_x, const double* stn_y, const spx_sparse_cov* sp,
                              const spx_vg* vg, double min_vg_val, double base_f, void* stream);
/* One warp per coefficient row: row i is the solution for time step row_step[i] of data
 * (device [*, n_stn] pitch ld, NaN = missing).  Writes coef [n_rows, kpad] row-major (columns
 * 0..n_stn: x and nu; the caller zeroes the padding), the transposed copy coef_t [kpad,
 * coef_t_ld] (may be NULL), base[i] = base_f * sum(x) + nu, and adds the number of rows whose
 * Cholesky failed to *info (device int32).  scratch: 2 * n_rows * n_stn doubles. */
int spx_krige_sparse_ok_dev(const double* data, int32_t n_stn, int64_t ld, const int32_t* row_step,
                            int64_t n_rows, const spx_sparse_cov* sp, double base_f, int32_t kpad,
                            double* coef, double* coef_t, int64_t coef_t_ld, double* base,
                            double* scratch, int32_t* info, void* stream);

typedef struct spx_fast_cfg {
    int32_t n_stn, n_border, kpad;
    int32_t max_steps;           /* largest chunk the job will see */
    int32_t n_slots;             /* ring depth, 2..8 */
    int32_t min_systems;         /* fewer systems than this: status 1 (not eligible) */
    double min_var_thr;          /* interp/steps.py:760-765 */
    const double* ginv;          /* device [M, M], M = n_stn + n_border, symmetric */
    double lambda_bound;         /* |rhs| bound of the sum(lambda) screening */
    double lambda_tol;
    int32_t estimator;           /* 0: local (template `local`), 1: contraction (`gemm`) */
    int32_t want_coef_t;         /* local: emit the transposed coefficients (streamlined kernel) */
    double base_f;               /* local: F of spx_local */
    spx_local local;             /* estimate templates: everything but coef / base / coef_t /
                                    n_rows / row_dst / out is taken from here */
    spx_gemm gemm;
    int32_t profile;             /* record events around the estimate launch (spx_fast_times) */
    int32_t solve_stream;        /* 1: solve phase on the job's own high-priority stream (runs
                                    beside the previous chunk's estimate), 0: on the caller's
                                    stream in front of the estimate */
    spx_sparse_cov sparse;       /* n_comp > 0 (estimator 0, n_border 1): the solve phase is
                                    spx_krige_sparse_ok_dev instead of Ut = Bt.G + downdate */
} spx_fast_cfg;

typedef struct spx_fast_result {
    int32_t status;              /* 0 queued, 1 not eligible (nothing was queued) */
    int32_t slot;
    int32_t n_grps, n_krige, n_sys, max_r;
    int32_t n_none, n_single, n_mean;   /* steps left to the caller */
    int32_t launches;            /* kernels queued by this call */
    int64_t h2d_bytes;
    const double* d_data;        /* the slot's device copy of the data block [n_steps, n_stn] */
    double host_ms[6];           /* host time of the call: slot wait, group scan, plan, upload
                                    calls, solve launches, estimate launch */
} spx_fast_result;

/* Device and pinned-host bytes one slot needs (the caller allocates n_slots of each). */
int64_t spx_fast_slot_bytes(const spx_fast_cfg* cfg, int32_t pinned_host);
/* dev_arena / host_arena: n_slots * spx_fast_slot_bytes(cfg, 0 / 1) bytes of device /
 * pinned (mapped) host memory owned by the caller for the lifetime of the job. */
int spx_fast_create(const spx_fast_cfg* cfg, void* dev_arena, void* host_arena, void** job);
int spx_fast_destroy(void* job);
/* data: HOST [n_steps, n_stn] row pitch ld (pageable or pinned).  out: device field
 * [*, out_ld] of the template's dtype.  Per-step outputs (host, caller-allocated, n_steps
 * entries; grp_bits [n_steps, ceil(n_stn / 64)]): as spx_avail_groups_host. */
int spx_fast_submit(void* job, const double* data, int64_t n_steps, int64_t ld, void* out,
                    void* main_stream, int32_t* grp_of_step, int32_t* grp_first,
                    int32_t* grp_n, int32_t* n_avail, uint8_t* step_flag, uint64_t* grp_bits,
                    spx_fast_result* res);
/* Waits for the slot's solve phase and reads its health flags:
 * verdict 0 healthy, 1 unhealthy elimination (info != 0), 2 sum(lambda) screening failed. */
int spx_fast_check(void* job, int32_t slot, int32_t* verdict);
/* Milliseconds of the slot's last estimate launch (profile = 1; waits for it). */
int spx_fast_times(void* job, int32_t slot, float* estimate_ms, float* solve_ms);
/* Timeline of a slot's last use relative to the solve start of slot `ref_slot` (profile = 1;
 * waits for the slot's estimate): t_ms[0..3] = solve start, solve end, estimate start,
 * estimate end.  A measuring aid for the overlap of the two streams. */
int spx_fast_timeline(void* job, int32_t ref_slot, int32_t slot, float* t_ms);

/* Ut[n_rows, M] = Bt . G with Bt generated from the resident data block exactly like
 * spx_build_bt_dev (row i < n_data: data[src_step[i]] with NaN -> 0, else the availability
 * mask; border columns zero) and G [M, M] symmetric; FP64 tensor-core (DMMA) tiles. */
int spx_ut_gemm_dev(const double* data, int32_t n_stn, int64_t data_ld, const int32_t* src_step,
                    int64_t n_rows, int64_t n_data, int32_t n_border, const double* ginv,
                    double* ut, void* stream);

/* ---- grid preparation ---------------------------------------------
 * Points (cells, stations) inside or within buffer_dist of polygons given as outer rings
 * (misc.py:407-540 chk_pt_cntmnt_in_polys_mp: OGR Contains on polygons buffered by the
 * station / cell buffer distance, one Python call per point).  Edges (x1, y1) -> (x2, y2)
 * of all rings, ering = ring id of each edge, non-decreasing; even-odd rule per ring,
 * union over rings; with buffer_dist > 0 also every point closer than buffer_dist to an
 * edge.  chunk_ymin / chunk_ymax (optional, both or neither): y-range of every chunk of
 * spx_points_in_polygons_chunk() consecutive edges, used to skip chunks that cannot touch
 * a block of points.  inside [n_pts] 0 / 1. */
int spx_points_in_polygons_dev(const double* px, const double* py, int64_t n_pts,
                               const double* ex1, const double* ey1, const double* ex2,
                               const double* ey2, const int32_t* ering, int64_t n_edges,
                               const double* chunk_ymin, const double* chunk_ymax,
                               double buffer_dist, uint8_t* inside, void* stream);
int spx_points_in_polygons_chunk(void);
/* out[i] = ras[rows[i], cols[i]] of a row-major [n_rows, n_cols] raster; NaN where the value
 * is np.isclose to the no-data value (has_ndv) or the index lies outside the raster
 * (interp/drift.py:165-226: drift values at cells and stations). */
int spx_sample_raster_dev(const double* ras, int64_t n_rows, int64_t n_cols, const int64_t* rows,
                          const int64_t* cols, int64_t n, double ndv, int32_t has_ndv,
                          double* out, void* stream);

/* Host -> device copy of n_bytes on `stream` (cudaMemcpyAsync; from pageable memory the call
 * returns once the source has been staged).  The engine's small per-chunk uploads go
 * through this on a dedicated upload stream. */
int spx_upload_dev(void* dst_dev, const void* src_host, int64_t n_bytes, void* stream);

/* Round a field to `decimals` decimal places in its own dtype, in place, exactly like
 * np.round (interp/steps.py:907-912: rint(x * 10^d) / 10^d; decimals < 0 = no rounding),
 * and return per-row statistics of the (rounded) values, NaN ignored:
 * stats[0..4][row] = min, mean, max, std (ddof 0), count of finite values -- what
 * interp/main.py:474-525 obtains by re-reading the netCDF file (there in the field
 * dtype; here accumulated in FP64).  Rows without any value give NaN, count 0.
 * workspace: spx_round_stats_workspace(n_rows, row_len) bytes of device memory. */
int64_t spx_round_stats_workspace(int64_t n_rows, int64_t row_len);
int spx_round_stats_dev(void* fld, int32_t is_f64, int64_t n_rows, int64_t row_len, int64_t ld,
                        int32_t decimals, double* stats, void* workspace, void* stream);

/* ---- lossless 2-byte transport of rounded f32 fields ------------------------------
 * The netCDF writer stores np.round(fld, nmrl_prcn) as float32 (interp/steps.py:907-945):
 * every value is fdiv(q, 10^d) for an integer q.  spx_pack_field_dev turns each row (time
 * step) of such a field into 16-bit codes q - qmin (0xFFFF = NaN, 0xFFFE = -0.0) after VERIFYING, element
 * by element, that the host's decode reproduces the float bit for bit; rows that do not
 * qualify (not rounded, |q| >= 2^31, infinities, range > 65533) get mode SPX_PACK_RAW and
 * are transferred as floats by the caller.  spx_unpack_field_host rebuilds the floats
 * (threaded, AVX2).  Halves the device -> host bytes of a chunk. */
#define SPX_PACK_U16 0
#define SPX_PACK_RAW 2
typedef struct spx_pack_row {
    int32_t mode;                /* SPX_PACK_U16 / SPX_PACK_RAW */
    int32_t qmin, qmax;          /* integer range of the row (qmin = offset of the codes) */
    int32_t n_nan;
} spx_pack_row;
/* Codes per row of the packed buffer (row_len rounded up to a multiple of 8). */
int64_t spx_pack_stride(int64_t row_len);
/* fld: device f32 [n_rows, row_len] pitch ld; hdr: device [n_rows]; codes: device
 * [n_rows, spx_pack_stride(row_len)] uint16.  decimals 0..9. */
int spx_pack_field_dev(const float* fld, int64_t n_rows, int64_t row_len, int64_t ld,
                       int32_t decimals, spx_pack_row* hdr, uint16_t* codes, void* stream);
/* HOST pointers.  Writes the rows of mode SPX_PACK_U16 into out [n_rows, row_len] pitch
 * out_ld (raw rows are left untouched: the caller copies them).  n_threads <= 0: default. */
int spx_unpack_field_host(const spx_pack_row* hdr, const uint16_t* codes, int64_t n_rows,
                          int64_t row_len, int32_t decimals, float* out, int64_t out_ld,
                          int32_t n_threads);

/* ---- lossless delta transport of rounded f32 fields (variable rate) ------------------
 * Same integer lattice as above, but neighbouring cells of an interpolated field differ by
 * few lattice steps: every row is cut into tiles of SPX_DPACK_TILE cells, the chain of q
 * along a tile is delta (first or second differences) + zigzag coded and bit-packed per
 * group of 8 cells with the group's own width (0..12, 14, 16 or 32 bits); NaN and -0.0 cells
 * sit in per-tile bitmaps, a tile with a value that is not on the lattice is stored as raw
 * floats.  The records of SPX_DPACK_SEGMENT cells (32 tiles) are contiguous in `payload`,
 * seg_off[row * segments + segment] is their offset in units of 4 bytes (record layout:
 * csrc/spx_pack.cu).  Typical interpolated fields need 0.2-0.6 bytes per cell instead of 4;
 * the host decode returns the identical floats (what interp/steps.py:907-945 writes).
 *
 * With SPX_DPACK_ROUND the call is the writer's whole output stage in one pass over the
 * UNROUNDED field: np.round(fld, decimals) in float32 (interp/steps.py:907-912) -- the
 * transported integers are np.round's own intermediate rint(x * 10^d) --, the per-step
 * statistics of interp/main.py:474-525 (stats != NULL, layout as spx_round_stats_dev), the
 * encoding; SPX_DPACK_WRITE_BACK also stores the rounded values to fld.  Without
 * SPX_DPACK_ROUND the field must already be rounded; every value is then checked to
 * round-trip bit for bit (else its tile travels raw) and fld is only read. */
#define SPX_DPACK_TILE 256
#define SPX_DPACK_SEGMENT 8192
#define SPX_DPACK_ROUND 1
#define SPX_DPACK_WRITE_BACK 2
/* Segments per row; worst-case payload bytes of a field (every tile raw); bytes of the
 * device workspace the statistics need. */
int64_t spx_dpack_segments(int64_t row_len);
int64_t spx_dpack_capacity(int64_t n_rows, int64_t row_len);
int64_t spx_dpack_stats_workspace(int64_t n_rows, int64_t row_len);
/* fld: device f32 [n_rows, row_len] pitch ld; decimals 0..9; stats: device double
 * [5, n_rows] or NULL (then workspace may be NULL); seg_off: device uint32
 * [n_rows * spx_dpack_segments(row_len)]; payload: device buffer of capacity_bytes (any size:
 * segments that do not fit are not written, their seg_off is 0xFFFFFFFF); counters: device
 * uint64[2], written by the call: [0] = 4-byte words the records need in total (may exceed
 * the capacity), [1] = 1 if a segment did not fit.  The order of the segments in the payload
 * is not deterministic; their content and the decoded field are. */
int spx_dpack_field_dev(float* fld, int64_t n_rows, int64_t row_len, int64_t ld,
                        int32_t decimals, int32_t flags, double* stats, void* workspace,
                        uint32_t* seg_off, void* payload, int64_t capacity_bytes,
                        uint64_t* counters, void* stream);
/* HOST pointers.  Decodes n_rows rows (seg_off points at the first of them) into out
 * [n_rows, row_len] pitch out_ld; payload_bytes bounds every access.  n_threads <= 0:
 * default, 1: on the calling thread (safe to call from several threads at once). */
int spx_dunpack_rows_host(const uint32_t* seg_off, const void* payload, int64_t payload_bytes,
                          int64_t n_rows, int64_t row_len, int32_t decimals, float* out,
                          int64_t out_ld, int32_t n_threads);

/* Copy a small device buffer into pinned (UVA-mapped) host memory with a kernel
 * instead of a DMA engine, so that the copy cannot queue behind a large field
 * download in flight; n_bytes and both pointers multiples of 4. */
int spx_copy_to_mapped_host_dev(void* dst_host_mapped, const void* src_dev, int64_t n_bytes,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPX_B200_H */
