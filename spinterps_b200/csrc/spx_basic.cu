// Drop-in equivalents of the free functions of cyth/interpmthds.pyx:
// distance fill, variogram fill, 2-D gather, IDW weights / dot product.
// Host-pointer shims (copy in, kernel, copy out) + device-pointer variants.
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "spx_common.cuh"

namespace spx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SPX_ECUDA;
}

// ---------------------------------------------------------------- kernels

// dists[i, j] for a [n1, n2] C-order matrix.  One thread per 2 consecutive
// columns (16-byte stores, coalesced along the row); the x2/y2 coordinates of
// the column tile are staged in shared memory and reused by every row of the
// block's row strip.
constexpr int DT_COLS = 512;  // columns per block tile
constexpr int DT_ROWS = 16;   // rows per block

__global__ void __launch_bounds__(256) k_fill_dists(const double* __restrict__ x1s,
                                                    const double* __restrict__ y1s, int64_t n1,
                                                    const double* __restrict__ x2s,
                                                    const double* __restrict__ y2s, int64_t n2,
                                                    double* __restrict__ dists) {
    __shared__ double sx[DT_COLS];
    __shared__ double sy[DT_COLS];
    const int64_t c0 = (int64_t)blockIdx.x * DT_COLS;
    const int64_t r0 = (int64_t)blockIdx.y * DT_ROWS;
    for (int c = threadIdx.x; c < DT_COLS; c += blockDim.x) {
        const int64_t cc = c0 + c;
        sx[c] = cc < n2 ? x2s[cc] : 0.0;
        sy[c] = cc < n2 ? y2s[cc] : 0.0;
    }
    __syncthreads();
    const bool vec_ok = ((n2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(dists) & 15) == 0);  // 16-byte aligned rows
    for (int r = 0; r < DT_ROWS; ++r) {
        const int64_t row = r0 + r;
        if (row >= n1) break;
        const double x = x1s[row];
        const double y = y1s[row];
        double* out = dists + row * n2;
        const int c = threadIdx.x * 2;
        const int64_t cc = c0 + c;
        if (cc + 1 < n2 && vec_ok) {
            double2 v;
            v.x = dist_rn(x, y, sx[c], sy[c]);
            v.y = dist_rn(x, y, sx[c + 1], sy[c + 1]);
            *reinterpret_cast<double2*>(out + cc) = v;
        } else {
            if (cc < n2) out[cc] = dist_rn(x, y, sx[c], sy[c]);
            if (cc + 1 < n2) out[cc + 1] = dist_rn(x, y, sx[c + 1], sy[c + 1]);
        }
    }
}

// in_vars = vg(dists) elementwise.  diag_mat_flag only changes WHICH entry of a
// symmetric pair is evaluated (upper triangle, mirrored to the lower), so the
// lower triangle reads the transposed distance.
__global__ void __launch_bounds__(256) k_fill_vg(const double* __restrict__ dists,
                                                 double* __restrict__ in_vars, int64_t rows,
                                                 int64_t cols, int covar_flag, int diag_mat_flag,
                                                 VgDev vg, double min_vg_val) {
    const int64_t n = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double h = dists[i];
        if (diag_mat_flag) {
            const int64_t r = i / cols, c = i - r * cols;
            if (c < r) h = dists[c * cols + r];
        }
        in_vars[i] = vg_eval(vg, h, covar_flag, min_vg_val);
    }
}

// Streaming variant for rectangular (dst <-> station) matrices and the
// division-free variogram families: 4 elements per thread as two 16-byte
// accesses in flight, 16 bytes of HBM traffic per element.
__global__ void __launch_bounds__(256) k_fill_vg_fast(const double2* __restrict__ dists,
                                                      double2* __restrict__ in_vars,
                                                      int64_t n_pairs, int covar_flag,
                                                      VgFast vg, double min_vg_val) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n_pairs;
         i += stride) {
        const double2 a = dists[i];
        const bool two = (i + 1 < n_pairs);
        const double2 b = two ? dists[i + 1] : a;
        double2 ra, rb;
        ra.x = vg_eval_fast(vg, a.x, covar_flag, min_vg_val);
        ra.y = vg_eval_fast(vg, a.y, covar_flag, min_vg_val);
        rb.x = vg_eval_fast(vg, b.x, covar_flag, min_vg_val);
        rb.y = vg_eval_fast(vg, b.y, covar_flag, min_vg_val);
        in_vars[i] = ra;
        if (two) in_vars[i + 1] = rb;
    }
}

// sub[r, c] = arr[row_idxs[r], col_idxs[c]]: threads walk the columns of a row
// strip (coalesced stores, near-coalesced loads when the column list is mostly
// contiguous); 4 rows per thread keep 4 independent loads in flight.
constexpr int GA_ROWS = 4;

__global__ void __launch_bounds__(256) k_gather_2d(const double* __restrict__ arr,
                                                   int64_t arr_cols,
                                                   const int64_t* __restrict__ row_idxs,
                                                   int64_t n_rows,
                                                   const int64_t* __restrict__ col_idxs,
                                                   int64_t n_cols, double* __restrict__ sub,
                                                   int64_t sub_cols) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t src_c = col_idxs[c];
    for (int64_t r0 = (int64_t)blockIdx.y * GA_ROWS; r0 < n_rows;
         r0 += (int64_t)gridDim.y * GA_ROWS) {
        double v[GA_ROWS];
#pragma unroll
        for (int k = 0; k < GA_ROWS; ++k)
            if (r0 + k < n_rows) v[k] = arr[row_idxs[r0 + k] * arr_cols + src_c];
#pragma unroll
        for (int k = 0; k < GA_ROWS; ++k)
            if (r0 + k < n_rows) sub[(r0 + k) * sub_cols + c] = v[k];
    }
}

__global__ void k_theo_vg(int type, const double* __restrict__ h, int64_t n, double r, double s,
                          double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += vg_term(type, h[i], r, s);
}

__global__ void k_dists_one_pt(double x, double y, const double* __restrict__ xs,
                               const double* __restrict__ ys, int64_t n,
                               double* __restrict__ dists) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dists[i] = dist_rn(x, y, xs[i], ys[i]);
}

__global__ void k_idw_wts(const double* __restrict__ dists, double* __restrict__ wts, int64_t n,
                          double p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) wts[i] = 1.0 / pow(dists[i], p);
}

// Sequential-order sums (cyth/interpmthds.pyx:791-793, :805-806 add left to
// right): a single thread walks the vector so that the rounding sequence is the
// reference's.  These two functions are drop-in shims, not the hot path (the
// engine uses the fused IDW contraction).
__global__ void k_seq_sum(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                          double* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        if (b == nullptr) {
            for (int64_t i = 0; i < n; ++i) s = __dadd_rn(s, a[i]);
        } else {
            for (int64_t i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn(a[i], b[i]));
        }
        *out = s;
    }
}

// cyth/interpmthds.pyx:811-890 for one destination point: thread per reference point.
__global__ void k_sel_equidist(double x, double y, const double* __restrict__ xs,
                               const double* __restrict__ ys, int n, int n_pies,
                               double min_dist_thresh, long long not_neb_flag,
                               double* __restrict__ dists, long long* __restrict__ sel,
                               unsigned long long* __restrict__ pidx,
                               unsigned long long* __restrict__ cts) {
    extern __shared__ double sd[];       // [n] distances, then [n] sectors (as int)
    int* sp = reinterpret_cast<int*>(sd + n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sd[i] = dist_rn(x, y, xs[i], ys[i]);
        sp[i] = pie_sector(__dsub_rn(xs[i], x), __dsub_rn(ys[i], y), n_pies);
        dists[i] = sd[i];
        sel[i] = not_neb_flag;
    }
    __syncthreads();
    // a reference point within the threshold: only the nearest such point is selected
    __shared__ int s_near;
    if (threadIdx.x == 0) {
        int best = -1;
        for (int i = 0; i < n; ++i)
            if (sd[i] <= min_dist_thresh && (best < 0 || sd[i] < sd[best])) best = i;
        s_near = best;
        if (best >= 0) sel[best] = 0;
    }
    __syncthreads();
    if (s_near >= 0) return;
    for (int j = threadIdx.x; j < n_pies; j += blockDim.x) {
        unsigned long long c = 0;
        for (int i = 0; i < n; ++i) c += (sp[i] == j);
        cts[j] = c;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        pidx[i] = (unsigned long long)sp[i];
        long long rank = 0;              // points of the same sector that come first
        for (int q = 0; q < n; ++q)
            if (sp[q] == sp[i] && (sd[q] < sd[i] || (sd[q] == sd[i] && q < i))) ++rank;
        sel[i] = rank;
    }
}

// cyth/interpmthds.pyx:893-925: pair (i > j) -> slot i (i - 1) / 2 + j.
__global__ void k_nd_dists(const double* __restrict__ pts, int64_t n_pts, int64_t n_dims,
                           double* __restrict__ out) {
    const int64_t i = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= i) return;
    double d = 0.0;
    for (int64_t k = 0; k < n_dims; ++k) {
        const double t = __dsub_rn(pts[i * n_dims + k], pts[j * n_dims + k]);
        d = __dadd_rn(d, __dmul_rn(t, t));
    }
    out[i * (i - 1) / 2 + j] = __dsqrt_rn(d);
}

// ---------------------------------------------------------------- helpers

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        SPX_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return SPX_OK;
    }
    template <typename T>
    T* as() {
        return reinterpret_cast<T*>(p);
    }
};

static int upload(DevBuf& b, const void* src, size_t bytes) {
    int rc = b.alloc(bytes);
    if (rc) return rc;
    if (bytes) SPX_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return SPX_OK;
}

static int vg_type_from_name(const std::string& n) {
    static const char* names[] = {"Rng", "Nug", "Sph", "Exp", "Lin", "Gau", "Pow", "Hol"};
    for (int i = 0; i < 8; ++i)
        if (n == names[i]) return i;
    return -1;
}

static std::string strip(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) ++b;
    while (e > b && isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}

static bool to_double(const std::string& s, double* out) {
    const std::string t = strip(s);
    if (t.empty()) return false;
    char* end = nullptr;
    *out = strtod(t.c_str(), &end);
    return end && *end == '\0';
}

}  // namespace spx

using namespace spx;

// ---------------------------------------------------------------- C ABI

extern "C" {

int spx_version(void) { return 100; }

const char* spx_last_error(void) { return g_err; }

int spx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int spx_parse_vg_str(const char* vg_models_str, int clamp_range, int max_terms, int* n_terms,
                     int* types, double* sills, double* ranges) {
    if (!vg_models_str || !n_terms || !types || !sills || !ranges) {
        set_error("spx_parse_vg_str: null argument");
        return SPX_EINVAL;
    }
    const std::string all(vg_models_str);
    int cnt = 0;
    size_t pos = 0;
    while (true) {
        size_t nxt = all.find('+', pos);
        std::string term = strip(all.substr(pos, nxt == std::string::npos ? nxt : nxt - pos));
        // "<sill> <Name>(<range>)": exactly one space (pyx:178 unpacks 2 parts)
        const size_t sp = term.find(' ');
        if (sp == std::string::npos || term.find(' ', sp + 1) != std::string::npos) {
            set_error("malformed variogram term '%s' in '%s'", term.c_str(), vg_models_str);
            return SPX_EPARSE;
        }
        const std::string sill_s = term.substr(0, sp);
        const std::string rest = term.substr(sp + 1);
        const size_t lp = rest.find('(');
        if (lp == std::string::npos || rest.find('(', lp + 1) != std::string::npos) {
            set_error("malformed variogram term '%s' in '%s'", term.c_str(), vg_models_str);
            return SPX_EPARSE;
        }
        const std::string name = rest.substr(0, lp);
        std::string range_s = rest.substr(lp + 1);
        const size_t rp = range_s.find(')');
        if (rp != std::string::npos) range_s = range_s.substr(0, rp);
        double sill, rng;
        const int type = vg_type_from_name(name);
        if (type < 0 || !to_double(sill_s, &sill) || !to_double(range_s, &rng)) {
            set_error("malformed variogram term '%s' in '%s'", term.c_str(), vg_models_str);
            return SPX_EPARSE;
        }
        if (cnt >= max_terms) {
            set_error("more than %d nested variogram terms in '%s'", max_terms, vg_models_str);
            return SPX_EINVAL;
        }
        if (clamp_range && !(rng > 1e-5)) rng = (rng != rng) ? rng : 1e-5;  // max(1e-5, r)
        types[cnt] = type;
        sills[cnt] = sill;
        ranges[cnt] = rng;
        ++cnt;
        if (nxt == std::string::npos) break;
        pos = nxt + 1;
    }
    *n_terms = cnt;
    return SPX_OK;
}

// ------------------------------------------------------- device variants

int spx_fill_dists_2d_mat_dev(const double* x1s, const double* y1s, int64_t n1, const double* x2s,
                              const double* y2s, int64_t n2, double* dists, void* stream) {
    if (n1 < 0 || n2 < 0) {
        set_error("fill_dists_2d_mat: negative size");
        return SPX_EINVAL;
    }
    if (n1 == 0 || n2 == 0) return SPX_OK;
    dim3 grid((unsigned)((n2 + DT_COLS - 1) / DT_COLS), (unsigned)((n1 + DT_ROWS - 1) / DT_ROWS));
    if (grid.y > 65535) {
        // fold very tall matrices: loop over row super-blocks
        const int64_t rows_per = (int64_t)65535 * DT_ROWS;
        for (int64_t r0 = 0; r0 < n1; r0 += rows_per) {
            const int64_t nr = (n1 - r0 < rows_per) ? (n1 - r0) : rows_per;
            dim3 g2(grid.x, (unsigned)((nr + DT_ROWS - 1) / DT_ROWS));
            k_fill_dists<<<g2, 256, 0, (cudaStream_t)stream>>>(x1s + r0, y1s + r0, nr, x2s, y2s,
                                                               n2, dists + r0 * n2);
        }
    } else {
        k_fill_dists<<<grid, 256, 0, (cudaStream_t)stream>>>(x1s, y1s, n1, x2s, y2s, n2, dists);
    }
    SPX_CHECK_LAUNCH("k_fill_dists");
    return SPX_OK;
}

int spx_fill_vg_var_arr_dev(const double* dists, double* in_vars, int64_t rows, int64_t cols,
                            int covar_flag, int diag_mat_flag, int n_terms, const int* types,
                            const double* sills, const double* ranges, double min_vg_val,
                            void* stream) {
    if (n_terms < 0 || n_terms > SPX_VG_MAX_TERMS) {
        set_error("fill_vg_var_arr: %d variogram terms (max %d)", n_terms, SPX_VG_MAX_TERMS);
        return SPX_EINVAL;
    }
    if (diag_mat_flag && rows != cols) {
        set_error("fill_vg_var_arr: diag_mat_flag needs a square matrix");
        return SPX_EINVAL;
    }
    if (rows * cols == 0) return SPX_OK;
    VgDev vg{};
    vg.n_terms = n_terms;
    for (int i = 0; i < n_terms; ++i) {
        vg.types[i] = types[i];
        vg.sills[i] = sills[i];
        vg.ranges[i] = ranges[i];
    }
    const int64_t n = rows * cols;
    const VgFast vf = make_vg_fast(n_terms, types, sills, ranges);
    const bool aligned = ((reinterpret_cast<uintptr_t>(dists) | reinterpret_cast<uintptr_t>(in_vars)) & 15) == 0;
    if (vf.all_fast && !diag_mat_flag && aligned && (n % 2 == 0)) {
        const int64_t n_pairs = n / 2;
        const int64_t want = (n_pairs / 2 + 255) / 256;
        const int blocks = (int)(want < 148 * 32 ? (want < 1 ? 1 : want) : 148 * 32);
        k_fill_vg_fast<<<blocks, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const double2*>(dists), reinterpret_cast<double2*>(in_vars), n_pairs,
            covar_flag, vf, min_vg_val);
        SPX_CHECK_LAUNCH("k_fill_vg_fast");
        return SPX_OK;
    }
    int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_fill_vg<<<blocks, 256, 0, (cudaStream_t)stream>>>(dists, in_vars, rows, cols, covar_flag,
                                                        diag_mat_flag, vg, min_vg_val);
    SPX_CHECK_LAUNCH("k_fill_vg");
    return SPX_OK;
}

int spx_copy_2d_arr_at_idxs_dev(const double* arr, int64_t arr_cols, const int64_t* row_idxs,
                                int64_t n_row_idxs, const int64_t* col_idxs, int64_t n_col_idxs,
                                double* subset_arr, int64_t subset_cols, void* stream) {
    const int64_t n = n_row_idxs * n_col_idxs;
    if (n == 0) return SPX_OK;
    const int64_t row_blocks = (n_row_idxs + GA_ROWS - 1) / GA_ROWS;
    dim3 grid((unsigned)((n_col_idxs + 255) / 256),
              (unsigned)(row_blocks < 65535 ? row_blocks : 65535));
    k_gather_2d<<<grid, 256, 0, (cudaStream_t)stream>>>(arr, arr_cols, row_idxs, n_row_idxs,
                                                        col_idxs, n_col_idxs, subset_arr,
                                                        subset_cols);
    SPX_CHECK_LAUNCH("k_gather_2d");
    return SPX_OK;
}

// ------------------------------------------------------- host shims

int spx_fill_dists_2d_mat(const double* x1s, const double* y1s, int64_t n1, const double* x2s,
                          const double* y2s, int64_t n2, double* dists) {
    if (n1 < 0 || n2 < 0) {
        set_error("fill_dists_2d_mat: negative size");
        return SPX_EINVAL;
    }
    if (n1 == 0 || n2 == 0) return SPX_OK;
    DevBuf a, b, c, d, o;
    int rc;
    if ((rc = upload(a, x1s, n1 * 8)) || (rc = upload(b, y1s, n1 * 8)) ||
        (rc = upload(c, x2s, n2 * 8)) || (rc = upload(d, y2s, n2 * 8)) ||
        (rc = o.alloc((size_t)n1 * n2 * 8)))
        return rc;
    rc = spx_fill_dists_2d_mat_dev(a.as<double>(), b.as<double>(), n1, c.as<double>(),
                                   d.as<double>(), n2, o.as<double>(), nullptr);
    if (rc) return rc;
    SPX_CUDA(cudaMemcpy(dists, o.p, (size_t)n1 * n2 * 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_fill_vg_var_arr(const double* dists, double* in_vars, int64_t rows, int64_t cols,
                        int covar_flag, int diag_mat_flag, const char* vg_models_str,
                        double min_vg_val) {
    int n_terms, types[SPX_VG_MAX_TERMS];
    double sills[SPX_VG_MAX_TERMS], ranges[SPX_VG_MAX_TERMS];
    int rc = spx_parse_vg_str(vg_models_str, 1, SPX_VG_MAX_TERMS, &n_terms, types, sills, ranges);
    if (rc) return rc;
    if (rows < 0 || cols < 0) {
        set_error("fill_vg_var_arr: negative size");
        return SPX_EINVAL;
    }
    if (rows * cols == 0) return SPX_OK;
    DevBuf d, o;
    const size_t bytes = (size_t)rows * cols * 8;
    if ((rc = upload(d, dists, bytes)) || (rc = o.alloc(bytes))) return rc;
    rc = spx_fill_vg_var_arr_dev(d.as<double>(), o.as<double>(), rows, cols, covar_flag,
                                 diag_mat_flag, n_terms, types, sills, ranges, min_vg_val,
                                 nullptr);
    if (rc) return rc;
    SPX_CUDA(cudaMemcpy(in_vars, o.p, bytes, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_copy_2d_arr_at_idxs(const double* arr, int64_t arr_rows, int64_t arr_cols,
                            const int64_t* row_idxs, int64_t n_row_idxs, const int64_t* col_idxs,
                            int64_t n_col_idxs, double* subset_arr, int64_t subset_rows,
                            int64_t subset_cols) {
    if (n_row_idxs > subset_rows || n_col_idxs > subset_cols) {
        set_error("copy_2d_arr_at_idxs: subset_arr smaller than the index lists");
        return SPX_EINVAL;
    }
    for (int64_t i = 0; i < n_row_idxs; ++i)
        if (row_idxs[i] < 0 || row_idxs[i] >= arr_rows) {
            set_error("copy_2d_arr_at_idxs: row index %lld out of range", (long long)row_idxs[i]);
            return SPX_EINVAL;
        }
    for (int64_t i = 0; i < n_col_idxs; ++i)
        if (col_idxs[i] < 0 || col_idxs[i] >= arr_cols) {
            set_error("copy_2d_arr_at_idxs: column index %lld out of range",
                      (long long)col_idxs[i]);
            return SPX_EINVAL;
        }
    if (n_row_idxs * n_col_idxs == 0) return SPX_OK;
    DevBuf a, r, c, s;
    int rc;
    const size_t sub_bytes = (size_t)subset_rows * subset_cols * 8;
    if ((rc = upload(a, arr, (size_t)arr_rows * arr_cols * 8)) ||
        (rc = upload(r, row_idxs, n_row_idxs * 8)) || (rc = upload(c, col_idxs, n_col_idxs * 8)) ||
        (rc = upload(s, subset_arr, sub_bytes)))
        return rc;
    rc = spx_copy_2d_arr_at_idxs_dev(a.as<double>(), arr_cols, r.as<int64_t>(), n_row_idxs,
                                     c.as<int64_t>(), n_col_idxs, s.as<double>(), subset_cols,
                                     nullptr);
    if (rc) return rc;
    SPX_CUDA(cudaMemcpy(subset_arr, s.p, sub_bytes, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_fill_theo_vg_vals(const char* vg_name, const double* h_arr, int64_t n, double r, double s,
                          double* vg_arr) {
    const int type = vg_type_from_name(strip(vg_name ? vg_name : ""));
    if (type < 0) {
        set_error("fill_theo_vg_vals: unknown variogram '%s'", vg_name ? vg_name : "(null)");
        return SPX_EPARSE;
    }
    if (n <= 0 || !(s >= 0) || !(r >= 0)) {  // asserts at pyx:112-115
        set_error("fill_theo_vg_vals: needs n > 0, s >= 0, r >= 0");
        return SPX_EINVAL;
    }
    DevBuf h, o;
    int rc;
    if ((rc = upload(h, h_arr, n * 8)) || (rc = upload(o, vg_arr, n * 8))) return rc;
    k_theo_vg<<<(unsigned)((n + 255) / 256), 256>>>(type, h.as<double>(), n, r, s, o.as<double>());
    SPX_CHECK_LAUNCH("k_theo_vg");
    SPX_CUDA(cudaMemcpy(vg_arr, o.p, n * 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_fill_dists_one_pt(double x, double y, const double* xs, const double* ys, int64_t n,
                          double* dists) {
    if (n < 0) {
        set_error("fill_dists_one_pt: negative size");
        return SPX_EINVAL;
    }
    if (n == 0) return SPX_OK;
    DevBuf a, b, o;
    int rc;
    if ((rc = upload(a, xs, n * 8)) || (rc = upload(b, ys, n * 8)) || (rc = o.alloc(n * 8)))
        return rc;
    k_dists_one_pt<<<(unsigned)((n + 255) / 256), 256>>>(x, y, a.as<double>(), b.as<double>(), n,
                                                         o.as<double>());
    SPX_CHECK_LAUNCH("k_dists_one_pt");
    SPX_CUDA(cudaMemcpy(dists, o.p, n * 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_sel_equidist_refs(double dst_x, double dst_y, const double* ref_xs, const double* ref_ys,
                          int64_t n_refs, uint64_t n_pies, double min_dist_thresh,
                          int64_t not_neb_flag, double* dists, double* tem_ref_sel_dists,
                          int64_t* ref_sel_pie_idxs, uint64_t* ref_pie_idxs,
                          uint64_t* ref_pie_cts) {
    if (n_refs < 0 || n_pies < 1 || n_pies > 1024 || n_refs > 4096) {
        set_error("sel_equidist_refs: n_refs outside 0..4096 or n_pies outside 1..1024");
        return SPX_EINVAL;
    }
    if (n_refs == 0) return SPX_OK;
    (void)tem_ref_sel_dists;   // scratch of the reference's per-sector argsort; not needed
    DevBuf a, b, d, s, p, c;
    int rc;
    if ((rc = upload(a, ref_xs, n_refs * 8)) || (rc = upload(b, ref_ys, n_refs * 8)) ||
        (rc = d.alloc(n_refs * 8)) || (rc = s.alloc(n_refs * 8)) ||
        (rc = upload(p, ref_pie_idxs, n_refs * 8)) || (rc = upload(c, ref_pie_cts, n_pies * 8)))
        return rc;
    k_sel_equidist<<<1, 256, (size_t)n_refs * 12 + 16>>>(
        dst_x, dst_y, a.as<double>(), b.as<double>(), (int)n_refs, (int)n_pies, min_dist_thresh,
        (long long)not_neb_flag, d.as<double>(), s.as<long long>(),
        p.as<unsigned long long>(), c.as<unsigned long long>());
    SPX_CHECK_LAUNCH("k_sel_equidist");
    SPX_CUDA(cudaMemcpy(dists, d.p, n_refs * 8, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(ref_sel_pie_idxs, s.p, n_refs * 8, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(ref_pie_idxs, p.p, n_refs * 8, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(ref_pie_cts, c.p, n_pies * 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_get_nd_dists(const double* pts, int64_t n_pts, int64_t n_dims, double* dists) {
    if (n_pts < 0 || n_dims < 0 || n_pts > 65535) {
        set_error("get_nd_dists: bad sizes (n_pts <= 65535)");
        return SPX_EINVAL;
    }
    const int64_t n_d = n_pts * (n_pts - 1) / 2;
    if (n_d <= 0) return SPX_OK;
    DevBuf a, o;
    int rc;
    if ((rc = upload(a, pts, n_pts * n_dims * 8)) || (rc = o.alloc(n_d * 8))) return rc;
    dim3 grid((unsigned)((n_pts + 127) / 128), (unsigned)n_pts);
    k_nd_dists<<<grid, 128>>>(a.as<double>(), n_pts, n_dims, o.as<double>());
    SPX_CHECK_LAUNCH("k_nd_dists");
    SPX_CUDA(cudaMemcpy(dists, o.p, n_d * 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_fill_wts_and_sum(const double* dists, double* wts, int64_t n, double idw_exp,
                         double* wts_sum) {
    if (n < 0 || !wts_sum) {
        set_error("fill_wts_and_sum: bad argument");
        return SPX_EINVAL;
    }
    if (n == 0) {
        *wts_sum = 0.0;
        return SPX_OK;
    }
    DevBuf d, w, s;
    int rc;
    if ((rc = upload(d, dists, n * 8)) || (rc = w.alloc(n * 8)) || (rc = s.alloc(8))) return rc;
    k_idw_wts<<<(unsigned)((n + 255) / 256), 256>>>(d.as<double>(), w.as<double>(), n, idw_exp);
    k_seq_sum<<<1, 32>>>(w.as<double>(), nullptr, n, s.as<double>());
    SPX_CHECK_LAUNCH("k_idw_wts");
    SPX_CUDA(cudaMemcpy(wts, w.p, n * 8, cudaMemcpyDeviceToHost));
    SPX_CUDA(cudaMemcpy(wts_sum, s.p, 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int spx_get_mults_sum(const double* wts, const double* data, int64_t n, double* mults_sum) {
    if (n < 0 || !mults_sum) {
        set_error("get_mults_sum: bad argument");
        return SPX_EINVAL;
    }
    if (n == 0) {
        *mults_sum = 0.0;
        return SPX_OK;
    }
    DevBuf w, z, s;
    int rc;
    if ((rc = upload(w, wts, n * 8)) || (rc = upload(z, data, n * 8)) || (rc = s.alloc(8)))
        return rc;
    k_seq_sum<<<1, 32>>>(w.as<double>(), z.as<double>(), n, s.as<double>());
    SPX_CHECK_LAUNCH("k_seq_sum");
    SPX_CUDA(cudaMemcpy(mults_sum, s.p, 8, cudaMemcpyDeviceToHost));
    return SPX_OK;
}

}  // extern "C"
