// Supporting kernels of the engine: coefficient packing, nearest-neighbour
// index / gather, constant-row fill and the sum(lambda) ~ 1 check.
#include <type_traits>

#include "spx_common.cuh"
#include "spx_stats.cuh"

namespace spx {

__host__ __device__ __forceinline__ int64_t coef_offset2(int64_t row, int64_t col, int64_t kpad) {
    return ((row / SPX_BM) * (kpad >> 2) + (col >> 2)) * (int64_t)(SPX_BM * 4) +
           ((row % SPX_BM) >> 3) * 32 + (row & 7) * 4 + (col & 3);
}

// One thread per (row, column); consecutive threads walk the columns of a row
// (coalesced reads of the dense source).
__global__ void __launch_bounds__(256) k_pack_rows(const double* __restrict__ src, int64_t src_ld,
                                                   const int32_t* __restrict__ src_rows,
                                                   int64_t n_rows, int n_cols, int kpad,
                                                   int mask_mode, double* __restrict__ coef,
                                                   int64_t row0) {
    const int64_t total = n_rows * n_cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / n_cols;
        const int c = (int)(i - r * n_cols);
        const int64_t sr = src_rows ? (int64_t)src_rows[r] : r;
        double v = src[sr * src_ld + c];
        const bool isn = (v != v);
        if (mask_mode)
            v = isn ? 0.0 : 1.0;
        else if (isn)
            v = 0.0;
        coef[coef_offset2(row0 + r, c, kpad)] = v;
    }
}

// Nearest available station.  Block = 256 cells x one chunk of groups; station
// coordinates and the group's availability bytes are staged in shared memory.
constexpr int NNB_STN_TILE = 1024;

__global__ void __launch_bounds__(256) k_nnb_index(const double* __restrict__ stn_x,
                                                   const double* __restrict__ stn_y, int n_stn,
                                                   const uint8_t* __restrict__ grp_mask,
                                                   int n_grps, const double* __restrict__ cell_x,
                                                   const double* __restrict__ cell_y,
                                                   int64_t n_cells, int32_t* __restrict__ nnb) {
    __shared__ double sx[NNB_STN_TILE];
    __shared__ double sy[NNB_STN_TILE];
    __shared__ uint8_t sm[NNB_STN_TILE];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = (c < n_cells) ? cell_x[c] : 0.0;
    const double y = (c < n_cells) ? cell_y[c] : 0.0;
    for (int gi = blockIdx.y; gi < n_grps; gi += gridDim.y) {
        double best = CUDART_INF;
        int bidx = -1;
        for (int s0 = 0; s0 < n_stn; s0 += NNB_STN_TILE) {
            const int ns = min(NNB_STN_TILE, n_stn - s0);
            __syncthreads();
            for (int k = threadIdx.x; k < ns; k += blockDim.x) {
                sx[k] = stn_x[s0 + k];
                sy[k] = stn_y[s0 + k];
                sm[k] = grp_mask[(int64_t)gi * n_stn + s0 + k];
            }
            __syncthreads();
            for (int k = 0; k < ns; ++k) {
                if (!sm[k]) continue;
                const double d = dist_rn(x, y, sx[k], sy[k]);
                // np.argmin: first minimum; a NaN distance would win in NumPy but
                // coordinates are asserted finite (interp/grps.py:44-45, :264-265)
                if (d < best || bidx < 0) {
                    best = d;
                    bidx = s0 + k;
                }
            }
        }
        if (c < n_cells) nnb[(int64_t)gi * n_cells + c] = bidx;
    }
}

// Candidate lists: the NNB_L stations nearest to every cell among ALL stations,
// ordered by (IEEE distance, station index).  The nearest AVAILABLE station of a
// group is then the first available candidate -- the same station np.argmin
// returns (first minimum) -- and only cells whose NNB_L nearest stations are all
// missing need the full scan.  One distance pass per chunk instead of one per
// availability group.
constexpr int NNB_L = 16;

__global__ void __launch_bounds__(256) k_nnb_candidates(const double* __restrict__ stn_x,
                                                        const double* __restrict__ stn_y,
                                                        int n_stn,
                                                        const double* __restrict__ cell_x,
                                                        const double* __restrict__ cell_y,
                                                        int64_t n_cells,
                                                        int32_t* __restrict__ cand) {
    __shared__ double sx[NNB_STN_TILE];
    __shared__ double sy[NNB_STN_TILE];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = (c < n_cells) ? cell_x[c] : 0.0;
    const double y = (c < n_cells) ? cell_y[c] : 0.0;
    double bd[NNB_L];
    int bi[NNB_L];
#pragma unroll
    for (int j = 0; j < NNB_L; ++j) { bd[j] = CUDART_INF; bi[j] = -1; }
    for (int s0 = 0; s0 < n_stn; s0 += NNB_STN_TILE) {
        const int ns = min(NNB_STN_TILE, n_stn - s0);
        __syncthreads();
        for (int k = threadIdx.x; k < ns; k += blockDim.x) {
            sx[k] = stn_x[s0 + k];
            sy[k] = stn_y[s0 + k];
        }
        __syncthreads();
        for (int k = 0; k < ns; ++k) {
            const double d = dist_rn(x, y, sx[k], sy[k]);
            if (d < bd[NNB_L - 1]) {   // stations arrive in index order: ties stay behind
                bd[NNB_L - 1] = d;
                bi[NNB_L - 1] = s0 + k;
#pragma unroll
                for (int j = NNB_L - 1; j > 0; --j) {
                    if (bd[j] < bd[j - 1]) {
                        const double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                        const int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                    }
                }
            }
        }
    }
    if (c < n_cells) {
#pragma unroll
        for (int j = 0; j < NNB_L; ++j) cand[c * NNB_L + j] = bi[j];
    }
}

__global__ void __launch_bounds__(256) k_nnb_from_candidates(
    const double* __restrict__ stn_x, const double* __restrict__ stn_y, int n_stn,
    const uint8_t* __restrict__ grp_mask, int n_grps, const double* __restrict__ cell_x,
    const double* __restrict__ cell_y, int64_t n_cells, const int32_t* __restrict__ cand,
    int32_t* __restrict__ nnb) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int cd[NNB_L];
#pragma unroll
    for (int j = 0; j < NNB_L; ++j) cd[j] = cand[c * NNB_L + j];
    for (int gi = blockIdx.y; gi < n_grps; gi += gridDim.y) {
        const uint8_t* __restrict__ m = grp_mask + (int64_t)gi * n_stn;
        int found = -1;
#pragma unroll
        for (int j = 0; j < NNB_L; ++j) {
            if (found < 0 && cd[j] >= 0 && m[cd[j]]) found = cd[j];
        }
        if (found < 0) {   // every candidate missing: full scan (rare)
            const double x = cell_x[c], y = cell_y[c];
            double best = CUDART_INF;
            for (int k = 0; k < n_stn; ++k) {
                if (!m[k]) continue;
                const double d = dist_rn(x, y, stn_x[k], stn_y[k]);
                if (d < best || found < 0) { best = d; found = k; }
            }
        }
        nnb[(int64_t)gi * n_cells + c] = found;
    }
}

__global__ void __launch_bounds__(256) k_nnb_gather(
    const double* __restrict__ data, int n_stn, const int32_t* __restrict__ nnb,
    const int32_t* __restrict__ row_step, const int32_t* __restrict__ row_grp,
    const int32_t* __restrict__ row_dst, int64_t n_rows, const uint8_t* __restrict__ fail,
    const int32_t* __restrict__ row_fail, int64_t n_cells, const int32_t* __restrict__ cell_pos,
    void* __restrict__ out, int64_t out_ld, int out_f64, int has_lo, int has_hi, double lo,
    double hi) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int64_t col = cell_pos ? (int64_t)cell_pos[c] : c;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        if (fail != nullptr && !fail[(int64_t)row_fail[r] * n_cells + c]) continue;
        const int st = nnb[(int64_t)row_grp[r] * n_cells + c];
        double v = data[(int64_t)row_step[r] * n_stn + st];
        v = clampd(v, has_lo, has_hi, lo, hi);
        store_out(out, (int64_t)row_dst[r] * out_ld + col, v, out_f64);
    }
}

__global__ void __launch_bounds__(256) k_fill_rows(const double* __restrict__ vals,
                                                   const int32_t* __restrict__ row_dst,
                                                   int64_t n_rows, int64_t n_cells,
                                                   const int32_t* __restrict__ cell_pos,
                                                   void* __restrict__ out, int64_t out_ld,
                                                   int out_f64, int has_lo, int has_hi, double lo,
                                                   double hi) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int64_t col = cell_pos ? (int64_t)cell_pos[c] : c;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const double v = clampd(vals[r], has_lo, has_hi, lo, hi);
        store_out(out, (int64_t)row_dst[r] * out_ld + col, v, out_f64);
    }
}

// out[row_dst[r], cell_pos[c]] = aux[row_slot[r], c]  (aux == NULL: 0.0), only where
// fail[row_fail[r], c] != 0 when a fail mask is given.  No clamp: estimation
// variances are not passed through _mod_min_max (interp/steps.py:805-831).
__global__ void __launch_bounds__(256) k_bcast_rows(
    const double* __restrict__ aux, const int32_t* __restrict__ row_slot,
    const int32_t* __restrict__ row_dst, int64_t n_rows, const uint8_t* __restrict__ fail,
    const int32_t* __restrict__ row_fail, int64_t n_cells, const int32_t* __restrict__ cell_pos,
    void* __restrict__ out, int64_t out_ld, int out_f64) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int64_t col = cell_pos ? (int64_t)cell_pos[c] : c;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        if (fail != nullptr && !fail[(int64_t)row_fail[r] * n_cells + c]) continue;
        const double v = aux ? aux[(int64_t)row_slot[r] * n_cells + c] : 0.0;
        store_out(out, (int64_t)row_dst[r] * out_ld + col, v, out_f64);
    }
}

// np.isclose(a, 1.0): |a - 1| <= atol + rtol * |1| with rtol=1e-5, atol=1e-8;
// NaN / inf are not close (interp/steps.py:418).
__global__ void __launch_bounds__(256) k_lambda_check(const double* __restrict__ aux,
                                                      int64_t n_slots, int64_t n_cells,
                                                      const uint8_t* __restrict__ cell_bad,
                                                      uint8_t* __restrict__ fail) {
    const int64_t total = n_slots * n_cells;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const double a = aux[i];
        bool ok = fabs(a - 1.0) <= (1e-8 + 1e-5);
        if (!(a == a) || isinf(a)) ok = false;
        if (cell_bad != nullptr && cell_bad[i % n_cells]) ok = false;
        fail[i] = ok ? 0 : 1;
    }
}

// Column indices of the set (want = 1) or clear (want = 0) entries of selected rows
// of a [*, n_cols] byte mask, written at host-computed offsets: the station lists of
// the availability groups without a host-side np.where over the whole mask.  One
// warp per selected row, ballot / popc compaction (ascending order).
__global__ void __launch_bounds__(256) k_mask_lists(const uint8_t* __restrict__ mask, int n_cols,
                                                    const int32_t* __restrict__ row_sel,
                                                    int n_sel, const int64_t* __restrict__ off,
                                                    int want, int32_t* __restrict__ out) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_sel) return;
    const uint8_t* __restrict__ m = mask + (int64_t)row_sel[w] * n_cols;
    int32_t* __restrict__ o = out + off[w];
    int base = 0;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int c = c0 + lane;
        const bool hit = (c < n_cols) && ((m[c] != 0) == (want != 0));
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (hit) o[base + __popc(b & ((1u << lane) - 1u))] = c;
        base += __popc(b);
    }
}

// Small device -> (mapped, pinned) host copy done by SMs instead of a DMA engine:
// a cudaMemcpyAsync D2H would queue behind any large field download in flight on
// the copy engine and stall the compute stream for its whole duration.
__global__ void k_copy_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                             int64_t n_words) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride)
        dst[i] = src[i];
}

static inline unsigned grid_y(int64_t n) { return (unsigned)(n < 1 ? 1 : (n > 65535 ? 65535 : n)); }


// ------------------------------------------------------------- rounding + statistics
// interp/steps.py:907-912 rounds every field to nmrl_prcn decimals in the field dtype
// before it is written (np.round: rint(x * 10^d) / 10^d, both operations rounded in that
// dtype) and interp/main.py:474-525 re-reads the file to get per-step min / mean / max /
// std / count.  Both in one pass over the field while it is still in HBM.
constexpr int RS_THREADS = 256;

template <typename T, bool VEC>
__global__ void __launch_bounds__(RS_THREADS) k_round_stats(T* fld, int64_t row_len, int64_t ld,
                                                           int64_t seg_len, int do_round, T pw,
                                                           StatPart* parts) {
    const int64_t row = blockIdx.y;
    const int seg = blockIdx.x;
    T* base = fld + row * ld;
    const int64_t beg = (int64_t)seg * seg_len;
    const int64_t end = min(row_len, beg + seg_len);
    // shift by the first value of the segment (if usable) against cancellation
    double K = (beg < end) ? (double)base[beg] : 0.0;
    if (!(fabs(K) < 1.0e300)) K = 0.0;
    if (do_round && beg < end) {
        const double Kr = (double)round_dec<T>((T)K, pw);
        if (fabs(Kr) < 1.0e300) K = Kr;
    }
    double n = 0.0, s = 0.0, q = 0.0, nfin = 0.0;
    double mn = CUDART_INF, mx = -CUDART_INF;
    auto take = [&](T v) {
        if (v == v) {
            const double d = (double)v - K;
            n += 1.0;
            s += d;
            q = fma(d, d, q);
            mn = fmin(mn, (double)v);
            mx = fmax(mx, (double)v);
            if (fabs((double)v) <= 1.7976931348623157e308) nfin += 1.0;
        }
    };
    if (VEC) {
        constexpr int W = 16 / sizeof(T);
        typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type V;
        constexpr int64_t STEP = (int64_t)RS_THREADS * W;
        int64_t i = beg + (int64_t)threadIdx.x * W;
        // four independent 16-byte loads in flight per thread (one was latency-bound)
        for (; i + 3 * STEP + W <= end; i += 4 * STEP) {
            V v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<V*>(base + i + q * STEP);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                T* e = reinterpret_cast<T*>(&v[q]);
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    if (do_round) e[u] = round_dec<T>(e[u], pw);
                    take(e[u]);
                }
                if (do_round) *reinterpret_cast<V*>(base + i + q * STEP) = v[q];
            }
        }
        for (; i < end; i += STEP) {
            if (i + W <= end) {
                V v = *reinterpret_cast<V*>(base + i);
                T* e = reinterpret_cast<T*>(&v);
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    if (do_round) e[u] = round_dec<T>(e[u], pw);
                    take(e[u]);
                }
                if (do_round) *reinterpret_cast<V*>(base + i) = v;
            } else {
                for (int64_t j = i; j < end; ++j) {
                    T v = base[j];
                    if (do_round) { v = round_dec<T>(v, pw); base[j] = v; }
                    take(v);
                }
            }
        }
    } else {
        for (int64_t i = beg + threadIdx.x; i < end; i += RS_THREADS) {
            T v = base[i];
            if (do_round) { v = round_dec<T>(v, pw); base[i] = v; }
            take(v);
        }
    }
    StatPart a;
    a.n = n;
    a.mean = (n > 0.0) ? K + s / n : 0.0;
    a.m2 = (n > 0.0) ? q - s * (s / n) : 0.0;
    a.mn = mn;
    a.mx = mx;
    a.nfin = nfin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const StatPart b = stat_shfl(a, o);
        stat_merge(a, b);
    }
    __shared__ StatPart sp[RS_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        StatPart t = sp[0];
        for (int w = 1; w < RS_THREADS / 32; ++w) stat_merge(t, sp[w]);
        parts[row * gridDim.x + seg] = t;
    }
}

// one warp per row: lanes merge every 32nd segment, then a shuffle tree (Chan et al.)
__global__ void __launch_bounds__(128) k_stats_final(const StatPart* __restrict__ parts, int n_seg,
                                                     int64_t n_rows, double* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    StatPart t;
    t.n = 0.0; t.mean = 0.0; t.m2 = 0.0; t.mn = CUDART_INF; t.mx = -CUDART_INF; t.nfin = 0.0;
    for (int sgi = lane; sgi < n_seg; sgi += 32) stat_merge(t, parts[row * n_seg + sgi]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const StatPart b = stat_shfl(t, o);
        stat_merge(t, b);
    }
    if (lane != 0) return;
    const bool any = t.n > 0.0;
    stats[0 * n_rows + row] = any ? t.mn : CUDART_NAN;
    stats[1 * n_rows + row] = any ? t.mean : CUDART_NAN;
    stats[2 * n_rows + row] = any ? t.mx : CUDART_NAN;
    stats[3 * n_rows + row] = any ? sqrt(fmax(t.m2, 0.0) / t.n) : CUDART_NAN;
    stats[4 * n_rows + row] = t.nfin;
}

void launch_stats_final(const StatPart* parts, int n_seg, int64_t n_rows, double* stats,
                        cudaStream_t st) {
    k_stats_final<<<(unsigned)((n_rows + 3) / 4), 128, 0, st>>>(parts, n_seg, n_rows, stats);
}

}  // namespace spx

using namespace spx;

extern "C" {

int spx_pack_rows_dev(const double* src, int64_t src_ld, const int32_t* src_rows, int64_t n_rows,
                      int32_t n_cols, int32_t kpad, int mask_mode, double* coef, int64_t row0,
                      void* stream) {
    if (n_cols > kpad || kpad % 4 != 0) {
        set_error("pack_rows: n_cols=%d kpad=%d", n_cols, kpad);
        return SPX_EINVAL;
    }
    const int64_t total = n_rows * n_cols;
    if (total == 0) return SPX_OK;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    k_pack_rows<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, src_ld, src_rows, n_rows, n_cols,
                                                          kpad, mask_mode, coef, row0);
    SPX_CHECK_LAUNCH("k_pack_rows");
    return SPX_OK;
}

int spx_nnb_index_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                      const uint8_t* grp_mask, int32_t n_grps, const double* cell_x,
                      const double* cell_y, int64_t n_cells, int32_t* nnb, void* stream) {
    if (n_grps == 0 || n_cells == 0) return SPX_OK;
    if (n_stn <= 0) {
        set_error("nnb_index: no stations");
        return SPX_EINVAL;
    }
    dim3 grid((unsigned)((n_cells + 255) / 256), grid_y(n_grps));
    k_nnb_index<<<grid, 256, 0, (cudaStream_t)stream>>>(stn_x, stn_y, n_stn, grp_mask, n_grps,
                                                        cell_x, cell_y, n_cells, nnb);
    SPX_CHECK_LAUNCH("k_nnb_index");
    return SPX_OK;
}

int spx_nnb_candidates_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                           const double* cell_x, const double* cell_y, int64_t n_cells,
                           int32_t* cand, void* stream) {
    if (n_cells == 0) return SPX_OK;
    if (n_stn <= 0) {
        set_error("nnb_candidates: no stations");
        return SPX_EINVAL;
    }
    k_nnb_candidates<<<(unsigned)((n_cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        stn_x, stn_y, n_stn, cell_x, cell_y, n_cells, cand);
    SPX_CHECK_LAUNCH("k_nnb_candidates");
    return SPX_OK;
}

int spx_nnb_candidates_width(void) { return NNB_L; }

int spx_nnb_index_cand_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                           const uint8_t* grp_mask, int32_t n_grps, const double* cell_x,
                           const double* cell_y, int64_t n_cells, const int32_t* cand,
                           int32_t* nnb, void* stream) {
    if (n_grps == 0 || n_cells == 0) return SPX_OK;
    dim3 grid((unsigned)((n_cells + 255) / 256), grid_y(n_grps < 64 ? n_grps : 64));
    k_nnb_from_candidates<<<grid, 256, 0, (cudaStream_t)stream>>>(
        stn_x, stn_y, n_stn, grp_mask, n_grps, cell_x, cell_y, n_cells, cand, nnb);
    SPX_CHECK_LAUNCH("k_nnb_from_candidates");
    return SPX_OK;
}

int spx_nnb_gather_dev(const double* data, int32_t n_stn, const int32_t* nnb,
                       const int32_t* row_step, const int32_t* row_grp, const int32_t* row_dst,
                       int64_t n_rows, const uint8_t* fail, const int32_t* row_fail,
                       int64_t n_cells, const int32_t* cell_pos, void* out, int64_t out_ld,
                       int32_t out_f64, int32_t has_lo, int32_t has_hi, double lo, double hi,
                       void* stream) {
    if (n_rows == 0 || n_cells == 0) return SPX_OK;
    if (fail != nullptr && row_fail == nullptr) {
        set_error("nnb_gather: fail without row_fail");
        return SPX_EINVAL;
    }
    dim3 grid((unsigned)((n_cells + 255) / 256), grid_y(n_rows));
    k_nnb_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(data, n_stn, nnb, row_step, row_grp,
                                                         row_dst, n_rows, fail, row_fail, n_cells,
                                                         cell_pos, out, out_ld, out_f64, has_lo,
                                                         has_hi, lo, hi);
    SPX_CHECK_LAUNCH("k_nnb_gather");
    return SPX_OK;
}

int spx_fill_rows_dev(const double* vals, const int32_t* row_dst, int64_t n_rows, int64_t n_cells,
                      const int32_t* cell_pos, void* out, int64_t out_ld, int32_t out_f64,
                      int32_t has_lo, int32_t has_hi, double lo, double hi, void* stream) {
    if (n_rows == 0 || n_cells == 0) return SPX_OK;
    dim3 grid((unsigned)((n_cells + 255) / 256), grid_y(n_rows));
    k_fill_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(vals, row_dst, n_rows, n_cells, cell_pos,
                                                        out, out_ld, out_f64, has_lo, has_hi, lo,
                                                        hi);
    SPX_CHECK_LAUNCH("k_fill_rows");
    return SPX_OK;
}

int spx_mask_lists_dev(const uint8_t* mask, int32_t n_cols, const int32_t* row_sel,
                       int32_t n_sel, const int64_t* off, int32_t want, int32_t* out,
                       void* stream) {
    if (n_sel == 0) return SPX_OK;
    const int64_t threads = (int64_t)n_sel * 32;
    k_mask_lists<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        mask, n_cols, row_sel, n_sel, off, want, out);
    SPX_CHECK_LAUNCH("k_mask_lists");
    return SPX_OK;
}

int spx_bcast_rows_dev(const double* aux, const int32_t* row_slot, const int32_t* row_dst,
                       int64_t n_rows, const uint8_t* fail, const int32_t* row_fail,
                       int64_t n_cells, const int32_t* cell_pos, void* out, int64_t out_ld,
                       int32_t out_f64, void* stream) {
    if (n_rows == 0 || n_cells == 0) return SPX_OK;
    if ((aux != nullptr && row_slot == nullptr) || (fail != nullptr && row_fail == nullptr)) {
        set_error("bcast_rows: aux without row_slot or fail without row_fail");
        return SPX_EINVAL;
    }
    dim3 grid((unsigned)((n_cells + 255) / 256), grid_y(n_rows));
    k_bcast_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(aux, row_slot, row_dst, n_rows, fail,
                                                         row_fail, n_cells, cell_pos, out, out_ld,
                                                         out_f64);
    SPX_CHECK_LAUNCH("k_bcast_rows");
    return SPX_OK;
}

int spx_copy_to_mapped_host_dev(void* dst_host_mapped, const void* src_dev, int64_t n_bytes,
                                void* stream) {
    if (n_bytes == 0) return SPX_OK;
    if (n_bytes % 4 != 0 || (reinterpret_cast<uintptr_t>(dst_host_mapped) & 3) ||
        (reinterpret_cast<uintptr_t>(src_dev) & 3)) {
        set_error("copy_to_mapped_host: size and pointers must be multiples of 4 bytes");
        return SPX_EINVAL;
    }
    const int64_t n_words = n_bytes / 4;
    const int blocks = (int)((n_words + 255) / 256 < 64 ? (n_words + 255) / 256 : 64);
    k_copy_words<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint32_t*>(dst_host_mapped), reinterpret_cast<const uint32_t*>(src_dev),
        n_words);
    SPX_CHECK_LAUNCH("k_copy_words");
    return SPX_OK;
}

int spx_upload_dev(void* dst_dev, const void* src_host, int64_t n_bytes, void* stream) {
    if (n_bytes == 0) return SPX_OK;
    if (!dst_dev || !src_host || n_bytes < 0) {
        set_error("upload: null pointer or negative size");
        return SPX_EINVAL;
    }
    SPX_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)n_bytes, cudaMemcpyHostToDevice,
                             (cudaStream_t)stream));
    return SPX_OK;
}

int spx_lambda_check_dev(const double* aux, int64_t n_slots, int64_t n_cells,
                         const uint8_t* cell_bad, uint8_t* fail, void* stream) {
    const int64_t total = n_slots * n_cells;
    if (total == 0) return SPX_OK;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    k_lambda_check<<<blocks, 256, 0, (cudaStream_t)stream>>>(aux, n_slots, n_cells, cell_bad,
                                                             fail);
    SPX_CHECK_LAUNCH("k_lambda_check");
    return SPX_OK;
}

static int rs_segments(int64_t n_rows, int64_t row_len) {
    // enough blocks to fill the GPU a few times over, at least 4096 elements each
    int64_t seg = (148 * 64 + n_rows - 1) / n_rows;       // many more blocks than slots: no tail
    const int64_t max_seg = (row_len + 4095) / 4096;
    if (seg > max_seg) seg = max_seg;
    if (seg < 1) seg = 1;
    if (seg > 1024) seg = 1024;
    return (int)seg;
}

int64_t spx_round_stats_workspace(int64_t n_rows, int64_t row_len) {
    return (int64_t)sizeof(StatPart) * n_rows * rs_segments(n_rows > 0 ? n_rows : 1, row_len);
}

int spx_round_stats_dev(void* fld, int32_t is_f64, int64_t n_rows, int64_t row_len, int64_t ld,
                        int32_t decimals, double* stats, void* workspace, void* stream) {
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    if (!fld || !stats || !workspace || ld < row_len) {
        set_error("round_stats: null argument or ld < row_len");
        return SPX_EINVAL;
    }
    if (decimals > 15) {
        set_error("round_stats: decimals > 15");
        return SPX_EINVAL;
    }
    if (n_rows > 65535) {
        set_error("round_stats: more than 65535 rows in one call");
        return SPX_EINVAL;
    }
    const int do_round = decimals >= 0;
    double pw = 1.0;
    for (int i = 0; i < decimals; ++i) pw *= 10.0;
    const int n_seg = rs_segments(n_rows, row_len);
    int64_t seg_len = (row_len + n_seg - 1) / n_seg;
    seg_len = (seg_len + 3) / 4 * 4;                      // keeps 16-byte alignment per segment
    const bool vec = (reinterpret_cast<uintptr_t>(fld) & 15) == 0 &&
                     (ld % (is_f64 ? 2 : 4)) == 0;
    dim3 grid((unsigned)n_seg, (unsigned)n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    StatPart* parts = reinterpret_cast<StatPart*>(workspace);
    if (is_f64) {
        if (vec) k_round_stats<double, true><<<grid, RS_THREADS, 0, st>>>(
            (double*)fld, row_len, ld, seg_len, do_round, pw, parts);
        else k_round_stats<double, false><<<grid, RS_THREADS, 0, st>>>(
            (double*)fld, row_len, ld, seg_len, do_round, pw, parts);
    } else {
        if (vec) k_round_stats<float, true><<<grid, RS_THREADS, 0, st>>>(
            (float*)fld, row_len, ld, seg_len, do_round, (float)pw, parts);
        else k_round_stats<float, false><<<grid, RS_THREADS, 0, st>>>(
            (float*)fld, row_len, ld, seg_len, do_round, (float)pw, parts);
    }
    SPX_CHECK_LAUNCH("k_round_stats");
    launch_stats_final(parts, n_seg, n_rows, stats, st);
    SPX_CHECK_LAUNCH("k_stats_final");
    return SPX_OK;
}

}  // extern "C"
