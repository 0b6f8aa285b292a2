// 'nrst' neighbour selection (interp/grps.py:147-166, :103-139) and the
// estimators that go with it: every cell uses only its k nearest available
// stations, cells sharing the same neighbour set share one (k + border) kriging
// system.  Four kernels:
//   k_topk        per cell: the k nearest available stations (IEEE distances,
//                 indices returned ascending like np.sort(np.argsort(d)[:k])) and a
//                 64-bit hash of the index row used to group cells
//   k_nrst_solve  per cell group: assemble the small system from coordinates in
//                 shared memory, LU with partial pivoting, solve every step of the
//                 availability group (dual coefficients) + the ones-vector
//   k_nrst_krige  per cell: right-hand side from coordinates, sum(lambda) test,
//                 NNB fallback, estimate, clamp, store (interp/steps.py:403-435)
//   k_nrst_idw    per cell: IDW over its neighbours (interp/steps.py:293-313)
#include <algorithm>
#include <cstdlib>

#include "spx_common.cuh"

namespace spx {

// Neighbours per cell: the per-thread work arrays are sized at compile time; two variants
// (<= 64: the reference's usual 10 - 50 neighbours; <= 160: as many as the shared-memory
// LU of k_nrst_solve can hold).
constexpr int NRST_KMAX = 160;  // neighbours per cell, largest variant
constexpr int NRST_BMAX = 8;    // border rows / columns (1 + drifts)

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}

template <int KMAX>
__global__ void __launch_bounds__(128) k_topk(const double* __restrict__ stn_x,
                                              const double* __restrict__ stn_y, int n_stn,
                                              const uint8_t* __restrict__ mask,  // [n_stn] or null
                                              const double* __restrict__ cell_x,
                                              const double* __restrict__ cell_y, int64_t n_cells,
                                              int k, int32_t* __restrict__ nb,   // [n_cells, k]
                                              int64_t* __restrict__ hash) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double x = cell_x[c], y = cell_y[c];
    double bd[KMAX];
    int bi[KMAX];
    int cnt = 0;
    for (int s = 0; s < n_stn; ++s) {
        if (mask != nullptr && !mask[s]) continue;
        const double d = dist_rn(x, y, stn_x[s], stn_y[s]);
        if (cnt == k && !(d < bd[k - 1])) continue;   // ties keep the earlier station
        int j = (cnt < k) ? cnt : k - 1;
        while (j > 0 && d < bd[j - 1]) {
            bd[j] = bd[j - 1];
            bi[j] = bi[j - 1];
            --j;
        }
        bd[j] = d;
        bi[j] = s;
        if (cnt < k) ++cnt;
    }
    // indices ascending (np.sort)
    for (int i = 1; i < cnt; ++i) {
        const int v = bi[i];
        int j = i;
        while (j > 0 && bi[j - 1] > v) {
            bi[j] = bi[j - 1];
            --j;
        }
        bi[j] = v;
    }
    uint64_t h = 0x243f6a8885a308d3ull;
    for (int i = 0; i < k; ++i) {
        const int v = (i < cnt) ? bi[i] : -1;
        nb[c * k + i] = v;
        h = mix64(h, (uint64_t)(uint32_t)v);
    }
    hash[c] = (int64_t)(h >> 1);   // non-negative
}

// The same selection with one WARP per cell (default for 'nrst').  The neighbour row is a
// SET -- the k smallest (distance, station index) pairs, written in index order -- so no
// sorted list is needed: the distances of the cell to all stations go to shared memory as
// 64-bit keys (bit pattern of the non-negative IEEE distance: same order), the k-th
// smallest key is found by bisection on the key value with warp-wide counting (leaves as
// soon as a threshold separates exactly k keys: about log2(n_stn) + 1 rounds), and one more
// pass writes the selected stations in index order (ties at the threshold: lowest indices
// first, what the thread-per-cell insertion keeps) and accumulates the row hash.  No
// divergence, no local-memory lists: 1e6 cells x 1,000 stations x 50 neighbours in ~3 ms
// instead of 17 ms; results are identical to k_topk.
constexpr int TOPK_WARPS = 4;

__device__ __forceinline__ int warp_count_le(const unsigned long long* keys, int n_it, int lane,
                                             unsigned long long t) {
    int c = 0;
#pragma unroll 4
    for (int it = 0; it < n_it; ++it) c += (keys[it * 32 + lane] <= t) ? 1 : 0;
    return (int)__reduce_add_sync(0xffffffffu, (unsigned)c);
}

__global__ void __launch_bounds__(TOPK_WARPS * 32) k_topk_warp(
    const double* __restrict__ stn_x, const double* __restrict__ stn_y, int n_stn,
    const uint8_t* __restrict__ mask, const double* __restrict__ cell_x,
    const double* __restrict__ cell_y, int64_t n_cells, int k, int32_t* __restrict__ nb,
    int64_t* __restrict__ hash) {
    extern __shared__ unsigned long long topk_keys[];      // [TOPK_WARPS][n_pad]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_it = (n_stn + 31) >> 5;
    unsigned long long* keys = topk_keys + (size_t)w * n_it * 32;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int64_t c = (int64_t)blockIdx.x * TOPK_WARPS + w; c < n_cells;
         c += (int64_t)gridDim.x * TOPK_WARPS) {
        const double x = cell_x[c], y = cell_y[c];
        unsigned long long kmin = ~0ull, kmax = 0ull;
        int nv = 0;
        for (int it = 0; it < n_it; ++it) {
            const int s = it * 32 + lane;
            unsigned long long key = ~0ull;                 // not a candidate
            if (s < n_stn && (mask == nullptr || mask[s])) {
                key = (unsigned long long)__double_as_longlong(
                    dist_rn(x, y, __ldg(stn_x + s), __ldg(stn_y + s)));
                if (key == ~0ull) key = ~0ull - 1;          // (a NaN pattern: keep it a candidate)
                kmin = min(kmin, key);
                kmax = max(kmax, key);
                ++nv;
            }
            keys[s] = key;
        }
        nv = (int)__reduce_add_sync(0xffffffffu, (unsigned)nv);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        __syncwarp();
        const int kk = min(k, nv);                          // stations to select
        // threshold: take every key < thr and the first n_eq (by index) keys == thr
        unsigned long long thr = 0ull;
        int n_eq = 0;
        if (kk == nv) {
            thr = ~0ull;                                    // every candidate
        } else if (kk > 0) {
            // invariant: count(key <= lo - 1) < kk <= count(key <= hi)
            unsigned long long lo = kmin, hi = kmax;
            bool exact = false;
            while (lo < hi) {
                const unsigned long long mid = lo + ((hi - lo) >> 1);
                const int cnt = warp_count_le(keys, n_it, lane, mid);
                if (cnt == kk) {
                    thr = mid + 1;
                    exact = true;
                    break;
                }
                if (cnt > kk) hi = mid; else lo = mid + 1;
            }
            if (!exact) {
                thr = lo;                                   // smallest key with count(<=) >= kk
                const int below = (lo == 0ull) ? 0 : warp_count_le(keys, n_it, lane, lo - 1);
                n_eq = kk - below;
            }
        }
        int written = 0, eq_seen = 0;
        uint64_t h = 0x243f6a8885a308d3ull;
        int32_t* row = nb + c * k;
        for (int it = 0; it < n_it; ++it) {
            const unsigned long long key = keys[it * 32 + lane];
            const bool eq = (key == thr) && (thr != ~0ull);
            const unsigned m_eq = __ballot_sync(0xffffffffu, eq);
            const bool take = (key < thr) || (eq && eq_seen + __popc(m_eq & lt_mask) < n_eq);
            unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) row[written + __popc(m & lt_mask)] = it * 32 + lane;
            written += __popc(m);
            eq_seen += __popc(m_eq);
            while (m) {                                     // warp-uniform: every lane hashes
                const int b = __ffs(m) - 1;
                m &= m - 1;
                h = mix64(h, (uint64_t)(uint32_t)(it * 32 + b));
            }
        }
        for (int i = written + lane; i < k; i += 32) row[i] = -1;
        for (int i = written; i < k; ++i) h = mix64(h, (uint64_t)(uint32_t)(-1));
        if (lane == 0) hash[c] = (int64_t)(h >> 1);         // non-negative
        __syncwarp();
    }
}

// 'pie' neighbour selection (interp/grps.py:168-247 + cyth/interpmthds.pyx:811-890): the
// stations are binned into n_pies angular sectors around the cell and ranked by distance
// inside their sector; the neighbours are the first k stations in (rank, distance) order,
// i.e. the nearest station of every sector, then the second nearest of every sector, ...
// One thread per cell, one pass over the stations per rank level (k / n_pies levels).
// Sector expression: pie_sector() in spx_common.cuh.
template <int KMAX>
__global__ void __launch_bounds__(128) k_pie_select(const double* __restrict__ stn_x,
                                                    const double* __restrict__ stn_y, int n_stn,
                                                    const uint8_t* __restrict__ mask,
                                                    const double* __restrict__ cell_x,
                                                    const double* __restrict__ cell_y,
                                                    int64_t n_cells, int k, int n_pies,
                                                    int32_t* __restrict__ nb,
                                                    int64_t* __restrict__ hash) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double x = cell_x[c], y = cell_y[c];
    double last_d[KMAX], cand_d[KMAX];      // n_pies <= k <= KMAX
    int last_i[KMAX], cand_i[KMAX];
    int sel[KMAX];
    for (int p = 0; p < n_pies; ++p) { last_d[p] = -1.0; last_i[p] = -1; }
    int cnt = 0;
    while (cnt < k) {
        for (int p = 0; p < n_pies; ++p) { cand_d[p] = CUDART_INF; cand_i[p] = -1; }
        for (int s = 0; s < n_stn; ++s) {
            if (mask != nullptr && !mask[s]) continue;
            const double sx = stn_x[s], sy = stn_y[s];
            const double d = dist_rn(x, y, sx, sy);
            const int p = pie_sector(__dsub_rn(sx, x), __dsub_rn(sy, y), n_pies);
            // strictly after the station taken at the previous level, in (d, index) order
            const bool after = d > last_d[p] || (d == last_d[p] && s > last_i[p]);
            if (after && (cand_i[p] < 0 || d < cand_d[p])) {
                cand_d[p] = d;
                cand_i[p] = s;
            }
        }
        // the level's stations in distance order
        int taken = 0;
        for (;;) {
            int bp = -1;
            for (int p = 0; p < n_pies; ++p)
                if (cand_i[p] >= 0 && (bp < 0 || cand_d[p] < cand_d[bp] ||
                                       (cand_d[p] == cand_d[bp] && cand_i[p] < cand_i[bp])))
                    bp = p;
            if (bp < 0) break;
            if (cnt < k) sel[cnt++] = cand_i[bp];
            last_d[bp] = cand_d[bp];
            last_i[bp] = cand_i[bp];
            cand_i[bp] = -1;
            ++taken;
        }
        if (taken == 0) break;   // fewer than k stations available
    }
    for (int i = 1; i < cnt; ++i) {           // indices ascending (np.sort)
        const int v = sel[i];
        int j = i;
        while (j > 0 && sel[j - 1] > v) { sel[j] = sel[j - 1]; --j; }
        sel[j] = v;
    }
    uint64_t h = 0x243f6a8885a308d3ull;
    for (int i = 0; i < k; ++i) {
        const int v = (i < cnt) ? sel[i] : -1;
        nb[c * k + i] = v;
        h = mix64(h, (uint64_t)(uint32_t)v);
    }
    hash[c] = (int64_t)(h >> 1);
}

struct NrstSolveArgs {
    int n_grp;                    // cell groups (systems)
    int k, n_border, n_drifts, kind;
    const int32_t* nbu;           // [n_grp, k] neighbour station indices, ascending
    const double* stn_x;
    const double* stn_y;
    const double* stn_drift;      // [n_stn, n_drifts]
    VgDev vg;
    double min_vg_val;
    const double* data;           // [n_steps_total, n_stn]
    int n_stn;
    const int32_t* steps;         // [n_t] step indices of this launch
    int n_t;
    double min_var_thr;
    const uint8_t* step_bypass;   // [n_t] 1 = nugget-only variogram -> mean
    double* coef;                 // [n_grp, n_t + 1, m]  (row n_t = ones-vector solution)
    double* ovr;                  // [n_grp, n_t]  NaN = krige, else the value to write
    int32_t* info;                // [n_grp]
    double* inv;                  // optional [u_end - u_beg, m, m]: A^-1 (estimation variance)
    int u_beg;                    // first system of this launch
    int tb;                       // > 0: right-hand sides per batch, one THREAD each; 0: one warp each
    int y_doubles;                // doubles of the right-hand side area behind the matrix
};

__global__ void __launch_bounds__(128) k_nrst_solve(NrstSolveArgs a) {
    extern __shared__ double ssm[];
    const int u = a.u_beg + blockIdx.x;
    const int k = a.k, m = a.k + a.n_border;
    const int ld = m | 1;
    double* S = ssm;                        // [ld * m] column-major
    double* ys = S + (size_t)ld * m;        // [4][ld], or [m][tb | 1] (thread per right-hand side)
    int* pv = reinterpret_cast<int*>(ys + (size_t)a.y_doubles);   // [m]
    int* st = pv + m;                       // [k]
    __shared__ double red_v[4];
    __shared__ int red_i[4];
    __shared__ int s_p;
    __shared__ double s_pv;
    __shared__ int s_info;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_info = 0;
    for (int i = tid; i < k; i += 128) st[i] = a.nbu[(int64_t)u * k + i];
    __syncthreads();
    const int covar = (a.kind == SPX_KRG_SK);
    for (int idx = tid; idx < m * m; idx += 128) {
        const int j = idx / m, i = idx - j * m;
        double v;
        if (i < k && j < k) {
            const double h = dist_rn(a.stn_x[st[i]], a.stn_y[st[i]], a.stn_x[st[j]], a.stn_y[st[j]]);
            v = vg_eval(a.vg, h, covar, a.min_vg_val);
        } else if (i >= k && j >= k) {
            v = 0.0;
        } else {
            const int b = (i >= k) ? (i - k) : (j - k);
            const int s = (i >= k) ? j : i;
            v = (b == 0) ? 1.0 : a.stn_drift[(int64_t)st[s] * a.n_drifts + (b - 1)];
        }
        S[i + (size_t)j * ld] = v;
    }
    __syncthreads();
    // ---- LU, partial pivoting
    for (int c = 0; c < m; ++c) {
        double bv = -1.0;
        int bi = c;
        const double* colc = S + (size_t)c * ld;
        for (int i = c + tid; i < m; i += 128) {
            const double v = fabs(colc[i]);
            if (v > bv) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 4; ++w)
                if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
            s_p = bi;
            s_pv = colc[bi];
            pv[c] = bi;
            if (!(bv > 0.0) && s_info == 0) s_info = c + 1;
        }
        __syncthreads();
        const int p = s_p;
        const double pvv = s_pv;
        if (p != c)
            for (int j = tid; j < m; j += 128) {
                const double t = S[c + (size_t)j * ld];
                S[c + (size_t)j * ld] = S[p + (size_t)j * ld];
                S[p + (size_t)j * ld] = t;
            }
        __syncthreads();
        double* cc = S + (size_t)c * ld;
        if (pvv != 0.0)
            for (int i = c + 1 + tid; i < m; i += 128) cc[i] = cc[i] / pvv;
        __syncthreads();
        // rank-1 update, four columns per warp and pass (loads of all four before the stores:
        // same arithmetic per element, a quarter of the dependent shared-memory round trips)
        for (int j = c + 1 + wid; j < m; j += 16) {
            double* cj[4];
            double uc[4];
            bool on[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int jj = j + 4 * r;
                on[r] = jj < m;
                cj[r] = S + (size_t)min(jj, m - 1) * ld;
                uc[r] = cj[r][c];
            }
            for (int i = c + 1 + lane; i < m; i += 32) {
                const double l = cc[i];
                double v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) v[r] = cj[r][i];
#pragma unroll
                for (int r = 0; r < 4; ++r) v[r] = fma(-l, uc[r], v[r]);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (on[r]) cj[r][i] = v[r];
            }
        }
        __syncthreads();
    }
    if (tid == 0) a.info[u] = s_info;
    // with inv: m more right-hand sides, the unit vectors (columns of A^-1)
    const int n_rhs = a.n_t + 1 + (a.inv ? m : 0);
    if (a.tb > 0) {
        // ---- right-hand sides, one per THREAD: a.tb of them at a time side by side in shared
        // memory (Y[i][q], odd pitch), every thread runs the two substitutions of its own
        // column -- the same operations in the same order as the warp-per-right-hand-side
        // path below, but tb independent dependency chains per block instead of 4, and the
        // matrix entries are warp-wide broadcasts.
        const int tbp = a.tb | 1;
        double* Y = ys;
        for (int q0 = 0; q0 < n_rhs; q0 += a.tb) {
            const int nq = min(a.tb, n_rhs - q0);
            if (tid < nq) {
                const int q = q0 + tid;
                const bool ones = (q == a.n_t);
                const int unit = q - a.n_t - 1;             // >= 0: unit vector e_unit
                double* y = Y + tid;
                if (unit >= 0) {
                    for (int i = 0; i < m; ++i) y[(size_t)i * tbp] = (i == unit) ? 1.0 : 0.0;
                } else if (ones) {
                    for (int i = 0; i < m; ++i) y[(size_t)i * tbp] = (i < k) ? 1.0 : 0.0;
                } else {
                    const double* __restrict__ z = a.data + (int64_t)a.steps[q] * a.n_stn;
                    double zmax = -CUDART_INF, zsum = 0.0;
                    for (int i = 0; i < k; ++i) {
                        const double v = z[st[i]];
                        zmax = fmax(zmax, v);
                        zsum += v;
                        y[(size_t)i * tbp] = v;
                    }
                    for (int i = k; i < m; ++i) y[(size_t)i * tbp] = 0.0;
                    // steps.py:760-765 (all values below the threshold) and :325-331
                    const bool bypass = !(zmax >= a.min_var_thr) || a.step_bypass[q];
                    a.ovr[(int64_t)u * a.n_t + q] = bypass ? (zsum / k) : CUDART_NAN;
                }
                for (int c = 0; c < m; ++c) {
                    const int p = pv[c];
                    if (p != c) {
                        const double t = y[(size_t)c * tbp];
                        y[(size_t)c * tbp] = y[(size_t)p * tbp];
                        y[(size_t)p * tbp] = t;
                    }
                }
                // Both substitutions in dot-product form, four rows at a time in registers:
                // every row still receives its updates in the order of the column sweeps
                // of the warp path (forward: c ascending, backward: c descending, then the
                // division), so the results are bit-identical, but the inner loops hold no
                // shared-memory stores and four independent FMA chains.
                for (int i0 = 1; i0 < m; i0 += 4) {
                    const int i1 = min(i0 + 1, m - 1), i2 = min(i0 + 2, m - 1), i3 = min(i0 + 3, m - 1);
                    double a0 = y[(size_t)i0 * tbp], a1 = y[(size_t)i1 * tbp];
                    double a2 = y[(size_t)i2 * tbp], a3 = y[(size_t)i3 * tbp];
#pragma unroll 2
                    for (int c = 0; c < i0; ++c) {
                        const double yc = y[(size_t)c * tbp];
                        const double* col = S + (size_t)c * ld;
                        a0 = fma(-col[i0], yc, a0);
                        a1 = fma(-col[i1], yc, a1);
                        a2 = fma(-col[i2], yc, a2);
                        a3 = fma(-col[i3], yc, a3);
                    }
                    y[(size_t)i0 * tbp] = a0;
                    if (i0 + 1 < m) {
                        a1 = fma(-S[(i0 + 1) + (size_t)i0 * ld], a0, a1);
                        y[(size_t)(i0 + 1) * tbp] = a1;
                    }
                    if (i0 + 2 < m) {
                        a2 = fma(-S[(i0 + 2) + (size_t)i0 * ld], a0, a2);
                        a2 = fma(-S[(i0 + 2) + (size_t)(i0 + 1) * ld], a1, a2);
                        y[(size_t)(i0 + 2) * tbp] = a2;
                    }
                    if (i0 + 3 < m) {
                        a3 = fma(-S[(i0 + 3) + (size_t)i0 * ld], a0, a3);
                        a3 = fma(-S[(i0 + 3) + (size_t)(i0 + 1) * ld], a1, a3);
                        a3 = fma(-S[(i0 + 3) + (size_t)(i0 + 2) * ld], a2, a3);
                        y[(size_t)(i0 + 3) * tbp] = a3;
                    }
                }
                for (int i0 = m - 1; i0 >= 0; i0 -= 4) {
                    const int i1 = max(i0 - 1, 0), i2 = max(i0 - 2, 0), i3 = max(i0 - 3, 0);
                    double a0 = y[(size_t)i0 * tbp], a1 = y[(size_t)i1 * tbp];
                    double a2 = y[(size_t)i2 * tbp], a3 = y[(size_t)i3 * tbp];
#pragma unroll 2
                    for (int c = m - 1; c > i0; --c) {
                        const double xc = y[(size_t)c * tbp];
                        const double* col = S + (size_t)c * ld;
                        a0 = fma(-col[i0], xc, a0);
                        a1 = fma(-col[i1], xc, a1);
                        a2 = fma(-col[i2], xc, a2);
                        a3 = fma(-col[i3], xc, a3);
                    }
                    const double x0 = a0 / S[i0 + (size_t)i0 * ld];
                    y[(size_t)i0 * tbp] = x0;
                    if (i0 - 1 >= 0) {
                        a1 = fma(-S[(i0 - 1) + (size_t)i0 * ld], x0, a1);
                        const double x1 = a1 / S[(i0 - 1) + (size_t)(i0 - 1) * ld];
                        y[(size_t)(i0 - 1) * tbp] = x1;
                        if (i0 - 2 >= 0) {
                            a2 = fma(-S[(i0 - 2) + (size_t)i0 * ld], x0, a2);
                            a2 = fma(-S[(i0 - 2) + (size_t)(i0 - 1) * ld], x1, a2);
                            const double x2 = a2 / S[(i0 - 2) + (size_t)(i0 - 2) * ld];
                            y[(size_t)(i0 - 2) * tbp] = x2;
                            if (i0 - 3 >= 0) {
                                a3 = fma(-S[(i0 - 3) + (size_t)i0 * ld], x0, a3);
                                a3 = fma(-S[(i0 - 3) + (size_t)(i0 - 1) * ld], x1, a3);
                                a3 = fma(-S[(i0 - 3) + (size_t)(i0 - 2) * ld], x2, a3);
                                y[(size_t)(i0 - 3) * tbp] = a3 / S[(i0 - 3) + (size_t)(i0 - 3) * ld];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nq * m; idx += 128) {       // coalesced rows out
                const int ql = idx / m, i = idx - ql * m;
                const int q = q0 + ql;
                const int unit = q - a.n_t - 1;
                double* dst = (unit >= 0)
                                  ? a.inv + ((int64_t)(u - a.u_beg) * m + unit) * m
                                  : a.coef + ((int64_t)u * (a.n_t + 1) + q) * m;
                dst[i] = Y[(size_t)i * tbp + ql];
            }
            __syncthreads();
        }
        return;
    }
    // ---- right-hand sides: n_t data steps + the ones vector, one per warp
    double* y = ys + (size_t)wid * ld;
    for (int q = wid; q < n_rhs; q += 4) {
        const bool ones = (q == a.n_t);
        const int unit = q - a.n_t - 1;                 // >= 0: unit vector e_unit
        double zmax = -CUDART_INF, zsum = 0.0;
        for (int i = lane; i < m; i += 32) {
            double v = 0.0;
            if (unit >= 0) {
                v = (i == unit) ? 1.0 : 0.0;
            } else if (i < k) {
                v = ones ? 1.0 : a.data[(int64_t)a.steps[q] * a.n_stn + st[i]];
                if (!ones) { zmax = fmax(zmax, v); zsum += v; }
            }
            y[i] = v;
        }
        __syncwarp();
        if (!ones && unit < 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
            }
            // steps.py:760-765 (all values below the threshold) and :325-331
            const bool bypass = !(zmax >= a.min_var_thr) || a.step_bypass[q];
            if (lane == 0)
                a.ovr[(int64_t)u * a.n_t + q] = bypass ? (zsum / k) : CUDART_NAN;
        }
        if (lane == 0)
            for (int c = 0; c < m; ++c) {
                const int p = pv[c];
                if (p != c) { const double t = y[c]; y[c] = y[p]; y[p] = t; }
            }
        __syncwarp();
        for (int c = 0; c < m - 1; ++c) {
            const double xc = y[c];
            const double* cc = S + (size_t)c * ld;
            for (int i = c + 1 + lane; i < m; i += 32) y[i] = fma(-cc[i], xc, y[i]);
            __syncwarp();
        }
        for (int c = m - 1; c >= 0; --c) {
            const double* cc = S + (size_t)c * ld;
            const double xc = y[c] / cc[c];
            __syncwarp();
            if (lane == 0) y[c] = xc;
            for (int i = lane; i < c; i += 32) y[i] = fma(-cc[i], xc, y[i]);
            __syncwarp();
        }
        double* dst = (unit >= 0)
                          ? a.inv + ((int64_t)(u - a.u_beg) * m + unit) * m
                          : a.coef + ((int64_t)u * (a.n_t + 1) + q) * m;
        for (int i = lane; i < m; i += 32) dst[i] = y[i];
        __syncwarp();
    }
}

struct NrstEstArgs {
    int64_t n_cells;
    int k, n_border, n_drifts, kind, n_stn;
    const int32_t* cell_grp;      // [n_cells] -> system
    const int32_t* nbu;           // [n_grp, k]
    const double* stn_x;
    const double* stn_y;
    const double* cell_x;
    const double* cell_y;
    const double* cell_drift;     // [n_drifts, n_cells]
    VgDev vg;
    double min_vg_val;
    const double* data;
    const int32_t* steps;
    int n_t;
    const double* coef;
    const double* ovr;
    const int32_t* info;
    const int32_t* cell_pos;
    void* out;
    int64_t out_ld;
    int out_f64, has_lo, has_hi;
    double lo, hi;
    double idw_exp;
    double min_var_thr;
    const double* inv;            // optional [u_end - u_beg, m, m] (k_nrst_solve)
    void* ev_out;                 // optional estimation-variance field (same layout as out)
    int u_beg, u_end;             // systems of this launch (cells of other systems skip)
};

template <int KMAX>
__global__ void __launch_bounds__(128) k_nrst_krige(NrstEstArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cells) return;
    const int k = a.k, m = a.k + a.n_border;
    const int u = a.cell_grp[c];
    if (u < a.u_beg || u >= a.u_end) return;
    const int32_t* __restrict__ st = a.nbu + (int64_t)u * k;
    const double x = a.cell_x[c], y = a.cell_y[c];
    double rhs[KMAX + NRST_BMAX];
    const int covar = (a.kind == SPX_KRG_SK);
    double dmin = CUDART_INF;
    int nn = st[0];
    for (int j = 0; j < k; ++j) {
        const int s = st[j];
        const double d = dist_rn(x, y, a.stn_x[s], a.stn_y[s]);
        if (d < dmin) { dmin = d; nn = s; }          // np.argmin: first minimum
        rhs[j] = vg_eval(a.vg, d, covar, a.min_vg_val);
    }
    for (int b = 0; b < a.n_border; ++b)
        rhs[k + b] = (b == 0) ? 1.0 : a.cell_drift[(int64_t)(b - 1) * a.n_cells + c];
    // sum(lambda) = s . rhs with s = A^-1 [1; 0]  (steps.py:418)
    const double* __restrict__ cu = a.coef + (int64_t)u * (a.n_t + 1) * m;
    double lsum = 0.0;
    {
        const double* sv = cu + (int64_t)a.n_t * m;
        for (int j = 0; j < m; ++j) lsum = fma(sv[j], rhs[j], lsum);
    }
    bool ok = fabs(lsum - 1.0) <= (1e-8 + 1e-5);
    if (!(lsum == lsum) || isinf(lsum) || a.info[u] != 0) ok = false;
    const int64_t col = a.cell_pos ? (int64_t)a.cell_pos[c] : c;
    // estimation variance (steps.py:431-434): sum(lambda * rhs) + lambda[n] with
    // lambda = A^-1 rhs; the same for every step of the system
    double est_var = 0.0;
    if (a.ev_out != nullptr && ok) {
        const double* __restrict__ iv = a.inv + (int64_t)(u - a.u_beg) * m * m;
        double lam_n = 0.0;
        for (int i = 0; i < m; ++i) {
            const double* __restrict__ ci = iv + (int64_t)i * m;    // column i = row i (symmetric)
            double li = 0.0;
            for (int j = 0; j < m; ++j) li = fma(ci[j], rhs[j], li);
            est_var = fma(li, rhs[i], est_var);
            if (i == k) lam_n = li;
        }
        est_var += lam_n;
    }
    for (int q = 0; q < a.n_t; ++q) {
        const int t = a.steps[q];
        const double ov = a.ovr[(int64_t)u * a.n_t + q];
        if (a.ev_out != nullptr)      // steps.py:329 (mean steps), :425-426 (NNB): 0
            store_out(a.ev_out, (int64_t)t * a.out_ld + col, (ov == ov || !ok) ? 0.0 : est_var,
                      a.out_f64);
        double v;
        if (ov == ov) {
            v = ov;
        } else if (!ok) {
            v = a.data[(int64_t)t * a.n_stn + nn];
        } else {
            const double* cq = cu + (int64_t)q * m;
            v = 0.0;
            for (int j = 0; j < m; ++j) v = fma(cq[j], rhs[j], v);
        }
        v = clampd(v, a.has_lo, a.has_hi, a.lo, a.hi);
        store_out(a.out, (int64_t)t * a.out_ld + col, v, a.out_f64);
    }
}

// IDW over the cell's neighbours; nb = the per-cell neighbour rows of k_topk.
template <int KMAX>
__global__ void __launch_bounds__(128) k_nrst_idw(NrstEstArgs a, const int32_t* __restrict__ nb) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cells) return;
    const int k = a.k;
    const int32_t* __restrict__ st = nb + c * k;
    const double x = a.cell_x[c], y = a.cell_y[c];
    double w[KMAX];
    double dmax = 0.0;
    for (int j = 0; j < k; ++j) {
        w[j] = dist_rn(x, y, a.stn_x[st[j]], a.stn_y[st[j]]);
        dmax = fmax(dmax, w[j]);
    }
    double wsum = 0.0;
    for (int j = 0; j < k; ++j) {
        double d = w[j];
        if (dmax > 0) d = d / dmax;              // steps.py:297-301
        w[j] = 1.0 / pow(d, a.idw_exp);          // pyx:792
        wsum += w[j];
    }
    const int64_t col = a.cell_pos ? (int64_t)a.cell_pos[c] : c;
    for (int q = 0; q < a.n_t; ++q) {
        const int t = a.steps[q];
        const double* __restrict__ z = a.data + (int64_t)t * a.n_stn;
        double acc = 0.0, zmax = -CUDART_INF, zsum = 0.0;
        for (int j = 0; j < k; ++j) {
            const double zj = z[st[j]];
            acc = fma(w[j], zj, acc);
            zmax = fmax(zmax, zj);
            zsum += zj;
        }
        double v = (zmax >= a.min_var_thr) ? (acc / wsum) : (zsum / k);   // steps.py:308-313
        v = clampd(v, a.has_lo, a.has_hi, a.lo, a.hi);
        store_out(a.out, (int64_t)t * a.out_ld + col, v, a.out_f64);
    }
}

}  // namespace spx

using namespace spx;

extern "C" {

int spx_nrst_max_neighbors(void) { return NRST_KMAX; }

static int g_topk_warp = -1;        // -1: environment SPX_TOPK_WARP (default 1)
static int g_nrst_thread_rhs = -1;  // -1: environment SPX_NRST_THREAD_RHS (default 1)

int spx_nrst_set_thread_rhs(int on) {
    const int prev = g_nrst_thread_rhs;
    g_nrst_thread_rhs = on;
    return prev;
}

int spx_nrst_set_topk_warp(int on) {
    const int prev = g_topk_warp;
    g_topk_warp = on;
    return prev;
}

int spx_nrst_topk_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                      const uint8_t* mask, const double* cell_x, const double* cell_y,
                      int64_t n_cells, int32_t k, int32_t* nb, int64_t* hash, void* stream) {
    if (n_cells == 0) return SPX_OK;
    if (k < 1 || k > NRST_KMAX) {
        set_error("nrst_topk: k=%d outside 1..%d", k, NRST_KMAX);
        return SPX_EINVAL;
    }
    // one warp per cell with the distances in shared memory (default); the thread-per-cell
    // insertion kernel when the keys of TOPK_WARPS cells do not fit, or by the knob
    static const int warp_env = getenv("SPX_TOPK_WARP") ? atoi(getenv("SPX_TOPK_WARP")) : 1;
    const int warp_knob = g_topk_warp < 0 ? warp_env : g_topk_warp;
    const size_t key_bytes = (size_t)TOPK_WARPS * (size_t)((n_stn + 31) / 32 * 32) * 8;
    if (warp_knob && key_bytes <= 200u * 1024u) {
        static size_t attr_bytes = 48u * 1024u;
        if (key_bytes > attr_bytes) {
            SPX_CUDA(cudaFuncSetAttribute(k_topk_warp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)key_bytes));
            attr_bytes = key_bytes;
        }
        const int64_t want = (n_cells + TOPK_WARPS - 1) / TOPK_WARPS;
        const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)148 * 16);
        k_topk_warp<<<grid, TOPK_WARPS * 32, key_bytes, (cudaStream_t)stream>>>(
            stn_x, stn_y, n_stn, mask, cell_x, cell_y, n_cells, k, nb, hash);
        SPX_CHECK_LAUNCH("k_topk_warp");
        return SPX_OK;
    }
    const unsigned nblk = (unsigned)((n_cells + 127) / 128);
    if (k <= 64)
        k_topk<64><<<nblk, 128, 0, (cudaStream_t)stream>>>(stn_x, stn_y, n_stn, mask, cell_x,
                                                          cell_y, n_cells, k, nb, hash);
    else
        k_topk<NRST_KMAX><<<nblk, 128, 0, (cudaStream_t)stream>>>(stn_x, stn_y, n_stn, mask, cell_x,
                                                                 cell_y, n_cells, k, nb, hash);
    SPX_CHECK_LAUNCH("k_topk");
    return SPX_OK;
}

int spx_pie_select_dev(const double* stn_x, const double* stn_y, int32_t n_stn,
                       const uint8_t* mask, const double* cell_x, const double* cell_y,
                       int64_t n_cells, int32_t k, int32_t n_pies, int32_t* nb, int64_t* hash,
                       void* stream) {
    if (n_cells == 0) return SPX_OK;
    if (k < 1 || k > NRST_KMAX || n_pies < 1 || n_pies > k) {
        set_error("pie_select: k=%d outside 1..%d or n_pies=%d outside 1..k", k, NRST_KMAX,
                  n_pies);
        return SPX_EINVAL;
    }
    const unsigned nblk = (unsigned)((n_cells + 127) / 128);
    if (k <= 64)
        k_pie_select<64><<<nblk, 128, 0, (cudaStream_t)stream>>>(
            stn_x, stn_y, n_stn, mask, cell_x, cell_y, n_cells, k, n_pies, nb, hash);
    else
        k_pie_select<NRST_KMAX><<<nblk, 128, 0, (cudaStream_t)stream>>>(
            stn_x, stn_y, n_stn, mask, cell_x, cell_y, n_cells, k, n_pies, nb, hash);
    SPX_CHECK_LAUNCH("k_pie_select");
    return SPX_OK;
}

int spx_nrst_solve_dev(const spx_nrst* n, void* stream) {
    if (!n) {
        set_error("nrst_solve: null argument");
        return SPX_EINVAL;
    }
    if (n->n_grp == 0) return SPX_OK;
    const int m = n->k + n->n_border;
    if (n->k < 1 || n->k > NRST_KMAX || n->n_border > NRST_BMAX) {
        set_error("nrst_solve: k=%d / border=%d outside the supported range (%d / %d)", n->k,
                  n->n_border, NRST_KMAX, NRST_BMAX);
        return SPX_EINVAL;
    }
    const int u_beg = n->u_beg, u_end = (n->u_end > 0) ? n->u_end : n->n_grp;
    if (u_beg < 0 || u_end > n->n_grp || u_beg >= u_end) {
        set_error("nrst_solve: bad system range");
        return SPX_EINVAL;
    }
    NrstSolveArgs a;
    a.n_grp = n->n_grp;
    a.k = n->k;
    a.n_border = n->n_border;
    a.n_drifts = n->n_drifts;
    a.kind = n->kind;
    a.nbu = n->nbu;
    a.stn_x = n->stn_x;
    a.stn_y = n->stn_y;
    a.stn_drift = n->stn_drift;
    a.vg = to_dev(n->vg);
    a.min_vg_val = n->min_vg_val;
    a.data = n->data;
    a.n_stn = n->n_stn;
    a.steps = n->steps;
    a.n_t = n->n_t;
    a.min_var_thr = n->min_var_thr;
    a.step_bypass = n->step_bypass;
    a.coef = n->coef;
    a.ovr = n->ovr;
    a.info = n->info;
    a.inv = n->inv;
    a.u_beg = u_beg;
    const int ld = m | 1;
    // right-hand sides per batch with one thread each: balanced batches of at most 128, as
    // long as the block stays below ~74 KB (three blocks per SM); larger systems keep the
    // warp-per-right-hand-side substitution, which needs four vectors only
    static const int rhs_env = getenv("SPX_NRST_THREAD_RHS") ? atoi(getenv("SPX_NRST_THREAD_RHS")) : 1;
    const int rhs_knob = g_nrst_thread_rhs < 0 ? rhs_env : g_nrst_thread_rhs;
    const int n_rhs = n->n_t + 1 + (n->inv ? m : 0);
    a.tb = 0;
    a.y_doubles = 4 * ld;
    if (rhs_knob) {
        const size_t budget = 74u * 1024u;
        const size_t fixed = (size_t)ld * m * sizeof(double) + ((size_t)m + n->k) * sizeof(int);
        int tb_max = 0;
        if (budget > fixed) tb_max = (int)((budget - fixed) / ((size_t)m * sizeof(double))) - 1;
        tb_max = std::min(tb_max, 128);
        if (tb_max >= 32) {
            const int n_batch = (n_rhs + tb_max - 1) / tb_max;
            a.tb = (n_rhs + n_batch - 1) / n_batch;
            a.y_doubles = std::max(4 * ld, m * (a.tb | 1));
        }
    }
    const size_t smem = ((size_t)ld * m + (size_t)a.y_doubles) * sizeof(double) +
                        ((size_t)m + n->k) * sizeof(int);
    if (smem > 220 * 1024) {
        set_error("nrst_solve: a system of %d unknowns does not fit shared memory", m);
        return SPX_ENOMEM;
    }
    if (smem > 48 * 1024)
        SPX_CUDA(cudaFuncSetAttribute(k_nrst_solve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    k_nrst_solve<<<(unsigned)(u_end - u_beg), 128, smem, (cudaStream_t)stream>>>(a);
    SPX_CHECK_LAUNCH("k_nrst_solve");
    return SPX_OK;
}

static void fill_est(NrstEstArgs& a, const spx_nrst* n) {
    a.n_cells = n->n_cells;
    a.k = n->k;
    a.n_border = n->n_border;
    a.n_drifts = n->n_drifts;
    a.kind = n->kind;
    a.n_stn = n->n_stn;
    a.cell_grp = n->cell_grp;
    a.nbu = n->nbu;
    a.stn_x = n->stn_x;
    a.stn_y = n->stn_y;
    a.cell_x = n->cell_x;
    a.cell_y = n->cell_y;
    a.cell_drift = n->cell_drift;
    a.vg = to_dev(n->vg);
    a.min_vg_val = n->min_vg_val;
    a.data = n->data;
    a.steps = n->steps;
    a.n_t = n->n_t;
    a.coef = n->coef;
    a.ovr = n->ovr;
    a.info = n->info;
    a.cell_pos = n->cell_pos;
    a.out = n->out;
    a.out_ld = n->out_ld;
    a.out_f64 = n->out_f64;
    a.has_lo = n->has_lo;
    a.has_hi = n->has_hi;
    a.lo = n->lo;
    a.hi = n->hi;
    a.idw_exp = n->idw_exp;
    a.min_var_thr = n->min_var_thr;
    a.inv = n->inv;
    a.ev_out = n->ev_out;
    a.u_beg = n->u_beg;
    a.u_end = (n->u_end > 0) ? n->u_end : n->n_grp;
}

int spx_nrst_krige_dev(const spx_nrst* n, void* stream) {
    if (!n) {
        set_error("nrst_krige: null argument");
        return SPX_EINVAL;
    }
    if (n->n_cells == 0 || n->n_t == 0) return SPX_OK;
    if (n->k > NRST_KMAX || n->n_border > NRST_BMAX || (n->ev_out && !n->inv)) {
        set_error("nrst_krige: system too large, or ev_out without inv");
        return SPX_EINVAL;
    }
    NrstEstArgs a;
    fill_est(a, n);
    const unsigned nblk = (unsigned)((n->n_cells + 127) / 128);
    if (n->k <= 64) k_nrst_krige<64><<<nblk, 128, 0, (cudaStream_t)stream>>>(a);
    else k_nrst_krige<NRST_KMAX><<<nblk, 128, 0, (cudaStream_t)stream>>>(a);
    SPX_CHECK_LAUNCH("k_nrst_krige");
    return SPX_OK;
}

int spx_nrst_idw_dev(const spx_nrst* n, const int32_t* nb, void* stream) {
    if (!n || !nb) {
        set_error("nrst_idw: null argument");
        return SPX_EINVAL;
    }
    if (n->n_cells == 0 || n->n_t == 0) return SPX_OK;
    if (n->k > NRST_KMAX) {
        set_error("nrst_idw: k too large");
        return SPX_EINVAL;
    }
    NrstEstArgs a;
    fill_est(a, n);
    const unsigned nblk = (unsigned)((n->n_cells + 127) / 128);
    if (n->k <= 64) k_nrst_idw<64><<<nblk, 128, 0, (cudaStream_t)stream>>>(a, nb);
    else k_nrst_idw<NRST_KMAX><<<nblk, 128, 0, (cudaStream_t)stream>>>(a, nb);
    SPX_CHECK_LAUNCH("k_nrst_idw");
    return SPX_OK;
}

}  // extern "C"

// ===========================================================================
// Estimate with ONE VARIOGRAM PER ROW (per-step variogram series, config 3):
//     Z[row, cell] = sum_k coef[row, k] * vg_row(dist(station k, cell))  (+ border)
// The contraction kernel of spx_gemm.cu regenerates its right-hand-side tile for
// every variogram, i.e. one sqrt + evaluation per (row, station, cell).  Here the
// DISTANCES of a cell tile are computed once into shared memory ([k][cell], 64
// cells) and every row only re-evaluates its variogram on them: 4 threads per
// cell split the stations, 4 rows are processed per pass so that each distance
// read feeds 4 evaluations.  FP64 ALU bound (no tensor-core shape: the operand
// changes with the row).
namespace spx {

constexpr int MV_CELLS = 64;
constexpr int MV_KQ = 4;       // threads per cell
constexpr int MV_ROWS = 4;     // rows per pass

// A variogram "compiled" for the inner loop: no per-term type dispatch and no
// divisions.  Two closed forms cover the common families:
//   polynomial  s * (hc * ca - hc^3 * cb), hc = min(h, r)   Sph (cb = 1/(2 r^3)) and
//                                                           Lin (cb = 0); at h >= r it
//                                                           gives s * (1.5 - 0.5) = s
//   exponential s * (1 - exp(ca * x)), x = h or h^2         Exp / Gau
// Nug terms are summed.  At most two terms of each form; anything else (Pow, Hol,
// Rng, longer nests) takes the generic vg_eval().
struct RowP {
    double nug, tot;
    double ps[2], pr[2], pca[2], pcb[2];
    double es[2], eca[2];
    int esq[2];
    int np, ne, ok;
};

__device__ __forceinline__ void build_row_p(RowP& o, const spx_vg& v) {
    o.nug = 0.0;
    o.tot = 0.0;
    o.np = o.ne = 0;
    o.ok = 1;
    for (int i = 0; i < 2; ++i) {
        o.ps[i] = o.pca[i] = o.pcb[i] = o.es[i] = o.eca[i] = 0.0;
        o.pr[i] = 1.0;
        o.esq[i] = 0;
    }
    for (int t = 0; t < v.n_terms; ++t) {
        const int ty = v.types[t];
        const double sl = v.sills[t], r = v.ranges[t];
        o.tot += sl;
        if (ty == SPX_VG_NUG) {
            o.nug += sl;
        } else if (ty == SPX_VG_SPH || ty == SPX_VG_LIN) {
            if (o.np >= 2) { o.ok = 0; continue; }
            const int i = o.np++;
            o.ps[i] = sl;
            o.pr[i] = r;
            o.pca[i] = (ty == SPX_VG_SPH) ? 1.5 / r : 1.0 / r;
            o.pcb[i] = (ty == SPX_VG_SPH) ? 1.0 / (2 * (r * r * r)) : 0.0;
        } else if (ty == SPX_VG_EXP || ty == SPX_VG_GAU) {
            if (o.ne >= 2) { o.ok = 0; continue; }
            const int i = o.ne++;
            o.es[i] = sl;
            o.eca[i] = (ty == SPX_VG_EXP) ? -3.0 / r : -3.0 / (r * r);
            o.esq[i] = (ty == SPX_VG_GAU);
        } else {
            o.ok = 0;
        }
    }
}

__device__ __forceinline__ double eval_row_p(const RowP& o, double h, int covar_flag,
                                             double min_vg_val) {
    double g = o.nug;
    if (o.np > 0) {
        const double hc = fmin(h, o.pr[0]);
        g = fma(o.ps[0], hc * o.pca[0] - hc * hc * hc * o.pcb[0], g);
        if (o.np > 1) {
            const double hd = fmin(h, o.pr[1]);
            g = fma(o.ps[1], hd * o.pca[1] - hd * hd * hd * o.pcb[1], g);
        }
    }
    if (o.ne > 0) {
        g = fma(o.es[0], 1.0 - exp(o.eca[0] * (o.esq[0] ? h * h : h)), g);
        if (o.ne > 1) g = fma(o.es[1], 1.0 - exp(o.eca[1] * (o.esq[1] ? h * h : h)), g);
    }
    if (covar_flag) g = o.tot - g;
    return (g <= min_vg_val) ? 0.0 : g;
}

struct MultiVgArgs {
    const double* coef;        // [n_rows, kpad] row-major
    int64_t n_rows;
    int kpad, n_stn, n_border;
    const double* stn_x;
    const double* stn_y;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    const double* cell_drift;
    const spx_vg* vgs;         // device table
    const int32_t* row_vg;     // [n_rows]
    int covar_flag;
    double min_vg_val;
    const int32_t* row_dst;
    void* out;
    int64_t out_ld;
    int out_f64;
    const int32_t* cell_pos;
    int has_lo, has_hi;
    double lo, hi;
};

template <bool FAST>
__global__ void __launch_bounds__(MV_CELLS * MV_KQ) k_estimate_multivg(MultiVgArgs a) {
    extern __shared__ double msm[];
    const int kp = a.n_stn + a.n_border;                 // used K
    double* D = msm;                                     // [kp][MV_CELLS] distances / border values
    double* Cs = D + (size_t)kp * MV_CELLS;              // [MV_ROWS][kp] coefficient rows
    double* Ps = Cs + (size_t)MV_ROWS * kp;              // [MV_ROWS][MV_KQ][MV_CELLS] partial sums
    __shared__ RowP rv[MV_ROWS];
    __shared__ VgDev vd[MV_ROWS];
    const int tid = threadIdx.x;
    const int cl = tid % MV_CELLS, kq = tid / MV_CELLS;
    const int64_t n_tiles = (a.n_cells + MV_CELLS - 1) / MV_CELLS;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t c = tile * MV_CELLS + cl;
        const bool cell_ok = c < a.n_cells;
        const double x = cell_ok ? a.cell_x[c] : 0.0, y = cell_ok ? a.cell_y[c] : 0.0;
        __syncthreads();
        for (int k = kq; k < kp; k += MV_KQ) {
            double v;
            if (k < a.n_stn) {
                const double dx = x - a.stn_x[k], dy = y - a.stn_y[k];
                v = sqrt(dx * dx + dy * dy);
            } else {
                const int b = k - a.n_stn;
                v = (b == 0) ? 1.0 : (cell_ok ? a.cell_drift[(int64_t)(b - 1) * a.n_cells + c] : 0.0);
            }
            D[(size_t)k * MV_CELLS + cl] = v;
        }
        for (int64_t r0 = 0; r0 < a.n_rows; r0 += MV_ROWS) {
            const int nr = (int)min((int64_t)MV_ROWS, a.n_rows - r0);
            __syncthreads();   // D ready / previous pass consumed
            for (int idx = tid; idx < nr * kp; idx += MV_CELLS * MV_KQ) {
                const int rr = idx / kp, k = idx - rr * kp;
                Cs[(size_t)rr * kp + k] = a.coef[(r0 + rr) * (int64_t)a.kpad + k];
            }
            if (tid < nr) {
                const spx_vg& v = a.vgs[a.row_vg[r0 + tid]];
                build_row_p(rv[tid], v);
                vd[tid].n_terms = v.n_terms;
                for (int t = 0; t < SPX_VG_MAX_TERMS; ++t) {
                    vd[tid].types[t] = v.types[t];
                    vd[tid].sills[t] = v.sills[t];
                    vd[tid].ranges[t] = v.ranges[t];
                }
            }
            __syncthreads();
            double acc[MV_ROWS];
#pragma unroll
            for (int rr = 0; rr < MV_ROWS; ++rr) acc[rr] = 0.0;
            // rows outermost: the row's compiled variogram lives in registers for the
            // whole station loop (two partial sums for instruction-level parallelism)
#pragma unroll 1
            for (int rr = 0; rr < nr; ++rr) {
                const RowP p = rv[rr];
                const double* __restrict__ crow = Cs + (size_t)rr * kp;
                double a0 = 0.0, a1 = 0.0;
                if (FAST || p.ok) {
                    int k = kq;
                    for (; k + MV_KQ < a.n_stn; k += 2 * MV_KQ) {
                        const double h0 = D[(size_t)k * MV_CELLS + cl];
                        const double h1 = D[(size_t)(k + MV_KQ) * MV_CELLS + cl];
                        a0 = fma(crow[k], eval_row_p(p, h0, a.covar_flag, a.min_vg_val), a0);
                        a1 = fma(crow[k + MV_KQ], eval_row_p(p, h1, a.covar_flag, a.min_vg_val), a1);
                    }
                    for (; k < a.n_stn; k += MV_KQ)
                        a0 = fma(crow[k], eval_row_p(p, D[(size_t)k * MV_CELLS + cl], a.covar_flag,
                                                     a.min_vg_val), a0);
                } else {
                    for (int k = kq; k < a.n_stn; k += MV_KQ)
                        a0 = fma(crow[k], vg_eval(vd[rr], D[(size_t)k * MV_CELLS + cl],
                                                  a.covar_flag, a.min_vg_val), a0);
                }
                for (int k = a.n_stn + kq; k < kp; k += MV_KQ)
                    a0 = fma(crow[k], D[(size_t)k * MV_CELLS + cl], a0);
                acc[rr] = a0 + a1;
            }
#pragma unroll
            for (int rr = 0; rr < MV_ROWS; ++rr)
                Ps[((size_t)rr * MV_KQ + kq) * MV_CELLS + cl] = acc[rr];
            __syncthreads();
            // MV_ROWS x MV_CELLS results, one per thread
            {
                const int rr = tid / MV_CELLS;   // MV_ROWS == MV_KQ
                if (rr < nr && cell_ok) {
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < MV_KQ; ++q) v += Ps[((size_t)rr * MV_KQ + q) * MV_CELLS + cl];
                    const int dst = a.row_dst[r0 + rr];
                    if (dst >= 0) {
                        v = clampd(v, a.has_lo, a.has_hi, a.lo, a.hi);
                        const int64_t col = a.cell_pos ? (int64_t)a.cell_pos[c] : c;
                        store_out(a.out, (int64_t)dst * a.out_ld + col, v, a.out_f64);
                    }
                }
            }
        }
    }
}

}  // namespace spx

extern "C" int spx_estimate_multivg_dev(const spx_multivg* g, void* stream) {
    using namespace spx;
    if (!g) {
        set_error("estimate_multivg: null argument");
        return SPX_EINVAL;
    }
    if (g->n_rows == 0 || g->n_cells == 0) return SPX_OK;
    if (g->kpad < g->n_stn + g->n_border || (g->n_border > 1 && !g->cell_drift)) {
        set_error("estimate_multivg: bad kpad / drift");
        return SPX_EINVAL;
    }
    MultiVgArgs a;
    a.coef = g->coef;
    a.n_rows = g->n_rows;
    a.kpad = g->kpad;
    a.n_stn = g->n_stn;
    a.n_border = g->n_border;
    a.stn_x = g->stn_x;
    a.stn_y = g->stn_y;
    a.cell_x = g->cell_x;
    a.cell_y = g->cell_y;
    a.n_cells = g->n_cells;
    a.cell_drift = g->cell_drift;
    a.vgs = g->vgs;
    a.row_vg = g->row_vg;
    a.covar_flag = g->covar_flag;
    a.min_vg_val = g->min_vg_val;
    a.row_dst = g->row_dst;
    a.out = g->out;
    a.out_ld = g->out_ld;
    a.out_f64 = g->out_f64;
    a.cell_pos = g->cell_pos;
    a.has_lo = g->has_lo;
    a.has_hi = g->has_hi;
    a.lo = g->lo;
    a.hi = g->hi;
    const int kp = g->n_stn + g->n_border;
    const size_t smem = ((size_t)kp * MV_CELLS + (size_t)MV_ROWS * kp +
                         (size_t)MV_ROWS * MV_KQ * MV_CELLS) * sizeof(double);
    int dev = 0, max_smem = 0, n_sm = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SPX_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (smem + 4096 > (size_t)max_smem) {
        set_error("estimate_multivg: %d stations do not fit in shared memory", g->n_stn);
        return SPX_ENOMEM;
    }
    SPX_CUDA(cudaFuncSetAttribute(k_estimate_multivg<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPX_CUDA(cudaFuncSetAttribute(k_estimate_multivg<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (g->n_cells + MV_CELLS - 1) / MV_CELLS;
    const int per_sm = (int)((size_t)max_smem / (smem + 4096));
    const int64_t want = (int64_t)n_sm * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    const int grid = (int)(tiles < want ? tiles : want);
    if (g->all_fast)
        k_estimate_multivg<true><<<grid, MV_CELLS * MV_KQ, smem, (cudaStream_t)stream>>>(a);
    else
        k_estimate_multivg<false><<<grid, MV_CELLS * MV_KQ, smem, (cudaStream_t)stream>>>(a);
    SPX_CHECK_LAUNCH("k_estimate_multivg");
    return SPX_OK;
}

// ===========================================================================
// LOCAL estimator for compactly supported variograms (Nug + Sph / Lin terms).
// Beyond the largest range R every entry of the cell<->station variogram matrix
// equals the constant F = sum(sills) (0 for the covariance form), exactly as in
// the reference (pyx:50-51, :63-64).  Hence
//   Z[row, cell] = F * sum_k coef[row, k] + border terms
//                  + sum_{k : dist(k, cell) < R} coef[row, k] * (vg(dist) - F)
// and only the stations within R of a cell (found through a uniform bin grid of
// size R) contribute individually.  Work per cell-step drops from 2 (N + k) flop
// to ~2 per near station; the kernel is bound by the HBM write of the field.
namespace spx {

struct LocalBuildArgs {
    const double* stn_x;
    const double* stn_y;
    const int32_t* bin_start;   // [nbx * nby + 1]
    const int32_t* bin_stn;     // station ids ordered by bin
    double x0, y0, inv_bin;
    int nbx, nby;
    double R, F;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    int cap;
    int32_t* cnt;               // [n_cells]
    int32_t* idx;               // [n_cells, cap]
    double* val;                // [n_cells, cap]
    VgDev vg;
    int covar_flag;
    double min_vg_val;
};

__global__ void __launch_bounds__(256) k_local_build(LocalBuildArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cells) return;
    const double x = a.cell_x[c], y = a.cell_y[c];
    const int bx = (int)floor((x - a.x0) * a.inv_bin);
    const int by = (int)floor((y - a.y0) * a.inv_bin);
    int n = 0;
    for (int jy = by - 1; jy <= by + 1; ++jy) {
        if (jy < 0 || jy >= a.nby) continue;
        for (int jx = bx - 1; jx <= bx + 1; ++jx) {
            if (jx < 0 || jx >= a.nbx) continue;
            const int b = jy * a.nbx + jx;
            for (int p = a.bin_start[b]; p < a.bin_start[b + 1]; ++p) {
                const int s = a.bin_stn[p];
                const double d = dist_rn(x, y, a.stn_x[s], a.stn_y[s]);
                if (d < a.R) {
                    if (n < a.cap) {
                        a.idx[(int64_t)n * a.n_cells + c] = s;
                        a.val[(int64_t)n * a.n_cells + c] = vg_eval(a.vg, d, a.covar_flag, a.min_vg_val) - a.F;
                    }
                    ++n;
                }
            }
        }
    }
    a.cnt[c] = n;   // may exceed cap: the caller rebuilds with a larger cap
}

constexpr int LOC_REG = 4;      // near stations kept in registers
constexpr int LOC_ROWS = 64;    // rows per block

struct LocalEstArgs {
    const double* coef;         // [n_rows, kpad] row-major
    const double* coef_t;       // [kpad, coef_t_ld] transposed copy (fast kernel) or NULL
    int64_t coef_t_ld;
    const double* base;         // [n_rows] F * sum_k coef + constant border term
    int64_t n_rows;
    int kpad, n_stn, n_drifts;
    const double* cell_drift;   // [n_drifts, n_cells]
    int64_t n_cells;
    int cap;
    const int32_t* cnt;
    const int32_t* idx;
    const double* val;
    const int32_t* row_dst;
    void* out;
    int64_t out_ld;
    int out_f64;
    const int32_t* cell_pos;
    int has_lo, has_hi;
    double lo, hi;
    const int32_t* tile_cnt;    // optional tile tables (see spx_local)
    const int32_t* tile_stn;
    const uint8_t* slot;
    int swap_grid;              // 1: blockIdx.x = row block, blockIdx.y = cell tile
};

// Distinct near stations of every tile of SPX_LOCAL_TILE cells: a station bitmap in shared
// memory (atomicOr), word prefix sums -> ascending list and the slot of every (cell, j).
constexpr int LOC_TILE_WORDS = 2048;        // up to 65536 stations

__global__ void __launch_bounds__(SPX_LOCAL_TILE) k_local_tiles(
    const int32_t* __restrict__ cnt, const int32_t* __restrict__ idx, int64_t n_cells, int cap,
    int n_stn, int32_t* __restrict__ tile_cnt, int32_t* __restrict__ tile_stn,
    uint8_t* __restrict__ slot) {
    __shared__ uint32_t bm[LOC_TILE_WORDS];
    __shared__ int pre[LOC_TILE_WORDS];
    __shared__ int part[SPX_LOCAL_TILE];
    const int tid = threadIdx.x;
    const int n_words = (n_stn + 31) >> 5;
    for (int w = tid; w < n_words; w += SPX_LOCAL_TILE) bm[w] = 0u;
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * SPX_LOCAL_TILE + tid;
    const int n = (c < n_cells) ? min(cnt[c], cap) : 0;
    for (int j = 0; j < n; ++j) {
        const int id = idx[(int64_t)j * n_cells + c];
        atomicOr(&bm[id >> 5], 1u << (id & 31));
    }
    __syncthreads();
    // exclusive prefix of the word popcounts: each thread owns a contiguous run of words
    const int per = (n_words + SPX_LOCAL_TILE - 1) / SPX_LOCAL_TILE;
    const int w0 = tid * per, w1 = min(n_words, w0 + per);
    int sum = 0;
    for (int w = w0; w < w1; ++w) sum += __popc(bm[w]);
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < SPX_LOCAL_TILE; o <<= 1) {
        const int v = (tid >= o) ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - sum;
    for (int w = w0; w < w1; ++w) {
        pre[w] = run;
        run += __popc(bm[w]);
    }
    const int total = part[SPX_LOCAL_TILE - 1];
    __syncthreads();
    if (total > SPX_LOCAL_TILE_CAP) {
        if (tid == 0) tile_cnt[blockIdx.x] = -1;
        return;
    }
    if (tid == 0) tile_cnt[blockIdx.x] = total;
    for (int w = tid; w < n_words; w += SPX_LOCAL_TILE) {
        uint32_t b = bm[w];
        int k = pre[w];
        while (b) {
            const int bit = __ffs(b) - 1;
            tile_stn[(int64_t)blockIdx.x * SPX_LOCAL_TILE_CAP + k++] = (w << 5) + bit;
            b &= b - 1;
        }
    }
    for (int j = 0; j < n; ++j) {
        const int id = idx[(int64_t)j * n_cells + c];
        slot[(int64_t)j * n_cells + c] =
            (uint8_t)(pre[id >> 5] + __popc(bm[id >> 5] & ((1u << (id & 31)) - 1u)));
    }
}

// One thread per cell, LOC_ROWS rows per block.  The cell's near stations live in
// registers; the loop over them is bounded by the warp-wide maximum so that the
// (typical) cells with 0-2 stations in range issue no dead instructions.  Per
// row: a broadcast read of base / destination, <= n gathers from the coefficient
// row (L1 / L2 resident) and one coalesced store.
template <typename OutT, bool DRIFT>
__global__ void __launch_bounds__(256) k_estimate_local(LocalEstArgs a) {
    __shared__ double sbase[LOC_ROWS];
    __shared__ int sdst[LOC_ROWS];
    const int64_t r_beg = (int64_t)blockIdx.y * LOC_ROWS;
    const int nr = (int)min((int64_t)LOC_ROWS, a.n_rows - r_beg);
    if (threadIdx.x < nr) {
        sbase[threadIdx.x] = a.base[r_beg + threadIdx.x];
        sdst[threadIdx.x] = a.row_dst[r_beg + threadIdx.x];
    }
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = c < a.n_cells;
    const int n = ok ? min(a.cnt[c], a.cap) : 0;
    int ri[LOC_REG];
    double rv[LOC_REG];
#pragma unroll
    for (int j = 0; j < LOC_REG; ++j) {
        ri[j] = (j < n) ? a.idx[(int64_t)j * a.n_cells + c] : 0;
        rv[j] = (j < n) ? a.val[(int64_t)j * a.n_cells + c] : 0.0;
    }
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    double dr[4] = {0.0, 0.0, 0.0, 0.0};
    if (DRIFT && ok) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (d < a.n_drifts) dr[d] = a.cell_drift[(int64_t)d * a.n_cells + c];
    }
    const int64_t col = ok ? (a.cell_pos ? (int64_t)a.cell_pos[c] : c) : 0;
    OutT* __restrict__ outp = reinterpret_cast<OutT*>(a.out) + col;
    const double* __restrict__ crow = a.coef + r_beg * a.kpad;
    constexpr int RU = 8;   // rows in flight: the gathers of RU rows are issued together
    for (int r0 = 0; r0 < nr; r0 += RU, crow += (int64_t)RU * a.kpad) {
        double g0[RU], g1[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const bool live = (r0 + u < nr);
            g0[u] = (live && 0 < n) ? crow[(int64_t)u * a.kpad + ri[0]] : 0.0;
            g1[u] = (live && 1 < n) ? crow[(int64_t)u * a.kpad + ri[1]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            if (r >= nr) break;
            const double* __restrict__ cr = crow + (int64_t)u * a.kpad;
            const int dst = sdst[r];
            double z = sbase[r];
            if (DRIFT) {
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if (d < a.n_drifts) z = fma(cr[a.n_stn + 1 + d], dr[d], z);
            }
            z = fma(g0[u], rv[0], z);
            z = fma(g1[u], rv[1], z);
            if (nmax > 2) {
                if (2 < n) z = fma(cr[ri[2]], rv[2], z);
                if (3 < n) z = fma(cr[ri[3]], rv[3], z);
                for (int j = LOC_REG; j < n; ++j)
                    z = fma(cr[a.idx[(int64_t)j * a.n_cells + c]], a.val[(int64_t)j * a.n_cells + c], z);
            }
            if (a.has_lo && z < a.lo) z = a.lo;
            if (a.has_hi && z > a.hi) z = a.hi;
            if (ok && dst >= 0) outp[(int64_t)dst * a.out_ld] = static_cast<OutT>(z);
        }
    }
}

// Fast variant: float output, identity cell order, no drift.  The kernel only writes
// (4 B per cell-step) and is bound by instruction issue unless every row is cheap, so:
//  * the coefficients are read from a TRANSPOSED copy  coef_t[k, row]: the four rows of
//    one unrolled step are two 16-byte loads at an immediate offset from a per-thread
//    pointer (row-major: one load + 64-bit address arithmetic per row and station);
//  * per-row base values / destination offsets come from shared memory, four rows per
//    pair of 16-byte loads;
//  * the number of gathers is a warp-uniform compile-time case (NG = 0, 1, 2, "many");
//  * the clamp is a compile-time option applied after the conversion (rounding is
//    monotone: float(clamp(z)) == clamp(float(z)) with float(lo), float(hi)).
constexpr int LOC_SLD = 130;    // row pitch of a staged coefficient slice (128 rows + 2: slices
                                // of different stations start 4 banks apart)

template <bool SM>
__device__ __forceinline__ double2 ld_coef2(const double* p) {
    if (SM) {
        // the pointer is known to be a shared-memory address only at run time: spell the
        // state space out, otherwise the compiler emits generic loads (LD.E, not LDS)
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                     : "=d"(v.x), "=d"(v.y)
                     : "r"((unsigned)__cvta_generic_to_shared(p)));
        return v;
    }
    return __ldg(reinterpret_cast<const double2*>(p));
}

// SM: the gathers read the tile's staged slices in shared memory (ct = its base, slice
// pitch LOC_SLD, station -> a.slot) instead of the global transposed copy.
// BULK (CPT == 1): the four values of a row group go to a shared-memory staging tile
// stage[2][LOC_RB rows][256 cells]; every LOC_RB rows thread 0 hands the tile to the TMA
// engine, one cp.async.bulk (shared -> global) per row segment of n_seg bytes, while the
// block computes the next LOC_RB rows into the other buffer.  One named barrier per
// LOC_RB rows; all 256 threads of the block run this loop (same trip count in every
// warp-uniform NG variant).
constexpr int LOC_RB = 8;

__device__ __forceinline__ void bulk_rows_out(const float* tile, float* outp,
                                              const int64_t* soff, int r0, int n_rows,
                                              int n_seg) {
    // caller: thread 0 only, after the barrier that published the tile
    for (int i = 0; i < n_rows; ++i) {
        float* g = outp + soff[r0 + i];
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tile + i * 256);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(g), "r"(sa), "r"(n_seg) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int NG, bool CLAMP, int CPT, bool SM, bool BULK = false>
__device__ __forceinline__ void local_rows_t(const LocalEstArgs& a, const double* sbase,
                                             const int64_t* soff, int nr, int64_t c0,
                                             const int* n, const double* const* p0,
                                             const double* const* p1, const double* v0,
                                             const double* v1, float* outp, const double* ct,
                                             float flo, float fhi, bool pair,
                                             float* stage = nullptr, int n_seg = 0,
                                             bool active = true) {
    int r = 0;
    const int tid = threadIdx.x;
    for (; r + 4 <= nr; r += 4) {
        const double2 b01 = *reinterpret_cast<const double2*>(sbase + r);
        const double2 b23 = *reinterpret_cast<const double2*>(sbase + r + 2);
        float f[CPT][4];
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            double z0 = b01.x, z1 = b01.y, z2 = b23.x, z3 = b23.y;
            if (NG >= 1 && 0 < n[q]) {
                const double2 g01 = ld_coef2<SM>(p0[q] + r);
                const double2 g23 = ld_coef2<SM>(p0[q] + r + 2);
                z0 = fma(g01.x, v0[q], z0); z1 = fma(g01.y, v0[q], z1);
                z2 = fma(g23.x, v0[q], z2); z3 = fma(g23.y, v0[q], z3);
            }
            if (NG >= 2 && 1 < n[q]) {
                const double2 g01 = ld_coef2<SM>(p1[q] + r);
                const double2 g23 = ld_coef2<SM>(p1[q] + r + 2);
                z0 = fma(g01.x, v1[q], z0); z1 = fma(g01.y, v1[q], z1);
                z2 = fma(g23.x, v1[q], z2); z3 = fma(g23.y, v1[q], z3);
            }
            if (NG >= 3) {
                const int64_t c = c0 + q;
                for (int j = 2; j < n[q]; ++j) {
                    const double* pj =
                        SM ? ct + (int)a.slot[(int64_t)j * a.n_cells + c] * LOC_SLD + r
                           : ct + (int64_t)a.idx[(int64_t)j * a.n_cells + c] * a.coef_t_ld + r;
                    const double vj = a.val[(int64_t)j * a.n_cells + c];
                    const double2 g01 = ld_coef2<SM>(pj);
                    const double2 g23 = ld_coef2<SM>(pj + 2);
                    z0 = fma(g01.x, vj, z0); z1 = fma(g01.y, vj, z1);
                    z2 = fma(g23.x, vj, z2); z3 = fma(g23.y, vj, z3);
                }
            }
            f[q][0] = (float)z0; f[q][1] = (float)z1; f[q][2] = (float)z2; f[q][3] = (float)z3;
            if (CLAMP) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    // comparisons with NaN are false: a NaN estimate stays NaN
                    f[q][u] = (f[q][u] < flo) ? flo : f[q][u];
                    f[q][u] = (f[q][u] > fhi) ? fhi : f[q][u];
                }
            }
        }
        if (BULK) {
            float* st = stage + ((r / LOC_RB) & 1) * (LOC_RB * 256) + (r & (LOC_RB - 1)) * 256 + tid;
#pragma unroll
            for (int u = 0; u < 4; ++u) st[u * 256] = f[0][u];
            const bool last = r + 8 > nr;                  // no further full row group
            if ((r & (LOC_RB - 1)) == LOC_RB - 4 || last) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // the other buffer's previous copies must have read it before anyone refills it
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) {
                    const int g0 = r & ~(LOC_RB - 1);
                    bulk_rows_out(stage + ((r / LOC_RB) & 1) * (LOC_RB * 256), outp - tid, soff,
                                  g0, r + 4 - g0, n_seg);
                }
            }
            continue;
        }
        const longlong2 o01 = *reinterpret_cast<const longlong2*>(soff + r);
        const longlong2 o23 = *reinterpret_cast<const longlong2*>(soff + r + 2);
        const int64_t o[4] = {o01.x, o01.y, o23.x, o23.y};
#pragma unroll
        // streaming stores (st.global.cs): the 5 GB field passes through L2 once and must
        // not evict the ~30 MB of tables / coefficient slices every row block re-reads
        for (int u = 0; u < 4; ++u) {
            if (CPT == 2) {
                if (pair)
                    __stcs(reinterpret_cast<float2*>(outp + o[u]), make_float2(f[0][u], f[CPT - 1][u]));
                else
                    __stcs(outp + o[u], f[0][u]);
            } else {
                __stcs(outp + o[u], f[0][u]);
            }
        }
    }
    if (BULK) {
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (!active) return;
    }
    for (; r < nr; ++r) {
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            if (q == 1 && !pair) break;
            const int64_t c = c0 + q;
            double z = sbase[r];
            if (NG >= 1 && 0 < n[q]) z = fma(p0[q][r], v0[q], z);
            if (NG >= 2 && 1 < n[q]) z = fma(p1[q][r], v1[q], z);
            if (NG >= 3)
                for (int j = 2; j < n[q]; ++j)
                    z = fma(SM ? ct[(int)a.slot[(int64_t)j * a.n_cells + c] * LOC_SLD + r]
                               : ct[(int64_t)a.idx[(int64_t)j * a.n_cells + c] * a.coef_t_ld + r],
                            a.val[(int64_t)j * a.n_cells + c], z);
            float fv = (float)z;
            if (CLAMP) {
                fv = (fv < flo) ? flo : fv;
                fv = (fv > fhi) ? fhi : fv;
            }
            __stcs(outp + soff[r] + q, fv);
        }
    }
}

// CPT cells per thread (2: one 8-byte store per row and thread; needs an even out_ld).
// TILE (CPT == 1, ROWS == 128, tile tables present): the block first copies the
// coefficient slices coef_t[stn, r_beg : r_beg + 128] of its tile's <= SPX_LOCAL_TILE_CAP
// distinct near stations into shared memory (coalesced 16-byte loads, ~1 KB per station
// against 128 KB of output per block) and every gather of the row loop becomes a
// shared-memory load: no L1 misses, a fraction of the latency of the global gathers
// the untiled variant waits on.
template <bool CLAMP, int ROWS, int CPT, bool TILE, bool BULK = false>
__global__ void __launch_bounds__(256, 4) k_estimate_local_fast(LocalEstArgs a) {
    extern __shared__ __align__(128) float stage[];      // BULK: [2][LOC_RB][256]
    __shared__ __align__(16) double sbase[ROWS];
    __shared__ __align__(16) int64_t soff[ROWS];
    __shared__ __align__(16) double sct[TILE ? SPX_LOCAL_TILE_CAP * LOC_SLD : 2];
    static_assert(!TILE || (CPT == 1 && ROWS == 128), "tile variant: 256 cells x 128 rows");
    static_assert(!BULK || TILE, "bulk stores come with the tile variant");
    const int tid = threadIdx.x;
    // swap_grid: the row blocks of one cell tile are launched back to back, so that all
    // but the first find the tile's tables and coefficient slices in L2
    const unsigned bx = a.swap_grid ? blockIdx.y : blockIdx.x;     // cell tile
    const unsigned by = a.swap_grid ? blockIdx.x : blockIdx.y;     // row block
    const int64_t r_beg = (int64_t)by * ROWS;
    const int nr = (int)min((int64_t)ROWS, a.n_rows - r_beg);
    const double* ct = a.coef_t + r_beg;
    const int64_t c0 = ((int64_t)bx * blockDim.x + tid) * CPT;
    const bool pair = (CPT == 2) && (c0 + 1 < a.n_cells);
    // A block lives for only ROWS / 4 loop iterations, so the latency of its prologue
    // matters: every table entry it may need is requested up front, unconditionally and
    // independently (entries j >= cnt are unspecified and masked after arrival), instead
    // of as a chain  cnt -> slot / val,  tile_cnt -> tile_stn -> slices.
    int n[CPT], t0[CPT], t1[CPT];
    double v0[CPT], v1[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int64_t c = min(c0 + q, a.n_cells - 1);
        n[q] = a.cnt[c];
        t0[q] = TILE ? (int)a.slot[c] : a.idx[c];
        t1[q] = TILE ? (int)a.slot[a.n_cells + c] : a.idx[a.n_cells + c];
        v0[q] = a.val[c];
        v1[q] = a.val[a.n_cells + c];
    }
    bool staged = false;
    if (TILE) {
        constexpr int SPT = SPX_LOCAL_TILE_CAP / 4;      // slices per thread (4 per pass)
        const int32_t* __restrict__ tl = a.tile_stn + (int64_t)bx * SPX_LOCAL_TILE_CAP;
        const int U = a.tile_cnt[bx];
        int stn[SPT];
#pragma unroll
        for (int it = 0; it < SPT; ++it) stn[it] = tl[(tid >> 6) + 4 * it];
        staged = U >= 0;                         // block-uniform
        const int r2 = 2 * (tid & 63);
        const int r_lim = (int)min((int64_t)ROWS, a.coef_t_ld - r_beg);   // even
        double2 g[SPT];
#pragma unroll
        for (int it = 0; it < SPT; ++it) {
            const int sl = (tid >> 6) + 4 * it;
            if (sl < U && r2 < r_lim)
                g[it] = __ldg(reinterpret_cast<const double2*>(
                    ct + (int64_t)stn[it] * a.coef_t_ld + r2));
        }
        for (int i = tid; i < nr; i += 256) {
            sbase[i] = a.base[r_beg + i];
            soff[i] = (int64_t)a.row_dst[r_beg + i] * a.out_ld;
        }
#pragma unroll
        for (int it = 0; it < SPT; ++it) {
            const int sl = (tid >> 6) + 4 * it;
            if (sl < U && r2 < r_lim)
                *reinterpret_cast<double2*>(sct + sl * LOC_SLD + r2) = g[it];
        }
    } else {
        for (int i = tid; i < nr; i += blockDim.x) {
            sbase[i] = a.base[r_beg + i];
            soff[i] = (int64_t)a.row_dst[r_beg + i] * a.out_ld;
        }
    }
    __syncthreads();
    const bool active = c0 < a.n_cells;
    if (!BULK && !active) return;                // no further block-wide barrier below
    const double* p0[CPT];
    const double* p1[CPT];
    int nloc = 0;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        n[q] = ((q == 0 || pair) && active) ? min(n[q], a.cap) : 0;
        v0[q] = (0 < n[q]) ? v0[q] : 0.0;
        v1[q] = (1 < n[q]) ? v1[q] : 0.0;
        if (TILE && staged) {
            p0[q] = sct + ((0 < n[q]) ? t0[q] : 0) * LOC_SLD;
            p1[q] = sct + ((1 < n[q]) ? t1[q] : 0) * LOC_SLD;
        } else {
            int i0 = t0[q], i1 = t1[q];
            if (TILE) {                          // overflowing tile: station ids instead of slots
                const int64_t c = min(c0 + q, a.n_cells - 1);
                i0 = a.idx[c];
                i1 = a.idx[a.n_cells + c];
            }
            p0[q] = ct + (int64_t)((0 < n[q]) ? i0 : 0) * a.coef_t_ld;
            p1[q] = ct + (int64_t)((1 < n[q]) ? i1 : 0) * a.coef_t_ld;
        }
        nloc = max(nloc, n[q]);
    }
    const int nmax = __reduce_max_sync(__activemask(), nloc);
    const float flo = a.has_lo ? (float)a.lo : -CUDART_INF_F;
    const float fhi = a.has_hi ? (float)a.hi : CUDART_INF_F;
    float* outp = reinterpret_cast<float*>(a.out) + c0;
    // bytes of the tile's row segment (the last tile of a row may be short)
    const int n_seg = (int)min((int64_t)256, a.n_cells - (int64_t)bx * 256) * 4;
#define SPX_LOCAL_ROWS_CALL(NG, SM, CT)                                                       \
    local_rows_t<NG, CLAMP, CPT, SM, BULK>(a, sbase, soff, nr, c0, n, p0, p1, v0, v1, outp, CT, \
                                           flo, fhi, pair, stage, n_seg, active)
    if (TILE && staged) {
        if (nmax == 0) SPX_LOCAL_ROWS_CALL(0, true, sct);
        else if (nmax == 1) SPX_LOCAL_ROWS_CALL(1, true, sct);
        else if (nmax == 2) SPX_LOCAL_ROWS_CALL(2, true, sct);
        else SPX_LOCAL_ROWS_CALL(3, true, sct);
    } else {
        if (nmax == 0) SPX_LOCAL_ROWS_CALL(0, false, ct);
        else if (nmax == 1) SPX_LOCAL_ROWS_CALL(1, false, ct);
        else if (nmax == 2) SPX_LOCAL_ROWS_CALL(2, false, ct);
        else SPX_LOCAL_ROWS_CALL(3, false, ct);
    }
#undef SPX_LOCAL_ROWS_CALL
}


}  // namespace spx

extern "C" int spx_local_build_dev(const spx_local* l, void* stream) {
    using namespace spx;
    if (!l || l->cap < 1) {
        set_error("local_build: bad argument");
        return SPX_EINVAL;
    }
    if (l->n_cells == 0) return SPX_OK;
    LocalBuildArgs a;
    a.stn_x = l->stn_x;
    a.stn_y = l->stn_y;
    a.bin_start = l->bin_start;
    a.bin_stn = l->bin_stn;
    a.x0 = l->x0;
    a.y0 = l->y0;
    a.inv_bin = l->inv_bin;
    a.nbx = l->nbx;
    a.nby = l->nby;
    a.R = l->R;
    a.F = l->F;
    a.cell_x = l->cell_x;
    a.cell_y = l->cell_y;
    a.n_cells = l->n_cells;
    a.cap = l->cap;
    a.cnt = l->cnt;
    a.idx = l->idx;
    a.val = l->val;
    a.vg = to_dev(l->vg);
    a.covar_flag = l->covar_flag;
    a.min_vg_val = l->min_vg_val;
    k_local_build<<<(unsigned)((l->n_cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    SPX_CHECK_LAUNCH("k_local_build");
    return SPX_OK;
}

extern "C" int spx_local_tiles_dev(const spx_local* l, void* stream) {
    using namespace spx;
    if (!l || !l->cnt || !l->idx || !l->tile_cnt || !l->tile_stn || !l->slot || l->cap < 1) {
        set_error("local_tiles: null argument");
        return SPX_EINVAL;
    }
    if (l->n_stn < 1 || l->n_stn > LOC_TILE_WORDS * 32) {
        set_error("local_tiles: n_stn must be in [1, %d]", LOC_TILE_WORDS * 32);
        return SPX_EINVAL;
    }
    if (l->n_cells == 0) return SPX_OK;
    const int64_t n_tiles = (l->n_cells + SPX_LOCAL_TILE - 1) / SPX_LOCAL_TILE;
    k_local_tiles<<<(unsigned)n_tiles, SPX_LOCAL_TILE, 0, (cudaStream_t)stream>>>(
        l->cnt, l->idx, l->n_cells, l->cap, l->n_stn, l->tile_cnt, l->tile_stn, l->slot);
    SPX_CHECK_LAUNCH("k_local_tiles");
    return SPX_OK;
}

static int g_local_bulk = -1;       // -1: environment SPX_LOCAL_BULK (default off: measured slower)

extern "C" int spx_local_set_bulk(int on) {
    const int prev = g_local_bulk;
    g_local_bulk = on;
    return prev;
}

extern "C" int spx_estimate_local_dev(const spx_local* l, void* stream) {
    using namespace spx;
    if (!l) {
        set_error("estimate_local: null argument");
        return SPX_EINVAL;
    }
    if (l->n_cells == 0 || l->n_rows == 0) return SPX_OK;
    if (l->n_drifts > 4) {
        set_error("estimate_local: more than 4 drifts");
        return SPX_EINVAL;
    }
    LocalEstArgs a;
    a.coef = l->coef;
    a.coef_t = l->coef_t;
    a.coef_t_ld = l->coef_t_ld;
    a.base = l->base;
    a.n_rows = l->n_rows;
    a.kpad = l->kpad;
    a.n_stn = l->n_stn;
    a.n_drifts = l->n_drifts;
    a.cell_drift = l->cell_drift;
    a.n_cells = l->n_cells;
    a.cap = l->cap;
    a.cnt = l->cnt;
    a.idx = l->idx;
    a.val = l->val;
    a.row_dst = l->row_dst;
    a.out = l->out;
    a.out_ld = l->out_ld;
    a.out_f64 = l->out_f64;
    a.cell_pos = l->cell_pos;
    a.has_lo = l->has_lo;
    a.has_hi = l->has_hi;
    a.lo = l->lo;
    a.hi = l->hi;
    a.tile_cnt = nullptr;
    a.tile_stn = nullptr;
    a.slot = nullptr;
    a.swap_grid = 0;
    const int64_t row_blocks = (l->n_rows + LOC_ROWS - 1) / LOC_ROWS;
    if (row_blocks > 65535) {
        set_error("estimate_local: too many rows in one launch");
        return SPX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool drift = l->n_drifts > 0;
    if (!l->out_f64 && !drift && l->cell_pos == nullptr && l->rows_all_valid &&
        l->coef_t != nullptr) {
        if (l->coef_t_ld % 4 != 0 || l->coef_t_ld < l->n_rows) {
            set_error("estimate_local: coef_t_ld must be a multiple of 4 and >= n_rows");
            return SPX_EINVAL;
        }
        static const int rows_knob = getenv("SPX_LOCAL_ROWS") ? atoi(getenv("SPX_LOCAL_ROWS")) : 128;
        static const int cpt_knob = getenv("SPX_LOCAL_CPT") ? atoi(getenv("SPX_LOCAL_CPT")) : 1;
        const bool clamp = l->has_lo || l->has_hi;
        const int rows = (rows_knob == 64) ? 64 : 128;
        // two cells per thread need 8-byte aligned row starts
        const int cpt = (cpt_knob == 2 && l->out_ld % 2 == 0 &&
                         (reinterpret_cast<uintptr_t>(l->out) & 7) == 0) ? 2 : 1;
        const int64_t per_blk = 256 * (int64_t)cpt;
        dim3 g1((unsigned)((l->n_cells + per_blk - 1) / per_blk),
                (unsigned)((l->n_rows + rows - 1) / rows));
        static const int swap_knob = getenv("SPX_LOCAL_SWAP") ? atoi(getenv("SPX_LOCAL_SWAP")) : 1;
        a.swap_grid = 0;
        if (swap_knob && g1.x <= 65535u) {
            a.swap_grid = 1;
            g1 = dim3(g1.y, g1.x);
        }
#define SPX_LOCAL_LAUNCH(CL, RW, CP) \
    k_estimate_local_fast<CL, RW, CP, false><<<g1, 256, 0, st>>>(a)
        static const int tile_knob = getenv("SPX_LOCAL_TILES") ? atoi(getenv("SPX_LOCAL_TILES")) : 1;
        if (tile_knob && rows == 128 && cpt == 1 && l->tile_cnt && l->tile_stn && l->slot) {
            a.tile_cnt = l->tile_cnt;
            a.tile_stn = l->tile_stn;
            a.slot = l->slot;
            // bulk (TMA) stores of shared-memory staged row segments: 16-byte aligned rows
            static const int bulk_env = getenv("SPX_LOCAL_BULK") ? atoi(getenv("SPX_LOCAL_BULK")) : 0;
            const int bulk_knob = g_local_bulk < 0 ? bulk_env : g_local_bulk;
            const bool bulk = bulk_knob && l->out_ld % 4 == 0 && l->n_cells % 4 == 0 &&
                              (reinterpret_cast<uintptr_t>(l->out) & 15) == 0;
            if (bulk) {
                constexpr int stage_bytes = 2 * LOC_RB * 256 * (int)sizeof(float);
                static bool attr_set = false;
                if (!attr_set) {
                    SPX_CUDA(cudaFuncSetAttribute(k_estimate_local_fast<true, 128, 1, true, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  stage_bytes));
                    SPX_CUDA(cudaFuncSetAttribute(k_estimate_local_fast<false, 128, 1, true, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  stage_bytes));
                    attr_set = true;
                }
                if (clamp)
                    k_estimate_local_fast<true, 128, 1, true, true><<<g1, 256, stage_bytes, st>>>(a);
                else
                    k_estimate_local_fast<false, 128, 1, true, true><<<g1, 256, stage_bytes, st>>>(a);
                SPX_CHECK_LAUNCH("k_estimate_local_fast(tile, bulk)");
                return SPX_OK;
            }
            if (clamp) k_estimate_local_fast<true, 128, 1, true><<<g1, 256, 0, st>>>(a);
            else k_estimate_local_fast<false, 128, 1, true><<<g1, 256, 0, st>>>(a);
            SPX_CHECK_LAUNCH("k_estimate_local_fast(tile)");
            return SPX_OK;
        }
        if (rows == 64) {
            if (cpt == 2) { if (clamp) SPX_LOCAL_LAUNCH(true, 64, 2); else SPX_LOCAL_LAUNCH(false, 64, 2); }
            else { if (clamp) SPX_LOCAL_LAUNCH(true, 64, 1); else SPX_LOCAL_LAUNCH(false, 64, 1); }
        } else {
            if (cpt == 2) { if (clamp) SPX_LOCAL_LAUNCH(true, 128, 2); else SPX_LOCAL_LAUNCH(false, 128, 2); }
            else { if (clamp) SPX_LOCAL_LAUNCH(true, 128, 1); else SPX_LOCAL_LAUNCH(false, 128, 1); }
        }
#undef SPX_LOCAL_LAUNCH
        SPX_CHECK_LAUNCH("k_estimate_local_fast");
        return SPX_OK;
    }
    dim3 grid((unsigned)((l->n_cells + 255) / 256), (unsigned)row_blocks);
    if (l->out_f64) {
        if (drift) k_estimate_local<double, true><<<grid, 256, 0, st>>>(a);
        else k_estimate_local<double, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (drift) k_estimate_local<float, true><<<grid, 256, 0, st>>>(a);
        else k_estimate_local<float, false><<<grid, 256, 0, st>>>(a);
    }
    SPX_CHECK_LAUNCH("k_estimate_local");
    return SPX_OK;
}
