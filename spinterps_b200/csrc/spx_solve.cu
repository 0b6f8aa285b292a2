// Batched kriging systems: assembly from station coordinates, LU factorisation
// with partial pivoting (one thread block per system, warp-shuffle pivot
// search) and multi-right-hand-side solves whose solutions are scattered into
// the packed coefficient matrix consumed by the estimate contraction.
//
// Replaces, per (availability group x variogram x kriging kind):
//   _get_vars_arr_subset   interp/steps.py:170-243   (assembly)
//   np.linalg.pinv         interp/steps.py:351       (factor; non-singular case)
//   np.matmul(inv, rhs_i)  interp/steps.py:416       (dual form)
#include <type_traits>

#include "spx_common.cuh"

namespace spx {

__host__ __device__ __forceinline__ int64_t coef_offset(int64_t row, int64_t col, int64_t kpad) {
    return ((row / SPX_BM) * (kpad >> 2) + (col >> 2)) * (int64_t)(SPX_BM * 4) +
           ((row % SPX_BM) >> 3) * 32 + (row & 7) * 4 + (col & 3);
}

__device__ __forceinline__ int border_of(int kind, int n_drifts) {
    return kind == SPX_KRG_OK ? 1 : (kind == SPX_KRG_SK ? 0 : 1 + n_drifts);
}

// ------------------------------------------------------------- assembly

__global__ void __launch_bounds__(256) k_assemble(spx_systems s, const spx_vg* __restrict__ vgs,
                                                  double min_vg_val) {
    const int sys = blockIdx.y;
    const int n = s.sys_n[sys];
    const int kind = s.sys_kind[sys];
    const int nd = s.n_drifts;
    const int m = n + border_of(kind, nd);
    double* __restrict__ W = s.work + s.sys_w_off[sys];
    const int32_t* __restrict__ stn = s.stn_list + s.sys_stn_off[sys];

    __shared__ VgDev vg;
    if (threadIdx.x == 0) {
        const spx_vg& v = vgs[s.sys_vg[sys]];
        vg.n_terms = v.n_terms;
        for (int i = 0; i < SPX_VG_MAX_TERMS; ++i) {
            vg.types[i] = v.types[i];
            vg.sills[i] = v.sills[i];
            vg.ranges[i] = v.ranges[i];
        }
    }
    __syncthreads();
    const int covar = (kind == SPX_KRG_SK);

    const int64_t total = (int64_t)m * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int j = (int)(idx / m);  // column
        const int i = (int)(idx - (int64_t)j * m);  // row (fastest: column-major)
        double v;
        if (i < n && j < n) {
            const int a = stn[i], b = stn[j];
            const double h = dist_rn(s.stn_x[a], s.stn_y[a], s.stn_x[b], s.stn_y[b]);
            v = vg_eval(vg, h, covar, min_vg_val);
        } else if (i >= n && j >= n) {
            v = 0.0;  // steps.py:215, :225
        } else {
            const int b = (i >= n) ? (i - n) : (j - n);   // border index
            const int a = (i >= n) ? j : i;               // station slot
            v = (b == 0) ? 1.0 : s.stn_drift[(int64_t)stn[a] * nd + (b - 1)];  // steps.py:213-234
        }
        W[idx] = v;
    }
}

// ------------------------------------------------------------- LU factor

// Unblocked right-looking LU, column-major, partial pivoting (first largest
// |value| like LAPACK idamax).  The matrix stays in global memory (L2 resident:
// <= 8 MB per system), all accesses run down columns (coalesced).
__global__ void __launch_bounds__(256) k_lu_factor(spx_systems s) {
    const int sys = blockIdx.x;
    const int m = s.sys_n[sys] + border_of(s.sys_kind[sys], s.n_drifts);
    double* __restrict__ W = s.work + s.sys_w_off[sys];
    int32_t* __restrict__ piv = s.piv + s.sys_piv_off[sys];

    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ int s_p;
    __shared__ double s_pv;
    __shared__ int s_info;

    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_info = 0;

    for (int k = 0; k < m; ++k) {
        // -- pivot search down column k
        double bv = -1.0;
        int bi = k;
        const double* colk = W + (int64_t)k * m;
        for (int i = k + tid; i < m; i += 256) {
            const double a = fabs(colk[i]);
            if (a > bv) { bv = a; bi = i; }  // first largest; NaN never selected
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bv = lane < 8 ? red_v[lane] : -2.0;
            bi = lane < 8 ? red_i[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                s_p = bi;
                s_pv = colk[bi];
                piv[k] = bi;
                if (!(bv > 0.0) && s_info == 0) s_info = k + 1;  // zero / NaN pivot
            }
        }
        __syncthreads();
        const int p = s_p;
        const double pv = s_pv;
        // -- swap rows k and p across every column
        if (p != k) {
            for (int j = tid; j < m; j += 256) {
                double* c = W + (int64_t)j * m;
                const double t = c[k];
                c[k] = c[p];
                c[p] = t;
            }
        }
        __syncthreads();
        // -- scale the sub-column (true division, like dgetf2 for tiny pivots)
        double* ck = W + (int64_t)k * m;
        if (pv != 0.0)
            for (int i = k + 1 + tid; i < m; i += 256) ck[i] = ck[i] / pv;
        __syncthreads();
        // -- rank-1 update of the trailing matrix
        for (int j = k + 1 + wid; j < m; j += 8) {
            double* cj = W + (int64_t)j * m;
            const double ukj = cj[k];
            for (int i = k + 1 + lane; i < m; i += 32) cj[i] = fma(-ck[i], ukj, cj[i]);
        }
        __syncthreads();
    }
    if (tid == 0) s.info[sys] = s_info;
}

// ------------------------------------------------------------- blocked LU

// Right-looking blocked LU with partial pivoting, one thread block per system.
// Per panel of NB columns: (a) the panel (all remaining rows) is staged in
// shared memory and factored there (warp-shuffle pivot search), (b) row swaps
// are applied to the other columns, (c) U12 = L11^-1 A12 one column per thread,
// (d) the trailing matrix is updated A22 -= L21 U12 with FP64 tensor-core MMA
// (DMMA m8n8k4): L21 fragments straight from the panel in shared memory, U12
// tiles staged in shared memory, 32x32 accumulator tiles per warp read from and
// written back to global memory (L2) once per panel.
constexpr int LU_UT = 64;        // columns of U12 staged per trailing-update block
constexpr int LU_UP = LU_UT + 4; // pitch: == 4 (mod 16) -> conflict-free B fragments

__device__ __forceinline__ void dmma_lu(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NB>
__global__ void __launch_bounds__(256) k_lu_blocked(spx_systems s, int ldp) {
    extern __shared__ double lsm[];
    double* P = lsm;                       // [NB][ldp] panel, column-major
    double* Us = P + (size_t)NB * ldp;     // [NB][LU_UP] U12 tile, row-major (k, col)
    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ int s_p;
    __shared__ double s_pv;
    __shared__ int s_info;
    __shared__ int s_piv[NB];

    const int sys = blockIdx.x;
    const int m = s.sys_n[sys] + border_of(s.sys_kind[sys], s.n_drifts);
    double* __restrict__ W = s.work + s.sys_w_off[sys];
    int32_t* __restrict__ piv = s.piv + s.sys_piv_off[sys];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    if (tid == 0) s_info = 0;

    for (int k0 = 0; k0 < m; k0 += NB) {
        const int nb = min(NB, m - k0);
        const int rows = m - k0;
        // ---- (a) stage + factor the panel
        for (int c = wid; c < nb; c += 8) {
            const double* src = W + (int64_t)(k0 + c) * m + k0;
            for (int i = lane; i < rows; i += 32) P[(size_t)c * ldp + i] = src[i];
        }
        __syncthreads();
        for (int j = 0; j < nb; ++j) {
            double bv = -1.0;
            int bi = j;
            const double* colj = P + (size_t)j * ldp;
            for (int i = j + tid; i < rows; i += 256) {
                const double a = fabs(colj[i]);
                if (a > bv) { bv = a; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
            __syncthreads();
            if (wid == 0) {
                bv = lane < 8 ? red_v[lane] : -2.0;
                bi = lane < 8 ? red_i[lane] : 0x7fffffff;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) {
                    s_p = bi;
                    s_pv = colj[bi];
                    s_piv[j] = bi;
                    piv[k0 + j] = k0 + bi;
                    if (!(bv > 0.0) && s_info == 0) s_info = k0 + j + 1;
                }
            }
            __syncthreads();
            const int p = s_p;
            const double pvv = s_pv;
            if (p != j && tid < nb) {
                const double t = P[(size_t)tid * ldp + j];
                P[(size_t)tid * ldp + j] = P[(size_t)tid * ldp + p];
                P[(size_t)tid * ldp + p] = t;
            }
            __syncthreads();
            double* cj = P + (size_t)j * ldp;
            if (pvv != 0.0)
                for (int i = j + 1 + tid; i < rows; i += 256) cj[i] = cj[i] / pvv;
            __syncthreads();
            for (int c = j + 1 + wid; c < nb; c += 8) {
                double* cc = P + (size_t)c * ldp;
                const double ujc = cc[j];
                for (int i = j + 1 + lane; i < rows; i += 32) cc[i] = fma(-cj[i], ujc, cc[i]);
            }
            __syncthreads();
        }
        // ---- write the factored panel back
        for (int c = wid; c < nb; c += 8) {
            double* dst = W + (int64_t)(k0 + c) * m + k0;
            for (int i = lane; i < rows; i += 32) dst[i] = P[(size_t)c * ldp + i];
        }
        // ---- (b) row swaps on the columns outside the panel, (c) U12
        for (int c = tid; c < m; c += 256) {
            if (c >= k0 && c < k0 + nb) continue;
            double* col = W + (int64_t)c * m + k0;
            for (int j = 0; j < nb; ++j) {
                const int p = s_piv[j];
                if (p != j) {
                    const double t = col[j];
                    col[j] = col[p];
                    col[p] = t;
                }
            }
            if (c >= k0 + nb) {
                double x[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? col[j] : 0.0;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (j < nb) {
#pragma unroll
                        for (int i = j + 1; i < NB; ++i)
                            if (i < nb) x[i] = fma(-P[(size_t)j * ldp + i], x[j], x[i]);
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (j < nb) col[j] = x[j];
            }
        }
        __syncthreads();
        // ---- (d) trailing update with DMMA
        const int r0 = k0 + nb;            // first trailing row / column
        const int nt = m - r0;             // trailing size
        if (nt <= 0) break;
        const int row_tiles = (nt + 31) / 32;
        for (int cb = 0; cb < nt; cb += LU_UT) {
            const int ncol = min(LU_UT, nt - cb);
            // stage U12[:, cb:cb+ncol] -> Us[k][col]
            for (int idx = tid; idx < NB * LU_UT; idx += 256) {
                const int col = idx / NB, k = idx - col * NB;  // k fastest: contiguous in W
                double v = 0.0;
                if (col < ncol && k < nb) v = W[(int64_t)(r0 + cb + col) * m + k0 + k];
                Us[(size_t)k * LU_UP + col] = v;
            }
            __syncthreads();
            const int col_tiles = (ncol + 31) / 32;
            for (int tile = wid; tile < row_tiles * col_tiles; tile += 8) {
                const int rt = tile / col_tiles, ctile = tile - rt * col_tiles;
                const int prow0 = nb + rt * 32;          // row inside the panel
                const int grow0 = r0 + rt * 32;          // global row
                const int ccol0 = ctile * 32;            // column inside the U tile
                double acc[4][4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < NB; kk += 4) {
                    double af[4], bf[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        af[i] = P[(size_t)(kk + t4) * ldp + prow0 + i * 8 + g];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        bf[j] = Us[(size_t)(kk + t4) * LU_UP + ccol0 + j * 8 + g];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dmma_lu(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = grow0 + i * 8 + g;
                    if (row >= m) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = r0 + cb + ccol0 + j * 8 + t4 * 2 + e;
                            if (col < m && (ccol0 + j * 8 + t4 * 2 + e) < ncol) {
                                double* a = W + (int64_t)col * m + row;
                                *a = *a - acc[i][j][e];
                            }
                        }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) s.info[sys] = s_info;
}

// ------------------------------------------------------------- solve + scatter

// One thread block per right-hand side: b in shared memory, pivots applied,
// column-oriented forward (unit L) and backward (U) substitution reading the
// factors from global memory down columns.
__global__ void __launch_bounds__(256) k_lu_solve(spx_systems s, spx_rhs r) {
    extern __shared__ double b[];
    __shared__ double red[8];
    const int q = blockIdx.x;
    const int sys = r.rhs_sys[q];
    const int n = s.sys_n[sys];
    const int m = n + border_of(s.sys_kind[sys], s.n_drifts);
    const double* __restrict__ W = s.work + s.sys_w_off[sys];
    const int32_t* __restrict__ piv = s.piv + s.sys_piv_off[sys];
    const int32_t* __restrict__ stn = s.stn_list + s.sys_stn_off[sys];
    const int kind = r.rhs_kind[q];
    const int arg = r.rhs_arg[q];
    const int tid = threadIdx.x;

    for (int i = tid; i < m; i += 256) {
        double v = 0.0;
        if (kind == 0) {
            if (i < n) v = r.data[(int64_t)arg * r.n_stn + stn[i]];
        } else if (kind == 1) {
            if (i < n) v = 1.0;
        } else {
            v = (i == arg) ? 1.0 : 0.0;
        }
        b[i] = v;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < m; ++k) {
            const int p = piv[k];
            if (p != k) {
                const double t = b[k];
                b[k] = b[p];
                b[p] = t;
            }
        }
    }
    __syncthreads();
    // forward: L y = Pb
    for (int k = 0; k < m - 1; ++k) {
        const double xk = b[k];
        const double* ck = W + (int64_t)k * m;
        for (int i = k + 1 + tid; i < m; i += 256) b[i] = fma(-ck[i], xk, b[i]);
        __syncthreads();
    }
    // backward: U x = y
    for (int k = m - 1; k >= 0; --k) {
        const double* ck = W + (int64_t)k * m;
        const double xk = b[k] / ck[k];
        __syncthreads();  // everyone has read b[k] before it is overwritten
        if (tid == 0) b[k] = xk;
        for (int i = tid; i < k; i += 256) b[i] = fma(-ck[i], xk, b[i]);
        __syncthreads();
    }
    // scatter into the packed coefficient row
    const int64_t row = r.rhs_row[q];
    if (row >= 0) {
        for (int i = tid; i < m; i += 256) {
            const int col = (i < n) ? stn[i] : (r.n_stn + (i - n));
            r.coef[r.coef_row_major ? row * (int64_t)r.kpad + col
                                    : coef_offset(row, col, r.kpad)] = b[i];
        }
    }
    if (r.dense != nullptr)
        for (int i = tid; i < m; i += 256) r.dense[(int64_t)q * r.dense_ld + i] = b[i];
    if (r.resid != nullptr) {
        // || x - e_n ||_1 : exact solution of A x = [1_n; 0] is e_n for OK / EDK
        double acc = 0.0;
        for (int i = tid; i < m; i += 256) acc += fabs(b[i] - ((i == n) ? 1.0 : 0.0));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            r.resid[q] = t;
        }
    }
}


// ------------------------------------------------------------- downdated solves

constexpr int DD_RB = 8;  // right-hand sides per batch (one per warp)

__global__ void __launch_bounds__(256) k_downdate(spx_downdate d, int r_lo) {
    extern __shared__ double dsm[];
    const int sys = d.sys_order ? d.sys_order[blockIdx.x] : blockIdx.x;
    const int r = d.sys_r[sys];
    if (r <= r_lo) return;   // handled by k_downdate_reg
    const int n = d.sys_n[sys];
    const int M = d.n_stn + d.n_border;
    const int ld = d.max_r | 1;  // odd pitch: row swaps hit distinct banks
    double* S = dsm;                                   // [ld * max_r] column-major
    double* ys = S + (size_t)ld * d.max_r;             // [DD_RB][ld]
    int* mi = reinterpret_cast<int*>(ys + (size_t)DD_RB * ld);  // [max_r]
    int* pv = mi + d.max_r;                            // [max_r]
    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ int s_p;
    __shared__ double s_pv;
    __shared__ int s_info;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* __restrict__ miss = d.miss_list + d.sys_miss_off[sys];
    const int32_t* __restrict__ stn = d.stn_list + d.sys_stn_off[sys];
    const double* __restrict__ G = d.ginv;
    if (tid == 0) s_info = 0;
    for (int i = tid; i < r; i += 256) mi[i] = miss[i];
    __syncthreads();
    // S = G[Mi, Mi]; row mi[j] of G is read along ascending mi[i] (G symmetric)
    for (int idx = tid; idx < r * r; idx += 256) {
        const int j = idx / r, i = idx - j * r;
        S[i + (size_t)j * ld] = G[(int64_t)mi[j] * M + mi[i]];
    }
    __syncthreads();
    // ---- LU of S in shared memory, partial pivoting, ONE block barrier per column.
    // For r ~ 100 the factorisation is a chain of r dependent steps, so barriers and
    // shared-memory round trips (not flops) set its duration:
    //  * pivoting is implicit (rows stay in place, each lane keeps a bit mask of the rows
    //    it owns that were already used as pivots) -- no swap phase;
    //  * every warp repeats the pivot search of column k for itself -- no broadcast phase;
    //  * the multipliers l_i are kept in registers and written over column k one step
    //    later, when nobody reads that column any more -- no scale phase.
    // Rows are brought into pivot order afterwards, one warp per column.
    {
        constexpr int RT = 128, NCG = 2, RQ = 2;   // 128 row slots x 2 column groups
        const int rslot = tid & (RT - 1), cg = tid >> 7;
        uint32_t used = 0;          // bit b: row lane + 32 b has been a pivot row
        double lprev[RQ] = {0.0, 0.0};
        for (int k = 0; k < r; ++k) {
            const double* ck = S + (size_t)k * ld;
            double bv = -1.0;
            int bi = 0x7fffffff;
            for (int i = lane, b = 0; i < r; i += 32, ++b) {
                if ((used >> b) & 1u) continue;
                const double a = fabs(ck[i]);
                if (a > bv) { bv = a; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            const int p = bi;
            if (tid == 0) {
                pv[k] = p;
                if (!(bv > 0.0) && s_info == 0) s_info = k + 1;
            }
            const double pvv = ck[p];
            const double* rowp = S + p;
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const int i = rslot + RT * q;
                const bool live = i < r && !((used >> (i >> 5)) & 1u) && i != p;
                // multipliers of the previous column (deferred store, see above)
                if (cg == 0 && k > 0 && i < r && !((used >> (i >> 5)) & 1u) )
                    S[i + (size_t)(k - 1) * ld] = lprev[q];
                if (!live) { lprev[q] = 0.0; continue; }
                const double li = (pvv != 0.0) ? ck[i] / pvv : ck[i];
                lprev[q] = li;
                double* rowi = S + i;
                int j = k + 1 + cg;
                for (; j + 3 * NCG < r; j += 4 * NCG) {
                    const size_t o0 = (size_t)j * ld, o1 = o0 + (size_t)NCG * ld,
                                 o2 = o1 + (size_t)NCG * ld, o3 = o2 + (size_t)NCG * ld;
                    const double u0 = rowp[o0], u1 = rowp[o1], u2 = rowp[o2], u3 = rowp[o3];
                    const double a0 = rowi[o0], a1 = rowi[o1], a2 = rowi[o2], a3 = rowi[o3];
                    rowi[o0] = fma(-li, u0, a0);
                    rowi[o1] = fma(-li, u1, a1);
                    rowi[o2] = fma(-li, u2, a2);
                    rowi[o3] = fma(-li, u3, a3);
                }
                for (; j < r; j += NCG) {
                    const size_t o0 = (size_t)j * ld;
                    rowi[o0] = fma(-li, rowp[o0], rowi[o0]);
                }
            }
            if ((p & 31) == lane) used |= 1u << (p >> 5);
            __syncthreads();
        }
        // rows into pivot order: S[m, j] <- S[pv[m], j]; a column belongs to one warp
        for (int j = wid; j < r; j += 8) {
            double* cj = S + (size_t)j * ld;
            double t[6];
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int m = lane + 32 * b;
                t[b] = (m < r) ? cj[pv[m]] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int m = lane + 32 * b;
                if (m < r) cj[m] = t[b];
            }
        }
    }
    __syncthreads();
    if (tid == 0) d.info[sys] = s_info;

    // ---- right-hand sides, DD_RB at a time
    const int64_t q0 = d.sys_rhs_off[sys];
    const int nq = d.sys_rhs_cnt[sys];
    const int nk = n + d.n_border;
    for (int base = 0; base < nq; base += DD_RB) {
        const int nb = min(DD_RB, nq - base);
        if (wid < nb && r > 0) {
            // y = S^-1 u[Mi] : one warp per right-hand side, warp-synchronous
            double* y = ys + (size_t)wid * ld;
            const double* u = d.ut + (int64_t)d.rhs_urow[q0 + base + wid] * M;
            for (int i = lane; i < r; i += 32) y[i] = u[mi[pv[i]]];   // rows in pivot order
            __syncwarp();
            for (int k = 0; k < r - 1; ++k) {
                const double xk = y[k];
                const double* ck = S + (size_t)k * ld;
                for (int i = k + 1 + lane; i < r; i += 32) y[i] = fma(-ck[i], xk, y[i]);
                __syncwarp();
            }
            for (int k = r - 1; k >= 0; --k) {
                const double* ck = S + (size_t)k * ld;
                const double xk = y[k] / ck[k];
                __syncwarp();
                if (lane == 0) y[k] = xk;
                for (int i = lane; i < k; i += 32) y[i] = fma(-ck[i], xk, y[i]);
                __syncwarp();
            }
        }
        __syncthreads();
        // c_K = u_K - G[K, Mi] y  for the nb right-hand sides at once
        for (int i = tid; i < nk; i += 256) {
            const int ki = (i < n) ? stn[i] : (d.n_stn + (i - n));
            double acc[DD_RB];
#pragma unroll
            for (int w = 0; w < DD_RB; ++w)
                acc[w] = (w < nb) ? d.ut[(int64_t)d.rhs_urow[q0 + base + w] * M + ki] : 0.0;
            for (int j = 0; j < r; ++j) {
                const double gv = G[(int64_t)mi[j] * M + ki];
#pragma unroll
                for (int w = 0; w < DD_RB; ++w) acc[w] = fma(-gv, ys[(size_t)w * ld + j], acc[w]);
            }
#pragma unroll
            for (int w = 0; w < DD_RB; ++w) {
                if (w >= nb) break;
                const int64_t q = q0 + base + w;
                const int64_t row = d.rhs_row[q];
                if (row >= 0)
                    d.coef[d.coef_row_major ? row * (int64_t)d.kpad + ki
                                            : coef_offset(row, ki, d.kpad)] = acc[w];
                if (d.rhs_kind[q] == 1)
                    atomicAdd(&d.resid[q], fabs(acc[w] - ((i == n) ? 1.0 : 0.0)));
            }
        }
        __syncthreads();
    }
}


// Register-resident variant: S = G[Mi, Mi] never touches shared memory.  The CTA is a
// 16 x 32 thread grid (warp = row class, lane = column class, both cyclic), thread
// (w, l) keeps S[w + 16 a, l + 32 b] in registers.  S is inverted in place by
// Gauss-Jordan elimination with implicit row pivoting: rows never move (they could
// not, in registers); a step publishes column k and the scaled pivot row through two
// small shared buffers and every thread updates its TA x TB tile -- all rows, all
// columns, so that no substitution phase is left:  y = S^-1 u_Mi  is a product with
// the tile.  With rows in place the result is phys[p_k, j] = S^-1[k, p_j]  (p_k = pivot
// row of column k), hence  y_k = sum_j phys[p_k, j] u_Mi[p_j].
// Two block barriers per column, ~TA + TB shared loads and TA * TB DFMA per thread;
// the LU-in-shared-memory kernel above spends its time on shared-memory wavefronts
// (r^3 / 3 elements read + written) and on barriers instead.
constexpr int DD_RPMAX = 160;   // largest padded order of the register kernel

struct DdShared {
    double colbuf[2][DD_RPMAX];
    double rowbuf[2][DD_RPMAX];
    double ys[DD_RB][DD_RPMAX];
    double bp[DD_RB][DD_RPMAX];
    int mi[DD_RPMAX];
    int pv[DD_RPMAX];               // pivot row of column k
    int rk[DD_RPMAX];               // column whose pivot row is i
    int s_info;
    int s_fail;
    double s_sgn;
};

// Scale row k1 (slot A1 of warp k1 % 16) by 1 / pivot, publish it and column k1 (register
// column kb1 = A1 / 2 of lane k1 % 32 in every warp) into buffer `buf`; the column is
// replaced by zeros except for the pivot, which becomes 1 / pivot (in-place inverse).
template <int TA, int TB, int A1>
__device__ __forceinline__ void dd_publish(double (&S)[TA][TB], DdShared& sm, int k1, int buf,
                                           int wid, int lane) {
    constexpr int kb1 = A1 >> 1;
    const int kl1 = k1 & 31;
    const bool prow = wid == (k1 & 15);
    if (prow) {
        const double pvv = __shfl_sync(0xffffffffu, S[A1][kb1], kl1);
        const double inv = (pvv != 0.0) ? 1.0 / pvv : 1.0;
        if (lane == 0) {
            if (k1 == 0) sm.s_sgn = pvv;
            const double sg = (k1 == 0) ? pvv : sm.s_sgn;
            if (!(pvv * sg > 0.0) || !(fabs(pvv) < 1.0e300)) sm.s_fail = 1;
        }
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            const double v = (b == kb1 && lane == kl1) ? inv : S[A1][b] * inv;
            S[A1][b] = v;
            sm.rowbuf[buf][lane + 32 * b] = v;
        }
    }
    if (lane == kl1) {
#pragma unroll
        for (int a = 0; a < TA; ++a) {
            const bool piv = prow && a == A1;
            sm.colbuf[buf][wid + 16 * a] = piv ? 0.0 : S[a][kb1];
            if (!piv) S[a][kb1] = 0.0;
        }
    }
}

// Elimination step k with look-ahead (A1 = register slot of row k + 1, -1 = none).
template <int TA, int TB, int A1>
__device__ __forceinline__ void dd_step(double (&S)[TA][TB], DdShared& sm, int k, int r, int wid,
                                        int lane) {
    const int cur = k & 1;
    double u[TB], f[TA];
#pragma unroll
    for (int b = 0; b < TB; ++b) u[b] = sm.rowbuf[cur][lane + 32 * b];
#pragma unroll
    for (int a = 0; a < TA; ++a) f[a] = sm.colbuf[cur][wid + 16 * a];
    if (A1 >= 0) {
        constexpr int A1c = A1 >= 0 ? A1 : 0;
        constexpr int kb1 = A1c >> 1;
        const int k1 = k + 1;
        const bool nrow = wid == (k1 & 15);
#pragma unroll
        for (int a = 0; a < TA; ++a) S[a][kb1] = fma(-f[a], u[kb1], S[a][kb1]);
        if (nrow) {
#pragma unroll
            for (int b = 0; b < TB; ++b)
                if (b != kb1) S[A1c][b] = fma(-f[A1c], u[b], S[A1c][b]);
        }
        if (k1 < r) dd_publish<TA, TB, A1c>(S, sm, k1, cur ^ 1, wid, lane);
#pragma unroll
        for (int a = 0; a < TA; ++a) {
            if (a == A1c && nrow) continue;
#pragma unroll
            for (int b = 0; b < TB; ++b)
                if (b != kb1) S[a][b] = fma(-f[a], u[b], S[a][b]);
        }
    } else {
#pragma unroll
        for (int a = 0; a < TA; ++a)
#pragma unroll
            for (int b = 0; b < TB; ++b) S[a][b] = fma(-f[a], u[b], S[a][b]);
    }
    __syncthreads();
}

// Columns 16 A0 .. 16 A0 + 15 (pivot rows in register slot A0), then the next slot.
template <int TA, int TB, int A0>
__device__ __forceinline__ void dd_sweep(double (&S)[TA][TB], DdShared& sm, int r, int wid,
                                         int lane) {
    if constexpr (A0 < TA) {
#pragma unroll 1
        for (int w = 0; w < 15; ++w) {           // rows 16 A0 + w + 1 share the slot A0
            const int k = 16 * A0 + w;
            if (k >= r) break;
            dd_step<TA, TB, A0>(S, sm, k, r, wid, lane);
        }
        const int k = 16 * A0 + 15;              // the next row lives in slot A0 + 1
        if (k < r) dd_step<TA, TB, (A0 + 1 < TA ? A0 + 1 : -1)>(S, sm, k, r, wid, lane);
        dd_sweep<TA, TB, A0 + 1>(S, sm, r, wid, lane);
    }
}

template <int TA, int TB>
__device__ __forceinline__ void downdate_reg_body(const spx_downdate& d, DdShared& sm, int sys,
                                                  int r, int force_pivot) {
    constexpr int RP = 32 * TB;          // padded order (columns); rows: 16 * TA <= RP
    static_assert(16 * TA <= RP && RP <= DD_RPMAX, "tile shape");
    const int n = d.sys_n[sys];
    const int M = d.n_stn + d.n_border;
    auto& colbuf = sm.colbuf;
    auto& rowbuf = sm.rowbuf;
    auto& ys = sm.ys;
    auto& bp = sm.bp;
    auto& mi = sm.mi;
    auto& pv = sm.pv;
    auto& rk = sm.rk;
    int& s_info = sm.s_info;
    int& s_fail = sm.s_fail;
    double& s_sgn = sm.s_sgn;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* __restrict__ miss = d.miss_list + d.sys_miss_off[sys];
    const int32_t* __restrict__ stn = d.stn_list + d.sys_stn_off[sys];
    const double* __restrict__ G = d.ginv;
    if (tid == 0) s_info = 0;
    for (int i = tid; i < RP; i += 512) mi[i] = (i < r) ? miss[i] : 0;
    __syncthreads();

    double S[TA][TB];
    auto load_S = [&]() {
#pragma unroll
        for (int a = 0; a < TA; ++a) {
            const int i = wid + 16 * a;
#pragma unroll
            for (int b = 0; b < TB; ++b) {
                const int j = lane + 32 * b;
                S[a][b] = (i < r && j < r) ? G[(int64_t)mi[i] * M + mi[j]] : 0.0;
            }
        }
    };
    load_S();

    // ---- fast path: no pivoting.  For a valid variogram S is a principal submatrix of
    // the inverse of a (conditionally) definite system and is itself definite, so the
    // diagonal pivots are safe and known in advance: the owners of row k and column k
    // publish them at the top of step k straight from their registers, ONE barrier per
    // column, every tile index a compile-time constant (outer loops unrolled over the
    // register slots).  A pivot that is zero, not finite or of the wrong sign (S not
    // definite) abandons the result and the pivoted elimination below runs instead.
    if (tid < RP) { pv[tid] = tid; rk[tid] = tid; }
    if (tid == 0) s_fail = force_pivot;
    __syncthreads();
    if (!force_pivot) {
        // Look-ahead: step k first brings row k + 1 and column k + 1 up to date, scales /
        // publishes them for the next step (shuffle, division, shared-memory stores: the
        // latency chain of the elimination) and only then applies the bulk of its rank-1
        // update, so that the chain of step k + 1 hides behind the DFMAs of step k.  Every
        // element still receives the same FMAs in the same order as without look-ahead.
        dd_publish<TA, TB, 0>(S, sm, 0, 0, wid, lane);
        __syncthreads();
        dd_sweep<TA, TB, 0>(S, sm, r, wid, lane);
    }
    const bool pivoted = s_fail != 0;
    if (pivoted) {
    __syncthreads();
    load_S();
    uint32_t used = 0;   // bit q: row lane + 32 q has been a pivot row (search bookkeeping)
    // The column loop is split by register slot kb = k / 32 (unrolled), so that the tile
    // index of column k is a compile-time constant inside: no selects over registers.
#pragma unroll
    for (int kb = 0; kb < TB; ++kb) {
        const int kend = min(r, 32 * (kb + 1));
        for (int k = 32 * kb; k < kend; ++k) {
            const int cur = k & 1;
            const bool kcol = lane == (k & 31);
            // (1) publish column k and replace it by zeros: with the 1 / pivot put into
            // the pivot row below, the column becomes e_p * (1 / pivot) and the uniform
            // row operation of (3) turns it into column k of the in-place inverse
            if (kcol) {
#pragma unroll
                for (int a = 0; a < TA; ++a) {
                    colbuf[cur][wid + 16 * a] = S[a][kb];
                    S[a][kb] = 0.0;
                }
            }
            __syncthreads();
            // (2) every warp finds the pivot row for itself.  The comparison key is the
            // high word of |v| (sign stripped): monotone in |v|, 20 mantissa bits -- any
            // element within 1e-6 of the largest is as good a pivot -- and it turns the
            // warp arg-max into one REDUX + one ballot.
            unsigned key = 0;
            int bi = 0;
#pragma unroll
            for (int q = 0; q < TB; ++q) {
                const int i = lane + 32 * q;
                if (i < r && !((used >> q) & 1u)) {
                    const unsigned kk =
                        ((unsigned)__double2hiint(colbuf[cur][i]) & 0x7fffffffu) + 1u;
                    if (kk > key) { key = kk; bi = i; }
                }
            }
            const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
            const unsigned bal = __ballot_sync(0xffffffffu, key == kmax);
            const int p = __shfl_sync(0xffffffffu, bi, __ffs(bal) - 1);
            if ((p & 31) == lane) used |= 1u << (p >> 5);
            if (wid == (p & 15)) {       // the warp that owns the pivot row scales it
                const double pvv = colbuf[cur][p];
                const double inv = (pvv != 0.0) ? 1.0 / pvv : 1.0;
                const int a0 = p >> 4;
                if (lane == 0) {
                    pv[k] = p;
                    rk[p] = k;
                    if (!(fabs(pvv) > 0.0) && s_info == 0) s_info = k + 1;
                }
#pragma unroll
                for (int a = 0; a < TA; ++a)
                    if (a == a0) {
#pragma unroll
                        for (int b = 0; b < TB; ++b) {
                            const double v = (b == kb && kcol) ? inv : S[a][b] * inv;
                            S[a][b] = v;
                            rowbuf[cur][lane + 32 * b] = v;
                        }
                    }
            }
            __syncthreads();
            // (3) eliminate column k from every other row
            double u[TB];
#pragma unroll
            for (int b = 0; b < TB; ++b) u[b] = rowbuf[cur][lane + 32 * b];
#pragma unroll
            for (int a = 0; a < TA; ++a) {
                const int i = wid + 16 * a;
                const double f = (i == p) ? 0.0 : colbuf[cur][i];
#pragma unroll
                for (int b = 0; b < TB; ++b) S[a][b] = fma(-f, u[b], S[a][b]);
            }
            // the buffers of parity cur are rewritten two barriers from here: safe
        }
    }
    }   // pivoted
    __syncthreads();
    if (tid == 0) d.info[sys] = s_info;

    // ---- right-hand sides, DD_RB at a time
    const int64_t q0 = d.sys_rhs_off[sys];
    const int nq = d.sys_rhs_cnt[sys];
    const int nk = n + d.n_border;
    for (int base = 0; base < nq; base += DD_RB) {
        const int nb = min(DD_RB, nq - base);
        for (int idx = tid; idx < nb * RP; idx += 512) {
            const int w = idx / RP, j = idx - w * RP;
            bp[w][j] = (j < r) ? d.ut[(int64_t)d.rhs_urow[q0 + base + w] * M + mi[pv[j]]] : 0.0;
        }
        __syncthreads();
        for (int w = 0; w < nb; ++w) {
            double bw[TB];
#pragma unroll
            for (int b = 0; b < TB; ++b) bw[b] = bp[w][lane + 32 * b];
#pragma unroll
            for (int a = 0; a < TA; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < TB; ++b) acc = fma(S[a][b], bw[b], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                const int i = wid + 16 * a;
                if (lane == 0 && i < r) ys[w][rk[i]] = acc;
            }
        }
        __syncthreads();
        // c_K = u_K - G[K, Mi] y  for the nb right-hand sides at once
        const bool want_base = d.base != nullptr;
        if (want_base && lane == 0) {
#pragma unroll
            for (int w = 0; w < DD_RB; ++w) bp[w][wid] = 0.0;   // bp is free here (see above)
        }
        for (int i0 = 0; i0 < nk; i0 += 512) {
            const int i = i0 + tid;
            const bool act = i < nk;
            const int ki = !act ? 0 : ((i < n) ? stn[i] : (d.n_stn + (i - n)));
            double acc[DD_RB];
#pragma unroll
            for (int w = 0; w < DD_RB; ++w)
                acc[w] = (act && w < nb) ? d.ut[(int64_t)d.rhs_urow[q0 + base + w] * M + ki] : 0.0;
            if (act) {
#pragma unroll 4
                for (int j = 0; j < r; ++j) {
                    const double gv = G[(int64_t)mi[j] * M + ki];
#pragma unroll
                    for (int w = 0; w < DD_RB; ++w) acc[w] = fma(-gv, ys[w][j], acc[w]);
                }
#pragma unroll
                for (int w = 0; w < DD_RB; ++w) {
                    if (w >= nb) break;
                    const int64_t q = q0 + base + w;
                    const int64_t row = d.rhs_row[q];
                    if (row >= 0) {
                        d.coef[d.coef_row_major ? row * (int64_t)d.kpad + ki
                                                : coef_offset(row, ki, d.kpad)] = acc[w];
                        if (d.coef_t) d.coef_t[(int64_t)ki * d.coef_t_ld + row] = acc[w];
                    }
                    if (d.rhs_kind[q] == 1)
                        atomicAdd(&d.resid[q], fabs(acc[w] - ((i == n) ? 1.0 : 0.0)));
                }
            }
            if (want_base) {
                // local estimator: base = F * sum of the station coefficients + the
                // coefficient of the ones-border (drift coefficients are not part of it);
                // fixed-order reduction: lanes by shuffle, warps below
                const double bf = !act ? 0.0 : ((i < n) ? d.base_f : ((i == n) ? 1.0 : 0.0));
#pragma unroll
                for (int w = 0; w < DD_RB; ++w) {
                    double v = bf * acc[w];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) bp[w][wid] += v;
                }
            }
        }
        if (d.coef_t) {           // missing stations: explicit zeros (the buffer is not pre-set)
            for (int idx = tid; idx < nb * r; idx += 512) {
                const int w = idx / r, j = idx - w * r;
                const int64_t row = d.rhs_row[q0 + base + w];
                if (row >= 0) d.coef_t[(int64_t)mi[j] * d.coef_t_ld + row] = 0.0;
            }
        }
        if (want_base) {
            __syncthreads();
            if (tid < nb) {
                const int64_t row = d.rhs_row[q0 + base + tid];
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) v += bp[tid][k];
                if (row >= 0) d.base[row] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Blocked LU variant of the downdate (default for r <= DD_REG_MAX).
//
// The Gauss-Jordan kernel above pays one block-wide barrier and ~100 issued instructions
// per warp for EVERY column of S (r ~ 100 columns, all 16 warps in lock step).  Here S
// lives in shared memory and is factored in panels of DL_NB columns:
//   * the panel (rows p.., DL_NB columns) is factored by ONE warp in registers -- the
//     pivot row is broadcast by shuffles, no block barrier inside a panel;
//   * S is symmetric: S = L D L^T, the U block row of the LU is D L21^T and needs no
//     phase of its own;
//   * the trailing update  S22 -= L21 D L21^T  (lower tiles only) runs on the FP64 tensor
//     cores (DMMA m8n8k4, two per 8 x 8 tile);
//   * every right-hand side is solved by one warp (vector in registers, shuffles).
// One block barrier per panel instead of one per column, and two blocks per SM
// (r <= 112) overlap each other's latency chains.  No pivoting: S is a principal
// submatrix of the inverse of a (conditionally) definite system, hence definite; a zero,
// non-finite or wrong-sign pivot flags the system and k_downdate_reg redoes it with
// pivoting (repair pass).
constexpr int DL_NB = 8;
constexpr int DL_THREADS = 256;
constexpr int DL_SMALL = 112;    // r <= DL_SMALL: two blocks per SM

static size_t dl_smem_bytes(int rp) {
    return ((size_t)(rp + 1) * rp + 2 * rp + (size_t)DD_RB * rp) * sizeof(double) +
           (size_t)rp * sizeof(int);
}

template <int NQ>
__global__ void __launch_bounds__(DL_THREADS) k_downdate_lu(spx_downdate d, int r_lo, int r_hi,
                                                           int rp_max, int test_fail) {
    extern __shared__ double dsm[];
    __shared__ int s_info;
    __shared__ double s_part[DD_RB][DL_THREADS / 32];
    const int sys = d.sys_order ? d.sys_order[blockIdx.x] : blockIdx.x;
    const int r = d.sys_r[sys];
    if (r <= r_lo || r > r_hi) return;
    const int n = d.sys_n[sys];
    const int M = d.n_stn + d.n_border;
    const int ld = rp_max + 1;                         // odd pitch
    double* __restrict__ S = dsm;                      // [ld * rp_max] column-major
    double* __restrict__ rdiag = S + (size_t)ld * rp_max;   // 1 / d_j
    double* __restrict__ dg = rdiag + rp_max;          // d_j (pivots of S = L D L^T)
    double* __restrict__ ys = dg + rp_max;             // [DD_RB][rp_max]
    int* __restrict__ mi = reinterpret_cast<int*>(ys + (size_t)DD_RB * rp_max);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* __restrict__ miss = d.miss_list + d.sys_miss_off[sys];
    const int32_t* __restrict__ stn = d.stn_list + d.sys_stn_off[sys];
    const double* __restrict__ G = d.ginv;
    if (test_fail && (sys & 1)) {                      // SPX_DD_LU_FAIL=1: exercise the repair pass
        if (tid == 0) d.info[sys] = 1;
        return;
    }
    if (tid == 0) s_info = 0;
    for (int i = tid; i < r; i += DL_THREADS) mi[i] = miss[i];
    __syncthreads();
    const int rp = (r + 7) & ~7;                       // <= rp_max
    // S = G[Mi, Mi] (row mi[j] of G read along ascending mi[i]; G symmetric), identity pad
    for (int idx = tid; idx < rp * rp; idx += DL_THREADS) {
        const int j = idx / rp, i = idx - j * rp;
        S[i + (size_t)j * ld] = (i < r && j < r) ? G[(int64_t)mi[j] * M + mi[i]]
                                                 : ((i == j) ? 1.0 : 0.0);
    }
    __syncthreads();

    double sgn = 0.0;                                  // sign of the first pivot (warp 0)
    // ---- panel: rows p + lane + 32 q, columns p .. p + nb - 1, factored by ONE warp in
    // registers (pivot row broadcast by shuffles)
    auto panel = [&](int p) {
        const int nb = min(DL_NB, r - p);
        double a[NQ][DL_NB];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = p + lane + 32 * q;
#pragma unroll
            for (int c = 0; c < DL_NB; ++c)
                a[q][c] = (i < r && c < nb) ? S[i + (size_t)(p + c) * ld] : 0.0;
        }
        int bad = 0;
#pragma unroll
        for (int jc = 0; jc < DL_NB; ++jc) {
            if (jc < nb) {
                const double piv = __shfl_sync(0xffffffffu, a[0][jc], jc);
                if (p == 0 && jc == 0) sgn = piv;
                if (!(piv * sgn > 0.0) || !(fabs(piv) < 1.0e300)) bad = bad ? bad : p + jc + 1;
                const double rcp = (piv != 0.0) ? 1.0 / piv : 1.0;
                if (lane == 0) {
                    rdiag[p + jc] = rcp;
                    dg[p + jc] = piv;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (lane + 32 * q > jc) a[q][jc] *= rcp;
#pragma unroll
                for (int c = jc + 1; c < DL_NB; ++c) {
                    const double u = __shfl_sync(0xffffffffu, a[0][c], jc);
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        if (lane + 32 * q > jc) a[q][c] = fma(-a[q][jc], u, a[q][c]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = p + lane + 32 * q;
#pragma unroll
            for (int c = 0; c < DL_NB; ++c)
                if (i < r && c < nb) S[i + (size_t)(p + c) * ld] = a[q][c];
        }
        if (bad && lane == 0 && s_info == 0) s_info = bad;
    };
    // 8 x 8 tile of the trailing update  S22 -= L21 D L21^T  (two DMMA m8n8k4).  S is
    // symmetric, so the U block row of the LU is D L21^T: both operands come from the panel
    // columns and only the lower tiles (and the diagonal ones, whose upper triangle the
    // next panel reads) are kept up to date.
    const int fg = lane >> 2, ft = lane & 3;
    auto tile_update = [&](int p, int i0, int c0) {
        double* cp0 = S + i0 + fg + (size_t)(c0 + 2 * ft) * ld;
        double c_lo = cp0[0], c_hi = cp0[ld];
        const double a0 = -S[i0 + fg + (size_t)(p + ft) * ld];
        const double a1 = -S[i0 + fg + (size_t)(p + 4 + ft) * ld];
        const double b0 = S[c0 + fg + (size_t)(p + ft) * ld] * dg[p + ft];
        const double b1 = S[c0 + fg + (size_t)(p + 4 + ft) * ld] * dg[p + 4 + ft];
        dmma_lu(c_lo, c_hi, a0, b0);
        dmma_lu(c_lo, c_hi, a1, b1);
        cp0[0] = c_lo;
        cp0[ld] = c_hi;
    };
    if (wid == 0 && r > 0) panel(0);
    __syncthreads();
    for (int p = 0; p + DL_NB < r; p += DL_NB) {
        // ---- trailing update with look-ahead: warp 0 updates the columns of the next
        // panel and factors it right away while the other warps update the remaining
        // lower tiles; ONE block barrier per panel
        const int o = p + DL_NB;
        const int nt = (rp - o) >> 3;
        if (wid == 0) {
            for (int ti = 0; ti < nt; ++ti) tile_update(p, o + 8 * ti, o);
            __syncwarp();
            panel(o);
        } else {
            int cnt = 0;
            for (int ti = 1; ti < nt; ++ti)
                for (int tj = 1; tj <= ti; ++tj, ++cnt)
                    if (cnt % (DL_THREADS / 32 - 1) == wid - 1)
                        tile_update(p, o + 8 * ti, o + 8 * tj);
        }
        __syncthreads();
    }
    if (s_info != 0) {                                 // block-uniform: left to the repair pass
        if (tid == 0) d.info[sys] = s_info;
        return;
    }
    if (tid == 0) d.info[sys] = 0;

    // ---- right-hand sides, DD_RB at a time: y = S^-1 u_Mi, one warp per right-hand side
    const int64_t q0 = d.sys_rhs_off[sys];
    const int nq = d.sys_rhs_cnt[sys];
    const int nk = n + d.n_border;
    for (int base = 0; base < nq; base += DD_RB) {
        const int nbq = min(DD_RB, nq - base);
        if (wid < nbq && r > 0) {
            const double* __restrict__ urow = d.ut + (int64_t)d.rhs_urow[q0 + base + wid] * M;
            double x[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = lane + 32 * q;
                x[q] = (i < r) ? urow[mi[i]] : 0.0;
            }
            // L y' = b (unit diagonal)
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq) {
#pragma unroll 4
                for (int jl = 0; jl < 32; ++jl) {
                    const int j = 32 * qq + jl;
                    if (j >= r) break;
                    const double xj = __shfl_sync(0xffffffffu, x[qq], jl);
                    const double* __restrict__ col = S + (size_t)j * ld;
#pragma unroll
                    for (int q = qq; q < NQ; ++q) {
                        const int i = lane + 32 * q;
                        if (i > j && i < r) x[q] = fma(-col[i], xj, x[q]);
                    }
                }
            }
            // D z = y'
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = lane + 32 * q;
                if (i < r) x[q] *= rdiag[i];
            }
            // L^T y = z  (row j of L: S[j + i * ld], i < j)
#pragma unroll
            for (int qq = NQ - 1; qq >= 0; --qq) {
#pragma unroll 4
                for (int jl = 31; jl >= 0; --jl) {
                    const int j = 32 * qq + jl;
                    if (j >= r) continue;
                    const double xj = __shfl_sync(0xffffffffu, x[qq], jl);
#pragma unroll
                    for (int q = 0; q <= qq; ++q) {
                        const int i = lane + 32 * q;
                        if (i < j) x[q] = fma(-S[j + (size_t)i * ld], xj, x[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = lane + 32 * q;
                if (i < r) ys[(size_t)wid * rp_max + i] = x[q];
            }
        }
        __syncthreads();
        // c_K = u_K - G[K, Mi] y  for the nbq right-hand sides at once (two accumulators
        // when the system has a single step: its data and the ones-vector)
        const bool want_base = d.base != nullptr;
        if (want_base && lane == 0) {
#pragma unroll
            for (int w = 0; w < DD_RB; ++w) s_part[w][wid] = 0.0;
        }
        auto emit = [&](auto wtag) {
            constexpr int W = decltype(wtag)::value;
            for (int i0 = 0; i0 < nk; i0 += DL_THREADS) {
                const int i = i0 + tid;
                const bool act = i < nk;
                const int ki = !act ? 0 : ((i < n) ? stn[i] : (d.n_stn + (i - n)));
                double acc[W];
#pragma unroll
                for (int w = 0; w < W; ++w)
                    acc[w] = (act && w < nbq)
                                 ? d.ut[(int64_t)d.rhs_urow[q0 + base + w] * M + ki] : 0.0;
                if (act) {
#pragma unroll 4
                    for (int j = 0; j < r; ++j) {
                        const double gv = G[(int64_t)mi[j] * M + ki];
#pragma unroll
                        for (int w = 0; w < W; ++w)
                            acc[w] = fma(-gv, ys[(size_t)w * rp_max + j], acc[w]);
                    }
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        if (w >= nbq) break;
                        const int64_t q = q0 + base + w;
                        const int64_t row = d.rhs_row[q];
                        if (row >= 0) {
                            d.coef[d.coef_row_major ? row * (int64_t)d.kpad + ki
                                                    : coef_offset(row, ki, d.kpad)] = acc[w];
                            if (d.coef_t) d.coef_t[(int64_t)ki * d.coef_t_ld + row] = acc[w];
                        }
                        if (d.rhs_kind[q] == 1)
                            atomicAdd(&d.resid[q], fabs(acc[w] - ((i == n) ? 1.0 : 0.0)));
                    }
                }
                if (want_base) {
                    const double bf = !act ? 0.0 : ((i < n) ? d.base_f : ((i == n) ? 1.0 : 0.0));
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        double v = bf * acc[w];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0) s_part[w][wid] += v;
                    }
                }
            }
        };
        if (nbq <= 2) emit(std::integral_constant<int, 2>{});
        else emit(std::integral_constant<int, DD_RB>{});
        if (d.coef_t) {           // missing stations: explicit zeros (the buffer is not pre-set)
            for (int idx = tid; idx < nbq * r; idx += DL_THREADS) {
                const int w = idx / r, j = idx - w * r;
                const int64_t row = d.rhs_row[q0 + base + w];
                if (row >= 0) d.coef_t[(int64_t)mi[j] * d.coef_t_ld + row] = 0.0;
            }
        }
        __syncthreads();
        if (want_base && tid < nbq) {
            const int64_t row = d.rhs_row[q0 + base + tid];
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < DL_THREADS / 32; ++k) v += s_part[tid][k];
            if (row >= 0) d.base[row] = v;
        }
        __syncthreads();
    }
}

// One launch for every tile shape: systems of different order run side by side (three
// launches, one per shape, serialise behind each other's stragglers).
__global__ void __launch_bounds__(512, 1) k_downdate_reg(spx_downdate d, int r_max, int force_pivot,
                                                        int only_failed) {
    __shared__ DdShared sm;
    const int sys = d.sys_order ? d.sys_order[blockIdx.x] : blockIdx.x;
    const int r = d.sys_r[sys];
    if (r > r_max) return;               // left to the shared-memory kernel
    // repair pass behind k_downdate_lu: only the systems it flagged (nothing written yet)
    if (only_failed && d.info[sys] == 0) return;
    if (r <= 112) downdate_reg_body<7, 4>(d, sm, sys, r, force_pivot);
    else if (r <= 128) downdate_reg_body<8, 4>(d, sm, sys, r, force_pivot);
    else downdate_reg_body<10, 5>(d, sm, sys, r, force_pivot);
}

}  // namespace spx

using namespace spx;

extern "C" {

int64_t spx_coef_offset(int64_t row, int64_t col, int64_t kpad) {
    return coef_offset(row, col, kpad);
}

int spx_krige_assemble_dev(const spx_systems* s, const spx_vg* vgs_dev, int n_vgs,
                           double min_vg_val, void* stream) {
    if (!s || !vgs_dev || n_vgs <= 0) {
        set_error("krige_assemble: null argument");
        return SPX_EINVAL;
    }
    if (s->n_sys == 0) return SPX_OK;
    if (s->n_sys > 65535) {
        set_error("krige_assemble: %d systems in one call (max 65535)", s->n_sys);
        return SPX_EINVAL;
    }
    // enough x-blocks to cover a 1024 x 1024 system in ~4 passes
    dim3 grid(s->n_sys >= 148 ? 8 : 64, s->n_sys);
    k_assemble<<<grid, 256, 0, (cudaStream_t)stream>>>(*s, vgs_dev, min_vg_val);
    SPX_CHECK_LAUNCH("k_assemble");
    return SPX_OK;
}

int spx_krige_factor_dev(const spx_systems* s, void* stream) {
    if (!s) {
        set_error("krige_factor: null argument");
        return SPX_EINVAL;
    }
    if (s->n_sys == 0) return SPX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, max_smem = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (s->max_m > 0) {
        // panel pitch: >= rows rounded up to the 32-row MMA tiles, == 4 (mod 16)
        const int ldp = ((s->max_m + 31) / 32) * 32 + 32 + 4;
        const int nbs[3] = {32, 16, 8};
        for (int nb : nbs) {
            const size_t smem = ((size_t)nb * ldp + (size_t)nb * LU_UP) * sizeof(double);
            if (smem + 1024 > (size_t)max_smem) continue;
            if (nb == 32) {
                SPX_CUDA(cudaFuncSetAttribute(k_lu_blocked<32>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
                k_lu_blocked<32><<<s->n_sys, 256, smem, st>>>(*s, ldp);
            } else if (nb == 16) {
                SPX_CUDA(cudaFuncSetAttribute(k_lu_blocked<16>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
                k_lu_blocked<16><<<s->n_sys, 256, smem, st>>>(*s, ldp);
            } else {
                SPX_CUDA(cudaFuncSetAttribute(k_lu_blocked<8>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
                k_lu_blocked<8><<<s->n_sys, 256, smem, st>>>(*s, ldp);
            }
            SPX_CHECK_LAUNCH("k_lu_blocked");
            return SPX_OK;
        }
    }
    k_lu_factor<<<s->n_sys, 256, 0, st>>>(*s);  // unblocked fallback (any size)
    SPX_CHECK_LAUNCH("k_lu_factor");
    return SPX_OK;
}

int spx_krige_solve_dev(const spx_systems* s, const spx_rhs* r, void* stream) {
    if (!s || !r) {
        set_error("krige_solve: null argument");
        return SPX_EINVAL;
    }
    if (r->n_rhs == 0) return SPX_OK;
    if (r->kpad % 4 != 0) {
        set_error("krige_solve: kpad must be a multiple of 4");
        return SPX_EINVAL;
    }
    // shared memory: the longest system; the caller guarantees m <= n_stn + border
    const int max_m = r->n_stn + 1 + s->n_drifts;
    const size_t smem = (size_t)max_m * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("krige_solve: system size %d exceeds the shared-memory vector limit", max_m);
        return SPX_ENOMEM;
    }
    if (smem > 48 * 1024)
        SPX_CUDA(cudaFuncSetAttribute(k_lu_solve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    k_lu_solve<<<r->n_rhs, 256, smem, (cudaStream_t)stream>>>(*s, *r);
    SPX_CHECK_LAUNCH("k_lu_solve");
    return SPX_OK;
}

static size_t dd_smem_bytes(int max_r) {
    const size_t ld = (size_t)(max_r | 1);
    return (ld * max_r + (size_t)DD_RB * ld) * sizeof(double) + 2 * (size_t)max_r * sizeof(int);
}

// r <= DD_REG_MAX: register-resident Gauss-Jordan (k_downdate_reg); larger r (up to what
// shared memory holds): LU in shared memory (k_downdate).  SPX_DD_SMEM=1 forces the latter.
constexpr int DD_REG_MAX = 160;

static bool dd_force_smem() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SPX_DD_SMEM");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// SPX_DD_LU=0 selects the register-resident Gauss-Jordan kernel instead of the blocked LU.
static bool dd_use_lu() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SPX_DD_LU");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static int dd_force_pivot() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SPX_DD_PIVOT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

int spx_krige_downdate_reg_max_r(void) { return DD_REG_MAX; }

int spx_krige_downdate_max_r(void) {
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) !=
            cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int r = 0;
    while (dd_smem_bytes(r + 8) + 256 <= (size_t)max_smem) r += 8;
    return r;
}

int spx_krige_downdate_dev(const spx_downdate* d, void* stream) {
    if (!d) {
        set_error("krige_downdate: null argument");
        return SPX_EINVAL;
    }
    if (d->n_sys == 0) return SPX_OK;
    if (d->kpad % 4 != 0 || d->max_r < 0) {
        set_error("krige_downdate: bad kpad / max_r");
        return SPX_EINVAL;
    }
    spx_downdate dd = *d;
    if (dd.max_r < 1) dd.max_r = 1;
    if ((d->coef_t || d->base) &&
        (d->max_r > DD_REG_MAX || dd_force_smem() || d->n_border < 1 ||
         (d->coef_t && d->coef_t_ld < 1))) {
        set_error("krige_downdate: coef_t / base need the register kernel (max_r <= %d), "
                  "n_border >= 1 and coef_t_ld >= 1", DD_REG_MAX);
        return SPX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int smem_lo = 0;   // systems with r > smem_lo go to the shared-memory kernel
    if (!dd_force_smem()) {
        const int fp = dd_force_pivot();
        if (dd_use_lu() && !fp) {
            // blocked LU in shared memory; large systems first (one block per SM), then
            // the bulk (two blocks per SM); flagged systems are redone with pivoting
            int dev = 0, max_smem = 0;
            SPX_CUDA(cudaGetDevice(&dev));
            SPX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                            dev));
            const size_t sm_big = dl_smem_bytes(DD_REG_MAX);
            size_t sm_small = dl_smem_bytes(DL_SMALL);
            // SPX_DL_SOLO=1: one solve block per SM (shared-memory request above half an
            // SM's), which leaves room for three estimate blocks of the previous chunk
            static const int solo = (getenv("SPX_DL_SOLO") && getenv("SPX_DL_SOLO")[0] == '1');
            if (solo && sm_small < (size_t)116 * 1024) sm_small = (size_t)116 * 1024;
            static const int tf = (getenv("SPX_DD_LU_FAIL") && getenv("SPX_DD_LU_FAIL")[0] == '1');
            if (sm_big + 1024 <= (size_t)max_smem) {
                if (dd.max_r > DL_SMALL) {
                    SPX_CUDA(cudaFuncSetAttribute(k_downdate_lu<5>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)sm_big));
                    k_downdate_lu<5><<<d->n_sys, DL_THREADS, sm_big, st>>>(dd, DL_SMALL, DD_REG_MAX,
                                                                          DD_REG_MAX, tf);
                    SPX_CHECK_LAUNCH("k_downdate_lu<5>");
                }
                SPX_CUDA(cudaFuncSetAttribute(k_downdate_lu<4>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)sm_small));
                k_downdate_lu<4><<<d->n_sys, DL_THREADS, sm_small, st>>>(dd, -1, DL_SMALL, DL_SMALL,
                                                                        tf);
                SPX_CHECK_LAUNCH("k_downdate_lu<4>");
                k_downdate_reg<<<d->n_sys, 512, 0, st>>>(dd, DD_REG_MAX, 1, 1);
                SPX_CHECK_LAUNCH("k_downdate_reg(repair)");
                if (dd.max_r <= DD_REG_MAX) return SPX_OK;
                smem_lo = DD_REG_MAX;
            }
        }
        if (smem_lo == 0) {
            k_downdate_reg<<<d->n_sys, 512, 0, st>>>(dd, DD_REG_MAX, fp, 0);
            SPX_CHECK_LAUNCH("k_downdate_reg");
            if (dd.max_r <= DD_REG_MAX) return SPX_OK;
            smem_lo = DD_REG_MAX;
        }
    }
    const size_t smem = dd_smem_bytes(dd.max_r);
    int dev = 0, max_smem = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem + 256 > (size_t)max_smem) {
        set_error("krige_downdate: max_r=%d needs %zu bytes of shared memory (limit %d)",
                  d->max_r, smem, max_smem);
        return SPX_ENOMEM;
    }
    if (smem > 48 * 1024)
        SPX_CUDA(cudaFuncSetAttribute(k_downdate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    k_downdate<<<d->n_sys, 256, smem, st>>>(dd, smem_lo);
    SPX_CHECK_LAUNCH("k_downdate");
    return SPX_OK;
}

}  // extern "C"
