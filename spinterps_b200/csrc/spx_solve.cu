// Batched kriging systems: assembly from station coordinates, LU factorisation
// with partial pivoting (one thread block per system, warp-shuffle pivot
// search) and multi-right-hand-side solves whose solutions are scattered into
// the packed coefficient matrix consumed by the estimate contraction.
//
// Replaces, per (availability group x variogram x kriging kind):
//   _get_vars_arr_subset   interp/steps.py:170-243   (assembly)
//   np.linalg.pinv         interp/steps.py:351       (factor; non-singular case)
//   np.matmul(inv, rhs_i)  interp/steps.py:416       (dual form)
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
            if (a > bv || !(a == a)) {  // NaN wins so that it surfaces in info
                if (!(bv != bv)) { bv = a; bi = i; }
            }
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
            r.coef[coef_offset(row, col, r.kpad)] = b[i];
        }
    }
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
    k_lu_factor<<<s->n_sys, 256, 0, (cudaStream_t)stream>>>(*s);
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

}  // extern "C"
