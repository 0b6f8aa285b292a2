// One native call per time chunk: the planned fast path of the kriging engine
// (include/spx_b200.h, "one native call per time chunk").
//
// Replaces, for the common case, the per-chunk Python of engine._krige_fast: the host
// part (availability groups, downdate descriptors) runs here without the interpreter, every
// buffer comes from a ring of pre-allocated slots, and the solve phase of a chunk is queued
// on its own high-priority stream so that it runs UNDERNEATH the HBM-bound estimate kernel
// of the previous chunk (interp/steps.py:673-833 is the loop nest this stands for).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <new>

#include "spx_b200.h"
#include "spx_common.cuh"

namespace spx {

// ------------------------------------------------------------------ Ut = Bt . G (DMMA)
// 32 x 64 tile per block (640 blocks for 2500 x 501: one wave, several blocks per SM),
// 8 warps (2 along rows x 4 along columns, 16 x 16 each), K chunks of 32 staged in shared
// memory with pitch 36 (== 4 mod 16: conflict-free 8 x 4 fragment loads); the global loads
// of chunk k + 1 are issued before the DMMA sweep of chunk k.  The A operand (Bt) is never
// stored: it is the resident data block with NaN -> 0 (data rows) or its availability
// mask (one row per system).  G is symmetric, so the col-major B fragment B[k][n] = G[k][n]
// is read as G[n][k]: contiguous along k.
constexpr int UT_BM = 32, UT_BN = 64, UT_BK = 32, UT_LD = UT_BK + 4;
constexpr int UT_EA = (UT_BM * UT_BK) / 256, UT_EB = (UT_BN * UT_BK) / 256;

__device__ __forceinline__ void dmma_ut(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_ut_gemm(const double* __restrict__ data, int n_stn,
                                                 int64_t data_ld,
                                                 const int32_t* __restrict__ src_step,
                                                 int64_t n_rows, int64_t n_data, int M,
                                                 const double* __restrict__ G,
                                                 double* __restrict__ ut) {
    __shared__ double As[UT_BM][UT_LD];
    __shared__ double Bs[UT_BN][UT_LD];
    __shared__ int64_t s_src[UT_BM];     // data row offset of each tile row, < 0 = beyond n_rows
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wr = wid >> 2, wc = wid & 3;
    const int64_t row0 = (int64_t)blockIdx.y * UT_BM;
    const int col0 = blockIdx.x * UT_BN;
    if (tid < UT_BM) {
        const int64_t r = row0 + tid;
        s_src[tid] = (r < n_rows) ? (int64_t)src_step[r] * data_ld : -1;
    }
    __syncthreads();
    double acc[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int fg = lane >> 2, ft = lane & 3;
    const int kk_l = tid & 31, rr_l = tid >> 5;          // element (rr_l + 8 e, kk_l) of a tile
    double ra[UT_EA], rb[UT_EB];
    auto fetch = [&](int k0) {
        const int c = k0 + kk_l;
#pragma unroll
        for (int e = 0; e < UT_EA; ++e) {
            const int rr = rr_l + 8 * e;
            const int64_t so = s_src[rr];
            double v = 0.0;
            if (so >= 0 && c < n_stn) {
                const double z = data[so + c];
                const bool fin = (z == z);
                v = (row0 + rr >= n_data) ? (fin ? 1.0 : 0.0) : (fin ? z : 0.0);
            }
            ra[e] = v;
        }
#pragma unroll
        for (int e = 0; e < UT_EB; ++e) {
            const int n = col0 + rr_l + 8 * e;
            rb[e] = (n < M && c < M) ? G[(int64_t)n * M + c] : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < M; k0 += UT_BK) {
#pragma unroll
        for (int e = 0; e < UT_EA; ++e) As[rr_l + 8 * e][kk_l] = ra[e];
#pragma unroll
        for (int e = 0; e < UT_EB; ++e) Bs[rr_l + 8 * e][kk_l] = rb[e];
        __syncthreads();
        if (k0 + UT_BK < M) fetch(k0 + UT_BK);
#pragma unroll
        for (int kk = 0; kk < UT_BK; kk += 4) {
            double af[2], bf[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = As[wr * 16 + 8 * i + fg][kk + ft];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[wc * 16 + 8 * j + fg][kk + ft];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_ut(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t r = row0 + wr * 16 + 8 * i + fg;
        if (r >= n_rows) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = col0 + wc * 16 + 8 * j + 2 * ft;
            if (c < M) ut[r * M + c] = acc[i][j][0];
            if (c + 1 < M) ut[r * M + c + 1] = acc[i][j][1];
        }
    }
}

// ------------------------------------------------------------------ sparse covariance solve
// Ordinary kriging with a compactly supported variogram (include/spx_b200.h, spx_sparse_cov):
// the station matrix is F 11' - C with C block diagonal over small connected components.
constexpr int SP_MAX = SPX_SPARSE_MAX_COMP;

__global__ void k_sparse_blocks(const double* __restrict__ stn_x, const double* __restrict__ stn_y,
                                spx_sparse_cov sp, spx_vg vgh, double min_vg_val, double base_f) {
    __shared__ VgDev vg;
    if (threadIdx.x == 0) {
        vg.n_terms = vgh.n_terms;
        for (int i = 0; i < SPX_VG_MAX_TERMS; ++i) {
            vg.types[i] = vgh.types[i];
            vg.sills[i] = vgh.sills[i];
            vg.ranges[i] = vgh.ranges[i];
        }
    }
    __syncthreads();
    for (int c = blockIdx.x; c < sp.n_comp; c += gridDim.x) {
        const int o = sp.comp_off[c], s = sp.comp_off[c + 1] - o;
        double* blk = sp.blk + sp.blk_off[c];
        for (int e = threadIdx.x; e < s * s; e += blockDim.x) {
            const int i = e / s, j = e - i * s;
            const int a = sp.comp_stn[o + i], b = sp.comp_stn[o + j];
            const double h = dist_rn(stn_x[a], stn_y[a], stn_x[b], stn_y[b]);
            blk[e] = base_f - vg_eval(vg, h, 0, min_vg_val);       // same entries as k_assemble
        }
    }
}

__global__ void __launch_bounds__(128) k_sparse_ok(
    const double* __restrict__ data, int n_stn, int64_t ld, const int32_t* __restrict__ row_step,
    int64_t n_rows, spx_sparse_cov sp, double base_f, int kpad, double* __restrict__ coef,
    double* __restrict__ coef_t, int64_t coef_t_ld, double* __restrict__ base,
    double* __restrict__ scr_a, double* __restrict__ scr_b, int32_t* __restrict__ info) {
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const double* __restrict__ z = data + (int64_t)row_step[row] * ld;
    double* __restrict__ A = scr_a + row * n_stn;      // C^-1 1   (0 at missing stations)
    double* __restrict__ B = scr_b + row * n_stn;      // C^-1 z
    double sa = 0.0, sb = 0.0;
    int fail = 0;
    // components are sorted by size: the first n_single are single stations (member list and
    // block offset == component index), then pairs, then the rest -- lanes of a warp run the
    // same code almost everywhere
#pragma unroll 4
    for (int c = lane; c < sp.n_single; c += 32) {
        const int k = sp.comp_stn[c];
        const double zz = z[k];
        double a = 0.0, b = 0.0;
        if (zz == zz) {
            const double d = sp.blk[c];
            if (!(d > 0.0)) fail = 1;
            a = 1.0 / d;
            b = zz * a;
        }
        A[k] = a;
        B[k] = b;
        sa += a;
        sb += b;
    }
    for (int c = sp.n_single + lane; c < sp.n_comp; c += 32) {
        const int o = sp.comp_off[c], s = sp.comp_off[c + 1] - o;
        const double* __restrict__ blk = sp.blk + sp.blk_off[c];
        if (s == 2) {
            const int k0 = sp.comp_stn[o], k1 = sp.comp_stn[o + 1];
            const double z0 = z[k0], z1 = z[k1];
            const bool h0 = (z0 == z0), h1 = (z1 == z1);
            const double c00 = blk[0], c01 = blk[1], c11 = blk[3];
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            if (h0 && h1) {
                const double det = fma(c00, c11, -c01 * c01);
                if (!(det > 0.0) || !(c00 > 0.0)) fail = 1;
                const double inv = 1.0 / det;
                a0 = (c11 - c01) * inv;
                a1 = (c00 - c01) * inv;
                b0 = fma(c11, z0, -c01 * z1) * inv;
                b1 = fma(c00, z1, -c01 * z0) * inv;
            } else if (h0) {
                if (!(c00 > 0.0)) fail = 1;
                a0 = 1.0 / c00;
                b0 = z0 * a0;
            } else if (h1) {
                if (!(c11 > 0.0)) fail = 1;
                a1 = 1.0 / c11;
                b1 = z1 * a1;
            }
            A[k0] = a0; B[k0] = b0;
            A[k1] = a1; B[k1] = b1;
            sa += a0 + a1;
            sb += b0 + b1;
            continue;
        }
        int idx[SP_MAX];
        double L[SP_MAX * SP_MAX], ra[SP_MAX], rb[SP_MAX];
        int m = 0;
        for (int i = 0; i < s; ++i) {
            const int k = sp.comp_stn[o + i];
            const double zz = z[k];
            if (zz == zz) {
                idx[m] = i;
                ra[m] = 1.0;
                rb[m] = zz;
                ++m;
            } else {
                A[k] = 0.0;
                B[k] = 0.0;
            }
        }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j <= i; ++j) L[i * SP_MAX + j] = blk[idx[i] * s + idx[j]];
        // Cholesky C = L L', then both right-hand sides
        for (int j = 0; j < m; ++j) {
            double d = L[j * SP_MAX + j];
            for (int k = 0; k < j; ++k) d = fma(-L[j * SP_MAX + k], L[j * SP_MAX + k], d);
            if (!(d > 0.0)) { fail = 1; d = 1.0; }
            const double dj = sqrt(d);
            L[j * SP_MAX + j] = dj;
            for (int i = j + 1; i < m; ++i) {
                double v = L[i * SP_MAX + j];
                for (int k = 0; k < j; ++k) v = fma(-L[i * SP_MAX + k], L[j * SP_MAX + k], v);
                L[i * SP_MAX + j] = v / dj;
            }
        }
        for (int i = 0; i < m; ++i) {
            double va = ra[i], vb = rb[i];
            for (int k = 0; k < i; ++k) {
                va = fma(-L[i * SP_MAX + k], ra[k], va);
                vb = fma(-L[i * SP_MAX + k], rb[k], vb);
            }
            ra[i] = va / L[i * SP_MAX + i];
            rb[i] = vb / L[i * SP_MAX + i];
        }
        for (int i = m - 1; i >= 0; --i) {
            double va = ra[i], vb = rb[i];
            for (int k = i + 1; k < m; ++k) {
                va = fma(-L[k * SP_MAX + i], ra[k], va);
                vb = fma(-L[k * SP_MAX + i], rb[k], vb);
            }
            ra[i] = va / L[i * SP_MAX + i];
            rb[i] = vb / L[i * SP_MAX + i];
        }
        for (int i = 0; i < m; ++i) {
            const int k = sp.comp_stn[o + idx[i]];
            A[k] = ra[i];
            B[k] = rb[i];
            sa += ra[i];
            sb += rb[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(FULL, sa, o);
        sb += __shfl_xor_sync(FULL, sb, o);
    }
    const double nu = sb / sa;
    if (!(sa > 0.0) || !(fabs(nu) <= 1.7976931348623157e308)) fail = 1;
    fail = __any_sync(FULL, fail);
    __syncwarp();
    double sx = 0.0;
    double* __restrict__ crow = coef + row * (int64_t)kpad;
#pragma unroll 4
    for (int k = lane; k < n_stn; k += 32) {
        const double x = fma(nu, A[k], -B[k]);
        crow[k] = x;
        if (coef_t) coef_t[(int64_t)k * coef_t_ld + row] = x;
        sx += x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(FULL, sx, o);
    if (lane == 0) {
        crow[n_stn] = nu;
        if (coef_t) coef_t[(int64_t)n_stn * coef_t_ld + row] = nu;
        base[row] = fma(base_f, sx, nu);
        if (fail) atomicAdd(info, 1);
    }
}

// ------------------------------------------------------------------ the job
static inline int64_t al256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct FastLayout {
    // device slot
    int64_t d_data, d_plan, d_ut, d_flags, d_coef, d_coef_t, d_base, d_bytes;
    // pinned host slot
    int64_t h_data, h_plan, h_flags, h_bytes;
    int64_t plan_dev_bytes, plan_host_bytes, coef_rows, ld_t, flags_bytes;
};

static FastLayout fast_layout(const spx_fast_cfg& c) {
    FastLayout L{};
    const int64_t T = c.max_steps, N = c.n_stn, M = c.n_stn + c.n_border;
    L.plan_dev_bytes = spx_downdate_plan_bytes(T, c.n_stn) + al256(4 * ((T + 255) / 256 * 256));
    L.plan_host_bytes = spx_downdate_plan_host_bytes(T) + al256(4 * ((T + 255) / 256 * 256));
    L.coef_rows = (c.estimator == 1) ? (T + SPX_BM - 1) / SPX_BM * SPX_BM : T;
    L.ld_t = (T + 3) / 4 * 4;
    L.flags_bytes = al256(8 * 2 * T + 4 * T);
    int64_t o = 0;
    L.d_data = o; o += al256(8 * T * N);
    L.d_plan = o; o += al256(L.plan_dev_bytes);
    L.d_ut = o; o += al256(8 * 2 * T * M);
    L.d_flags = o; o += L.flags_bytes;
    L.d_coef = o; o += al256(8 * L.coef_rows * c.kpad);
    L.d_coef_t = o; o += c.want_coef_t ? al256(8 * (int64_t)c.kpad * L.ld_t) : 0;
    L.d_base = o; o += al256(8 * T);
    L.d_bytes = o;
    o = 0;
    L.h_data = o; o += al256(8 * T * N);
    L.h_plan = o; o += al256(L.plan_host_bytes);
    L.h_flags = o; o += L.flags_bytes;
    L.h_bytes = o;
    return L;
}

struct FastSlot {
    uint8_t* dev = nullptr;
    uint8_t* host = nullptr;
    cudaEvent_t ev_solved = nullptr, ev_done = nullptr, ev_up = nullptr;
    cudaEvent_t ev_e0 = nullptr, ev_e1 = nullptr, ev_s0 = nullptr;
    bool used = false;
    spx_dd_plan plan{};
    int64_t n_krige = 0;
    bool sparse = false;
};

struct FastJob {
    spx_fast_cfg cfg;
    FastLayout L;
    FastSlot slots[8];
    cudaStream_t solve = nullptr;
    int next = 0;
    int device = 0;
};

}  // namespace spx

using namespace spx;

extern "C" {

int spx_ut_gemm_dev(const double* data, int32_t n_stn, int64_t data_ld, const int32_t* src_step,
                    int64_t n_rows, int64_t n_data, int32_t n_border, const double* ginv,
                    double* ut, void* stream) {
    if (n_rows == 0) return SPX_OK;
    if (!data || !src_step || !ginv || !ut || n_stn < 1 || n_border < 0 || n_data > n_rows) {
        set_error("ut_gemm: bad argument");
        return SPX_EINVAL;
    }
    const int M = n_stn + n_border;
    dim3 grid((unsigned)((M + UT_BN - 1) / UT_BN), (unsigned)((n_rows + UT_BM - 1) / UT_BM));
    if (grid.y > 65535u) {
        set_error("ut_gemm: too many rows in one launch");
        return SPX_EINVAL;
    }
    k_ut_gemm<<<grid, 256, 0, (cudaStream_t)stream>>>(data, n_stn, data_ld, src_step, n_rows,
                                                     n_data, M, ginv, ut);
    SPX_CHECK_LAUNCH("k_ut_gemm");
    return SPX_OK;
}

int spx_sparse_cov_blocks_dev(const double* stn_x, const double* stn_y, const spx_sparse_cov* sp,
                              const spx_vg* vg, double min_vg_val, double base_f, void* stream) {
    if (!stn_x || !stn_y || !sp || !vg || sp->n_comp < 1 || !sp->comp_off || !sp->comp_stn ||
        !sp->blk_off || !sp->blk || sp->max_size < 1 || sp->max_size > SPX_SPARSE_MAX_COMP) {
        set_error("sparse_cov_blocks: bad argument (components of at most %d stations)",
                  SPX_SPARSE_MAX_COMP);
        return SPX_EINVAL;
    }
    const int blocks = sp->n_comp < 1024 ? sp->n_comp : 1024;
    k_sparse_blocks<<<blocks, 64, 0, (cudaStream_t)stream>>>(stn_x, stn_y, *sp, *vg, min_vg_val,
                                                            base_f);
    SPX_CHECK_LAUNCH("k_sparse_blocks");
    return SPX_OK;
}

int spx_krige_sparse_ok_dev(const double* data, int32_t n_stn, int64_t ld, const int32_t* row_step,
                            int64_t n_rows, const spx_sparse_cov* sp, double base_f, int32_t kpad,
                            double* coef, double* coef_t, int64_t coef_t_ld, double* base,
                            double* scratch, int32_t* info, void* stream) {
    if (n_rows == 0) return SPX_OK;
    if (!data || !row_step || !sp || !coef || !base || !scratch || !info || n_stn < 1 ||
        ld < n_stn || kpad < n_stn + 1 || sp->n_comp < 1 || sp->max_size > SPX_SPARSE_MAX_COMP ||
        sp->n_single < 0 || sp->n_single > sp->n_comp ||
        (coef_t && coef_t_ld < n_rows)) {
        set_error("krige_sparse_ok: bad argument");
        return SPX_EINVAL;
    }
    const int64_t blocks = (n_rows + 3) / 4;           // 4 warps per block: spread over the SMs
    k_sparse_ok<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        data, n_stn, ld, row_step, n_rows, *sp, base_f, kpad, coef, coef_t, coef_t_ld, base,
        scratch, scratch + n_rows * n_stn, info);
    SPX_CHECK_LAUNCH("k_sparse_ok");
    return SPX_OK;
}

int64_t spx_fast_slot_bytes(const spx_fast_cfg* cfg, int32_t pinned_host) {
    if (!cfg || cfg->n_stn < 1 || cfg->max_steps < 1 || cfg->kpad < cfg->n_stn + cfg->n_border)
        return 0;
    const FastLayout L = fast_layout(*cfg);
    return pinned_host ? L.h_bytes : L.d_bytes;
}

int spx_fast_create(const spx_fast_cfg* cfg, void* dev_arena, void* host_arena, void** job_out) {
    if (!cfg || !dev_arena || !host_arena || !job_out || cfg->n_stn < 1 || cfg->n_border < 1 ||
        cfg->max_steps < 1 || cfg->n_slots < 2 || cfg->n_slots > 8 ||
        (!cfg->ginv && !(cfg->sparse.n_comp > 0 && cfg->estimator == 0 && cfg->n_border == 1)) ||
        cfg->kpad % 4 != 0 || cfg->kpad < cfg->n_stn + cfg->n_border ||
        (cfg->estimator != 0 && cfg->estimator != 1)) {
        set_error("fast_create: bad configuration");
        return SPX_EINVAL;
    }
    FastJob* j = new (std::nothrow) FastJob();
    if (!j) {
        set_error("fast_create: out of host memory");
        return SPX_ENOMEM;
    }
    j->cfg = *cfg;
    j->L = fast_layout(*cfg);
    SPX_CUDA(cudaGetDevice(&j->device));
    int lo = 0, hi = 0;
    SPX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SPX_CUDA(cudaStreamCreateWithPriority(&j->solve, cudaStreamNonBlocking, hi));
    for (int k = 0; k < cfg->n_slots; ++k) {
        FastSlot& s = j->slots[k];
        s.dev = static_cast<uint8_t*>(dev_arena) + (int64_t)k * j->L.d_bytes;
        s.host = static_cast<uint8_t*>(host_arena) + (int64_t)k * j->L.h_bytes;
        SPX_CUDA(cudaEventCreateWithFlags(&s.ev_solved,
                                          cfg->profile ? cudaEventDefault : cudaEventDisableTiming));
        SPX_CUDA(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        SPX_CUDA(cudaEventCreateWithFlags(&s.ev_up, cudaEventDisableTiming));
        if (cfg->profile) {
            SPX_CUDA(cudaEventCreate(&s.ev_e0));
            SPX_CUDA(cudaEventCreate(&s.ev_e1));
            SPX_CUDA(cudaEventCreate(&s.ev_s0));
        }
    }
    *job_out = j;
    return SPX_OK;
}

int spx_fast_destroy(void* job) {
    FastJob* j = static_cast<FastJob*>(job);
    if (!j) return SPX_OK;
    for (int k = 0; k < j->cfg.n_slots; ++k) {
        FastSlot& s = j->slots[k];
        if (s.used) cudaEventSynchronize(s.ev_done);
        if (s.ev_solved) cudaEventDestroy(s.ev_solved);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_up) cudaEventDestroy(s.ev_up);
        if (s.ev_e0) cudaEventDestroy(s.ev_e0);
        if (s.ev_e1) cudaEventDestroy(s.ev_e1);
        if (s.ev_s0) cudaEventDestroy(s.ev_s0);
    }
    if (j->solve) {
        cudaStreamSynchronize(j->solve);
        cudaStreamDestroy(j->solve);
    }
    cudaGetLastError();
    delete j;
    return SPX_OK;
}

int spx_fast_submit(void* job, const double* data, int64_t n_steps, int64_t ld, void* out,
                    void* main_stream, int32_t* grp_of_step, int32_t* grp_first,
                    int32_t* grp_n, int32_t* n_avail, uint8_t* step_flag, uint64_t* grp_bits,
                    spx_fast_result* res) {
    FastJob* j = static_cast<FastJob*>(job);
    if (!j || !data || !out || !grp_of_step || !grp_first || !grp_n || !n_avail || !step_flag ||
        !res || n_steps < 1 || n_steps > j->cfg.max_steps || ld < j->cfg.n_stn) {
        set_error("fast_submit: bad argument (n_steps %lld, max %d)", (long long)n_steps,
                  j ? j->cfg.max_steps : 0);
        return SPX_EINVAL;
    }
    const spx_fast_cfg& c = j->cfg;
    const FastLayout& L = j->L;
    const int N = c.n_stn, M = c.n_stn + c.n_border;
    cudaStream_t st_main = (cudaStream_t)main_stream;
    cudaStream_t st = c.solve_stream ? j->solve : st_main;
    std::memset(res, 0, sizeof(*res));
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](int i) {
        const auto t = std::chrono::steady_clock::now();
        res->host_ms[i] += std::chrono::duration<double, std::milli>(t - t_prev).count();
        t_prev = t;
    };

    const int k = j->next;
    FastSlot& s = j->slots[k];
    if (s.used) SPX_CUDA(cudaEventSynchronize(s.ev_done));   // ring: never blocks in steady state
    lap(0);

    // ---- host: availability groups (+ dense pinned copy of pageable input) --------------
    cudaPointerAttributes pa;
    bool pinned = false;
    if (cudaPointerGetAttributes(&pa, data) == cudaSuccess)
        pinned = (pa.type == cudaMemoryTypeHost);
    else
        cudaGetLastError();
    double* h_data = reinterpret_cast<double*>(s.host + L.h_data);
    int32_t n_grps = 0;
    int rc = spx_avail_groups_host(data, n_steps, N, ld, c.min_var_thr, grp_of_step, grp_first,
                                   grp_n, nullptr, n_avail, step_flag,
                                   pinned ? nullptr : h_data, grp_bits, &n_grps);
    if (rc != SPX_OK) return rc;
    res->n_grps = n_grps;
    lap(1);

    // ---- kriged steps (interp/steps.py:282-283, :325-331, :677-688) --------------------
    uint8_t* h_plan = s.host + L.h_plan;
    int32_t* h_rowdst = reinterpret_cast<int32_t*>(h_plan + spx_downdate_plan_host_bytes(c.max_steps));
    // the row list doubles as the plan's `steps` input; rows[i] = i
    int64_t* rows = reinterpret_cast<int64_t*>(s.host + L.h_flags);   // scratch until the check
    int64_t nk = 0;
    for (int64_t t = 0; t < n_steps; ++t) {
        const int na = n_avail[t];
        if (na == 0) ++res->n_none;
        else if (na == 1) ++res->n_single;
        else if (!step_flag[t]) ++res->n_mean;
        else {
            h_rowdst[nk] = (int32_t)t;
            rows[nk] = nk;
            ++nk;
        }
    }
    res->n_krige = (int32_t)nk;
    res->slot = k;
    if (nk == 0) {
        res->status = 1;
        return SPX_OK;
    }
    const int64_t coef_rows = (c.estimator == 1) ? (nk + SPX_BM - 1) / SPX_BM * SPX_BM : nk;
    for (int64_t i = nk; i < (c.estimator == 1 ? coef_rows : nk); ++i) h_rowdst[i] = -1;
    if (c.sparse.n_comp > 0 && c.estimator == 0 && c.n_border == 1) {
        // ---- sparse-covariance solve: no plan, no Ut, no factorisation of r x r blocks ----
        res->n_sys = (int32_t)nk;
        lap(2);
        j->next = (k + 1) % c.n_slots;
        s.used = true;
        s.n_krige = nk;
        s.sparse = true;
        double* d_data = reinterpret_cast<double*>(s.dev + L.d_data);
        uint8_t* d_plan = s.dev + L.d_plan;
        cudaStream_t st_up = j->solve;
        if (pinned && ld != N) {
            SPX_CUDA(cudaMemcpy2DAsync(d_data, 8 * (size_t)N, data, 8 * (size_t)ld, 8 * (size_t)N,
                                       (size_t)n_steps, cudaMemcpyHostToDevice, st_up));
        } else {
            SPX_CUDA(cudaMemcpyAsync(d_data, pinned ? data : h_data, 8 * (size_t)n_steps * N,
                                     cudaMemcpyHostToDevice, st_up));
        }
        int32_t* d_rowdst =
            reinterpret_cast<int32_t*>(d_plan + spx_downdate_plan_bytes(c.max_steps, N));
        SPX_CUDA(cudaMemcpyAsync(d_rowdst, h_rowdst, 4 * (size_t)nk, cudaMemcpyHostToDevice, st_up));
        if (st != st_up) {
            SPX_CUDA(cudaEventRecord(s.ev_up, st_up));
            SPX_CUDA(cudaStreamWaitEvent(st, s.ev_up, 0));
        }
        if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_s0, st));
        res->h2d_bytes = 8 * n_steps * N + 4 * nk;
        res->d_data = d_data;
        lap(3);
        uint8_t* d_flags = s.dev + L.d_flags;
        SPX_CUDA(cudaMemsetAsync(d_flags, 0, 8, st));
        double* d_coef = reinterpret_cast<double*>(s.dev + L.d_coef);
        SPX_CUDA(cudaMemsetAsync(d_coef, 0, 8 * (size_t)nk * c.kpad, st));
        double* d_base = reinterpret_cast<double*>(s.dev + L.d_base);
        double* d_coef_t = c.want_coef_t ? reinterpret_cast<double*>(s.dev + L.d_coef_t) : nullptr;
        const int64_t ld_t = (nk + 3) / 4 * 4;
        rc = spx_krige_sparse_ok_dev(d_data, N, N, d_rowdst, nk, &c.sparse, c.base_f, c.kpad,
                                     d_coef, d_coef_t, ld_t, d_base,
                                     reinterpret_cast<double*>(s.dev + L.d_ut),
                                     reinterpret_cast<int32_t*>(d_flags), st);
        if (rc != SPX_OK) return rc;
        rc = spx_copy_to_mapped_host_dev(s.host + L.h_flags, d_flags, 8, st);
        if (rc != SPX_OK) return rc;
        SPX_CUDA(cudaEventRecord(s.ev_solved, st));
        res->launches = 2;
        lap(4);
        if (c.solve_stream) SPX_CUDA(cudaStreamWaitEvent(st_main, s.ev_solved, 0));
        if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_e0, st_main));
        spx_local Lc = c.local;
        Lc.coef = d_coef;
        Lc.base = d_base;
        Lc.n_rows = nk;
        Lc.row_dst = d_rowdst;
        Lc.out = out;
        Lc.rows_all_valid = 1;
        Lc.coef_t = d_coef_t;
        Lc.coef_t_ld = d_coef_t ? ld_t : 0;
        rc = spx_estimate_local_dev(&Lc, st_main);
        if (rc != SPX_OK) return rc;
        if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_e1, st_main));
        SPX_CUDA(cudaEventRecord(s.ev_done, st_main));
        lap(5);
        res->launches += 1;
        res->status = 0;
        return SPX_OK;
    }
    s.sparse = false;
    spx_dd_plan& plan = s.plan;
    rc = spx_downdate_plan_host(grp_of_step, grp_n, n_grps, N, h_rowdst, rows, nk, h_plan,
                                spx_downdate_plan_host_bytes(c.max_steps), &plan);
    if (rc != SPX_OK) return rc;
    res->n_sys = plan.n_sys;
    res->max_r = plan.max_r;
    if (plan.n_sys < c.min_systems || plan.max_r > spx_krige_downdate_reg_max_r()) {
        res->status = 1;
        return SPX_OK;
    }
    lap(2);
    j->next = (k + 1) % c.n_slots;
    s.used = true;
    s.n_krige = nk;

    // ---- device: uploads + solve phase on the solve stream ------------------------------
    double* d_data = reinterpret_cast<double*>(s.dev + L.d_data);
    uint8_t* d_plan = s.dev + L.d_plan;
    // uploads on the job's own stream (they overlap the previous chunk's kernels whichever
    // stream those run on); the solve phase waits for them
    cudaStream_t st_up = j->solve;
    if (pinned && ld != N) {
        SPX_CUDA(cudaMemcpy2DAsync(d_data, 8 * (size_t)N, data, 8 * (size_t)ld, 8 * (size_t)N,
                                   (size_t)n_steps, cudaMemcpyHostToDevice, st_up));
    } else {
        SPX_CUDA(cudaMemcpyAsync(d_data, pinned ? data : h_data, 8 * (size_t)n_steps * N,
                                 cudaMemcpyHostToDevice, st_up));
    }
    SPX_CUDA(cudaMemcpyAsync(d_plan, h_plan, (size_t)plan.n_upload_bytes, cudaMemcpyHostToDevice,
                             st_up));
    // row_dst sits behind the plan's device buffer
    int32_t* d_rowdst = reinterpret_cast<int32_t*>(d_plan + spx_downdate_plan_bytes(c.max_steps, N));
    SPX_CUDA(cudaMemcpyAsync(d_rowdst, h_rowdst, 4 * (size_t)coef_rows, cudaMemcpyHostToDevice,
                             st_up));
    if (st != st_up) {
        SPX_CUDA(cudaEventRecord(s.ev_up, st_up));
        SPX_CUDA(cudaStreamWaitEvent(st, s.ev_up, 0));
    }
    if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_s0, st));
    res->h2d_bytes = 8 * n_steps * N + plan.n_upload_bytes + 4 * coef_rows;
    res->d_data = d_data;
    lap(3);

    const int n_sys = plan.n_sys, n_data = plan.n_data, n_rhs = plan.n_rhs;
    rc = spx_avail_lists_dev(d_data, N, N,
                             reinterpret_cast<const int32_t*>(d_plan + plan.off_bt_step) + n_data,
                             n_sys, reinterpret_cast<const int64_t*>(d_plan + plan.off_sys_stn_off),
                             reinterpret_cast<int32_t*>(d_plan + plan.off_stn_list),
                             reinterpret_cast<const int64_t*>(d_plan + plan.off_sys_miss_off),
                             reinterpret_cast<int32_t*>(d_plan + plan.off_miss_list), st);
    if (rc != SPX_OK) return rc;
    double* d_ut = reinterpret_cast<double*>(s.dev + L.d_ut);
    rc = spx_ut_gemm_dev(d_data, N, N, reinterpret_cast<const int32_t*>(d_plan + plan.off_bt_step),
                         n_rhs, n_data, c.n_border, c.ginv, d_ut, st);
    if (rc != SPX_OK) return rc;
    uint8_t* d_flags = s.dev + L.d_flags;
    const int64_t flags_used = 8 * (int64_t)n_rhs + 4 * (((int64_t)n_sys + 1) & ~(int64_t)1);
    SPX_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)flags_used, st));
    double* d_coef = reinterpret_cast<double*>(s.dev + L.d_coef);
    SPX_CUDA(cudaMemsetAsync(d_coef, 0, 8 * (size_t)coef_rows * c.kpad, st));
    double* d_base = reinterpret_cast<double*>(s.dev + L.d_base);
    double* d_coef_t = c.want_coef_t ? reinterpret_cast<double*>(s.dev + L.d_coef_t) : nullptr;
    const int64_t ld_t = (nk + 3) / 4 * 4;

    spx_downdate D;
    std::memset(&D, 0, sizeof(D));
    D.n_sys = n_sys;
    D.n_stn = N;
    D.n_border = c.n_border;
    D.max_r = plan.max_r;
    D.ginv = c.ginv;
    D.sys_r = reinterpret_cast<const int32_t*>(d_plan + plan.off_sys_r);
    D.sys_miss_off = reinterpret_cast<const int64_t*>(d_plan + plan.off_sys_miss_off);
    D.miss_list = reinterpret_cast<const int32_t*>(d_plan + plan.off_miss_list);
    D.sys_n = reinterpret_cast<const int32_t*>(d_plan + plan.off_sys_n);
    D.sys_stn_off = reinterpret_cast<const int64_t*>(d_plan + plan.off_sys_stn_off);
    D.stn_list = reinterpret_cast<const int32_t*>(d_plan + plan.off_stn_list);
    D.sys_rhs_off = reinterpret_cast<const int64_t*>(d_plan + plan.off_sys_rhs_off);
    D.sys_rhs_cnt = reinterpret_cast<const int32_t*>(d_plan + plan.off_sys_rhs_cnt);
    D.rhs_urow = reinterpret_cast<const int32_t*>(d_plan + plan.off_rhs_urow);
    D.rhs_row = reinterpret_cast<const int64_t*>(d_plan + plan.off_rhs_row);
    D.rhs_kind = reinterpret_cast<const int32_t*>(d_plan + plan.off_rhs_kind);
    D.sys_order = reinterpret_cast<const int32_t*>(d_plan + plan.off_sys_order);
    D.ut = d_ut;
    D.kpad = c.kpad;
    D.coef = d_coef;
    D.coef_row_major = (c.estimator == 0) ? 1 : 0;
    D.resid = reinterpret_cast<double*>(d_flags);
    D.info = reinterpret_cast<int32_t*>(d_flags + 8 * (int64_t)n_rhs);
    if (c.estimator == 0) {
        D.base = d_base;
        D.base_f = c.base_f;
        if (d_coef_t) {
            D.coef_t = d_coef_t;
            D.coef_t_ld = ld_t;
        }
    }
    rc = spx_krige_downdate_dev(&D, st);
    if (rc != SPX_OK) return rc;
    rc = spx_copy_to_mapped_host_dev(s.host + L.h_flags, d_flags, flags_used, st);
    if (rc != SPX_OK) return rc;
    SPX_CUDA(cudaEventRecord(s.ev_solved, st));
    // station lists, Ut, downdate (large systems first if any, bulk, repair pass), flags
    res->launches = 2 + (plan.max_r > 112 ? 3 : 2) + 1;

    lap(4);
    // ---- estimate on the caller's stream behind the solve ------------------------------
    if (c.solve_stream) SPX_CUDA(cudaStreamWaitEvent(st_main, s.ev_solved, 0));
    if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_e0, st_main));
    if (c.estimator == 0) {
        spx_local Lc = c.local;
        Lc.coef = d_coef;
        Lc.base = d_base;
        Lc.n_rows = nk;
        Lc.row_dst = d_rowdst;
        Lc.out = out;
        Lc.rows_all_valid = 1;
        Lc.coef_t = d_coef_t;
        Lc.coef_t_ld = d_coef_t ? ld_t : 0;
        rc = spx_estimate_local_dev(&Lc, st_main);
    } else {
        spx_gemm g = c.gemm;
        g.coef = d_coef;
        g.n_rows = nk;
        g.row_dst = d_rowdst;
        g.out = out;
        rc = spx_estimate_gemm_dev(&g, st_main);
    }
    if (rc != SPX_OK) return rc;
    if (c.profile) SPX_CUDA(cudaEventRecord(s.ev_e1, st_main));
    SPX_CUDA(cudaEventRecord(s.ev_done, st_main));
    lap(5);
    res->launches += 1;
    res->status = 0;
    return SPX_OK;
}

int spx_fast_check(void* job, int32_t slot, int32_t* verdict) {
    FastJob* j = static_cast<FastJob*>(job);
    if (!j || !verdict || slot < 0 || slot >= j->cfg.n_slots || !j->slots[slot].used) {
        set_error("fast_check: bad argument");
        return SPX_EINVAL;
    }
    FastSlot& s = j->slots[slot];
    SPX_CUDA(cudaEventSynchronize(s.ev_solved));
    if (s.sparse) {                        // number of rows whose Cholesky failed
        *verdict = (*reinterpret_cast<const volatile int32_t*>(s.host + j->L.h_flags) != 0) ? 1 : 0;
        return SPX_OK;
    }
    const spx_dd_plan& plan = s.plan;
    const uint8_t* hf = s.host + j->L.h_flags;
    const double* resid = reinterpret_cast<const double*>(hf);
    const int32_t* info = reinterpret_cast<const int32_t*>(hf + 8 * (int64_t)plan.n_rhs);
    const int64_t* pos_ones =
        reinterpret_cast<const int64_t*>(s.host + j->L.h_plan + plan.off_pos_ones);
    int v = 0;
    for (int i = 0; i < plan.n_sys; ++i)
        if (info[i] != 0) { v = 1; break; }
    if (v == 0) {
        for (int i = 0; i < plan.n_sys; ++i) {
            const double dev = resid[pos_ones[i]] * j->cfg.lambda_bound;
            if (!(dev <= j->cfg.lambda_tol)) { v = 2; break; }
        }
    }
    *verdict = v;
    return SPX_OK;
}

int spx_fast_times(void* job, int32_t slot, float* estimate_ms, float* solve_ms) {
    FastJob* j = static_cast<FastJob*>(job);
    if (!j || slot < 0 || slot >= j->cfg.n_slots || !j->slots[slot].used || !j->cfg.profile) {
        set_error("fast_times: bad argument or profiling off");
        return SPX_EINVAL;
    }
    FastSlot& s = j->slots[slot];
    SPX_CUDA(cudaEventSynchronize(s.ev_e1));
    if (estimate_ms) SPX_CUDA(cudaEventElapsedTime(estimate_ms, s.ev_e0, s.ev_e1));
    if (solve_ms) SPX_CUDA(cudaEventElapsedTime(solve_ms, s.ev_s0, s.ev_solved));
    return SPX_OK;
}

int spx_fast_timeline(void* job, int32_t ref_slot, int32_t slot, float* t_ms) {
    FastJob* j = static_cast<FastJob*>(job);
    if (!j || !t_ms || slot < 0 || slot >= j->cfg.n_slots || ref_slot < 0 ||
        ref_slot >= j->cfg.n_slots || !j->slots[slot].used || !j->slots[ref_slot].used ||
        !j->cfg.profile) {
        set_error("fast_timeline: bad argument or profiling off");
        return SPX_EINVAL;
    }
    FastSlot& s = j->slots[slot];
    cudaEvent_t base = j->slots[ref_slot].ev_s0;
    SPX_CUDA(cudaEventSynchronize(s.ev_e1));
    SPX_CUDA(cudaEventSynchronize(base));
    cudaEvent_t evs[4] = {s.ev_s0, s.ev_solved, s.ev_e0, s.ev_e1};
    for (int i = 0; i < 4; ++i) {
        if (cudaEventElapsedTime(&t_ms[i], base, evs[i]) != cudaSuccess) {
            cudaGetLastError();
            float neg = 0.f;                       // event earlier than the base
            SPX_CUDA(cudaEventElapsedTime(&neg, evs[i], base));
            t_ms[i] = -neg;
        }
    }
    return SPX_OK;
}

}  // extern "C"
