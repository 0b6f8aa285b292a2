// Fused "fill + contract" estimate kernel (the hot loop of the whole path).
//
//   Z[row, cell] = sum_k coef[row, k] * B[k, cell]
//
// * B (the dst<->station variogram / IDW-weight matrix with its border rows) is
//   NEVER written to HBM: each persistent thread block owns a tile of NT*8
//   cells, generates the full-K tile [kpad x NT*8] from coordinates once into
//   shared memory (already in DMMA B-fragment order) and keeps it resident.
// * The block then sweeps every coefficient row: 8 consumer warps x 32 rows per
//   M-tile, FP64 tensor-core MMA (mma.sync.aligned.m8n8k4.f64 -> SASS
//   DMMA.8x8x4), accumulators in registers.
// * Coefficients arrive pre-packed in A-fragment order; a producer warp streams
//   16 KB stages (256 rows x 8 k) with cp.async.bulk (TMA bulk copy, SASS
//   UBLKCP) into an mbarrier full/empty ring.
// * Epilogue: clamp (steps.py:466-476), cast, scatter to the masked field; or
//   raw FP64 to an auxiliary buffer (sum-of-weights rows); or divide by such a
//   row (IDW normalisation).
//
// Replaces interp/steps.py:403-435 (kriging estimate) and :293-313 (IDW)
// together with the dst<->station fills of :639-650 (pyx:123-226).
#include <cstdlib>

#include "spx_common.cuh"

namespace spx {

// consumer warps per block: 8 (32 rows each) or 16 (16 rows each)
constexpr int BK = 8;                          // k per pipeline stage
constexpr int STAGE_DOUBLES = SPX_BM * BK;     // 2048 doubles = 16 KB
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;

struct GemmArgs {
    const double* coef;
    int64_t n_rows;
    int kpad, n_stn, n_border;
    const double* stn_x;
    const double* stn_y;
    const double* cell_x;
    const double* cell_y;
    int64_t n_cells;
    const double* cell_drift;
    int gen, covar_flag;
    VgDev vg;
    VgFast vgf;
    double min_vg_val, idw_exp, inv_scale;
    int epi;
    const int32_t* row_dst;
    const int32_t* row_aux;
    void* out;
    int64_t out_ld;
    int out_f64;
    const int32_t* cell_pos;
    double* aux;
    int has_lo, has_hi;
    double lo, hi;
    int n_stages;
    int quad_slot;
    // K split over several passes (large K: few cells fit beside a full-K B tile): this
    // launch covers the k-chunks [kc_lo, kc_lo + kc_n) of 4 columns each; the accumulators
    // start from part_in (f64 [n_rows, n_cells], NULL = zero) and, when part_out is set, are
    // stored there raw instead of going through the epilogue (which only the last pass runs)
    int kc_lo, kc_n;
    const double* part_in;
    double* part_out;
    int* zk_glob;       // IDW: station on the cell centre, found by whichever pass holds it
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// (dist / scale) ** -p from the squared distance; small integer exponents avoid
// pow().  Reference: w = 1 / d**p on max-normalised distances
// (interp/steps.py:297-303, pyx:792); the common scale cancels in the ratio.
// element (row, col) of the packed coefficient matrix (spx_coef_offset of the header)
__device__ __forceinline__ int64_t gemm_coef_offset(int64_t row, int64_t col, int64_t kpad) {
    const int64_t mt = row / SPX_BM, r = row % SPX_BM;
    return (mt * (kpad / 4) + col / 4) * (SPX_BM * 4) + (r / 8) * 32 + (r % 8) * 4 + (col % 4);
}

__device__ __forceinline__ double idw_weight(double d2, double inv_scale, double p) {
    const double q2 = d2 * inv_scale * inv_scale;  // (d/scale)^2
    if (p == 2.0) return 1.0 / q2;
    const double q = sqrt(q2);
    if (p == 1.0) return 1.0 / q;
    if (p == 3.0) return 1.0 / (q2 * q);
    if (p == 4.0) return 1.0 / (q2 * q2);
    if (p == 5.0) return 1.0 / (q2 * q2 * q);
    return 1.0 / pow(q, p);
}

template <int NT, int CW, bool QUAD>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_estimate_gemm(const GemmArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int BN = NT * 8;
    constexpr int CONSUMER_WARPS = CW;
    constexpr int GEMM_THREADS = (CW + 1) * 32;
    constexpr int RW = SPX_BM / CW;   // rows per warp
    constexpr int MI = RW / 8;        // 8-row MMA tiles per warp
    const int KC = a.kc_n;                  // k-chunks of this pass (all of them: kpad / 4)
    const int KC_ALL = a.kpad >> 2;
    const int k_lo = a.kc_lo * 4;           // first coefficient column of this pass
    const int k_n = a.kc_n * 4;
    const int n_ksteps = k_n / BK;
    const int n_stages = a.n_stages;

    double* Bs = reinterpret_cast<double*>(smem_raw);                  // [KC][NT][32]
    double* As = Bs + (size_t)KC * NT * 32;                            // [stage][2][32][32]
    double* sx = As + (size_t)n_stages * STAGE_DOUBLES;                // [k_n] station x
    double* sy = sx + k_n;                                             // [k_n] station y
    double* cx = sy + k_n;                                             // [BN]
    double* cy = cx + BN;                                              // [BN]
    double* qs = cy + BN;                                              // [BN] quadratic forms
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(qs + BN);         // [n_stages]
    uint64_t* empty_bar = full_bar + n_stages;                         // [n_stages]
    // IDW: station that coincides with the cell centre (distance 0), -1 = none
    int* zk = reinterpret_cast<int*>(empty_bar + n_stages);            // [BN]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int k = tid; k < k_n; k += GEMM_THREADS) {
        sx[k] = (k_lo + k < a.n_stn) ? a.stn_x[k_lo + k] : 0.0;
        sy[k] = (k_lo + k < a.n_stn) ? a.stn_y[k_lo + k] : 0.0;
    }
    __syncthreads();

    const int64_t n_ctiles = (a.n_cells + BN - 1) / BN;
    const int64_t n_mtiles = (a.n_rows + SPX_BM - 1) / SPX_BM;
    int stage = 0;        // ring position at the start of the current tile (identical in
    uint32_t phase = 0;   // the producer and in every consumer)

    for (int64_t ct = blockIdx.x; ct < n_ctiles; ct += gridDim.x) {
        const int64_t cell0 = ct * BN;
        // ---- generate the resident B tile --------------------------------
        __syncthreads();  // previous tile fully consumed
        if (tid < BN) {
            if (QUAD) {
                if (ct != (int64_t)blockIdx.x) {   // result of the previous tile
                    const int64_t pc = cell0 - (int64_t)gridDim.x * BN + tid;
                    if (pc < a.n_cells) a.aux[(int64_t)a.quad_slot * a.n_cells + pc] = qs[tid];
                }
                qs[tid] = 0.0;
            }
            const int64_t c = cell0 + tid;
            cx[tid] = (c < a.n_cells) ? a.cell_x[c] : 0.0;
            cy[tid] = (c < a.n_cells) ? a.cell_y[c] : 0.0;
            zk[tid] = -1;
        }
        __syncthreads();
        const int n_ent = KC * NT * 32;
        for (int idx = tid; idx < n_ent; idx += GEMM_THREADS) {
            const int ln = idx & 31;
            const int q = idx >> 5;
            const int j = q % NT;
            const int kc = q / NT;
            const int k = kc * 4 + (ln & 3);          // column inside this pass
            const int kg = k_lo + k;                  // coefficient column
            const int n = j * 8 + (ln >> 2);
            const int64_t c = cell0 + n;
            double v = 0.0;
            if (c < a.n_cells) {
                if (kg < a.n_stn) {
                    if (a.gen == SPX_GEN_VG) {
                        const double dx = cx[n] - sx[k], dy = cy[n] - sy[k];
                        const double h = sqrt(dx * dx + dy * dy);
                        v = a.vgf.all_fast ? vg_eval_fast(a.vgf, h, a.covar_flag, a.min_vg_val)
                                           : vg_eval(a.vg, h, a.covar_flag, a.min_vg_val);
                    } else {
                        const double dx = cx[n] - sx[k], dy = cy[n] - sy[k];
                        const double d2 = dx * dx + dy * dy;
                        if (d2 == 0.0) {
                            // a station exactly on the cell centre: weight inf.  The reference
                            // (steps.py:293-313) gives NaN where that station is available and
                            // simply leaves it out where it is missing; 0 * inf would turn the
                            // latter into NaN too.  Its weight is dropped here, the
                            // sum-of-weights epilogue (SPX_EPI_AUX) restores the NaN for the
                            // groups that have the station.
                            v = 0.0;
                            zk[n] = kg;
                            if (a.zk_glob) a.zk_glob[c] = kg;
                        } else {
                            v = idw_weight(d2, a.inv_scale, a.idw_exp);
                        }
                    }
                } else if (kg < a.n_stn + a.n_border) {
                    const int b = kg - a.n_stn;
                    v = (b == 0) ? 1.0 : a.cell_drift[(int64_t)(b - 1) * a.n_cells + c];
                }
            }
            Bs[idx] = v;
        }
        __syncthreads();
        if (a.zk_glob && !a.part_out) {     // last pass: stations found by the earlier ones
            if (tid < BN && cell0 + tid < a.n_cells && zk[tid] < 0) zk[tid] = a.zk_glob[cell0 + tid];
            __syncthreads();
        }

        if (warp == CONSUMER_WARPS) {
            // ---- producer: stream coefficient stages ----------------------
            if (lane == 0) {
                int s = stage;
                uint32_t ph = phase;
                for (int64_t mt = 0; mt < n_mtiles; ++mt) {
                    const double* src = a.coef + ((size_t)mt * KC_ALL + a.kc_lo) * (SPX_BM * 4);
                    for (int ks = 0; ks < n_ksteps; ++ks) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                        bulk_g2s(As + (size_t)s * STAGE_DOUBLES, src + (size_t)ks * STAGE_DOUBLES,
                                 STAGE_BYTES, &full_bar[s]);
                        if (++s == n_stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else {
            // ---- consumers: DMMA sweep over all rows ---------------------
            const int g = lane >> 2, t4 = lane & 3;
            int s = stage;
            uint32_t ph = phase;
            double quad[QUAD ? NT : 1][2];
#pragma unroll
            for (int j = 0; j < (QUAD ? NT : 1); ++j) quad[j][0] = quad[j][1] = 0.0;
            for (int64_t mt = 0; mt < n_mtiles; ++mt) {
                const int64_t row_base = mt * SPX_BM + warp * RW;
                const bool active = row_base < a.n_rows;
                // destination rows of this lane's 4 accumulator rows, fetched before the
                // k loop so that the latency is hidden
                int dst_r[MI], aux_r[MI];
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int64_t R = row_base + i * 8 + g;
                    dst_r[i] = (R < a.n_rows) ? a.row_dst[R] : -1;
                    aux_r[i] = (a.epi == SPX_EPI_FIELD_DIV && R < a.n_rows) ? a.row_aux[R] : 0;
                }
                double acc[MI][NT][2];
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                if (a.part_in && active) {             // sums of the earlier K passes
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        const int64_t R = row_base + i * 8 + g;
                        if (R >= a.n_rows) continue;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const int64_t c = cell0 + j * 8 + t4 * 2;       // even
                            const double* src = a.part_in + R * a.n_cells + c;
                            if (c + 1 < a.n_cells && (a.n_cells & 1) == 0) {
                                const double2 v = __ldcs(reinterpret_cast<const double2*>(src));
                                acc[i][j][0] = v.x;
                                acc[i][j][1] = v.y;
                            } else {
                                if (c < a.n_cells) acc[i][j][0] = __ldcs(src);
                                if (c + 1 < a.n_cells) acc[i][j][1] = __ldcs(src + 1);
                            }
                        }
                    }
                }

                for (int ks = 0; ks < n_ksteps; ++ks) {
                    mbar_wait(&full_bar[s], ph);
                    if (active) {
                        const double* Ast = As + (size_t)s * STAGE_DOUBLES;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            double af[MI], bf[NT];
#pragma unroll
                            for (int i = 0; i < MI; ++i)
                                af[i] = Ast[(kk * 32 + warp * MI + i) * 32 + lane];
                            const double* Bk = Bs + (size_t)(ks * 2 + kk) * NT * 32;
#pragma unroll
                            for (int j = 0; j < NT; ++j) bf[j] = Bk[j * 32 + lane];
#pragma unroll
                            for (int i = 0; i < MI; ++i)
#pragma unroll
                                for (int j = 0; j < NT; ++j)
                                    dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    if (++s == n_stages) { s = 0; ph ^= 1; }
                }
                if (!active) continue;
                if (a.part_out) {                      // not the last K pass: raw sums
#pragma unroll
                    for (int i = 0; i < MI; ++i) {
                        const int64_t R = row_base + i * 8 + g;
                        if (R >= a.n_rows) continue;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const int64_t c = cell0 + j * 8 + t4 * 2;       // even
                            double* dstp = a.part_out + R * a.n_cells + c;
                            if (c + 1 < a.n_cells && (a.n_cells & 1) == 0) {
                                __stcs(reinterpret_cast<double2*>(dstp),
                                       make_double2(acc[i][j][0], acc[i][j][1]));
                            } else {
                                if (c < a.n_cells) __stcs(dstp, acc[i][j][0]);
                                if (c + 1 < a.n_cells) __stcs(dstp + 1, acc[i][j][1]);
                            }
                        }
                    }
                    continue;
                }
                // ---- epilogue -------------------------------------------
                const bool pair_ok = (a.cell_pos == nullptr) && ((a.out_ld & 1) == 0);
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const int dst = dst_r[i];
                    if (dst < 0) continue;
                    if (QUAD) {
                        // row = one row of A^-1 (K index dst): accumulate
                        // lambda[dst] * rhs[dst] (+ lambda[n], steps.py:431-434)
                        const int kc = dst >> 2, kt = dst & 3;
                        const double extra = (dst == a.n_stn) ? 1.0 : 0.0;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const double* Bk = Bs + ((size_t)kc * NT + j) * 32;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int n8 = t4 * 2 + e;
                                quad[QUAD ? j : 0][e] =
                                    fma(acc[i][j][e], Bk[n8 * 4 + kt] + extra, quad[QUAD ? j : 0][e]);
                            }
                        }
                        continue;
                    }
                    const int64_t aux_row = (int64_t)aux_r[i] * a.n_cells;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int64_t c = cell0 + j * 8 + t4 * 2;   // even
                        if (c >= a.n_cells) continue;
                        const bool has2 = (c + 1 < a.n_cells);
                        double v0 = acc[i][j][0], v1 = acc[i][j][1];
                        if (a.epi == SPX_EPI_AUX && a.gen == SPX_GEN_IDW) {
                            // rows are availability masks (1 / 0): the coincident station is
                            // part of this group -> sum of weights inf -> estimate NaN
                            const int64_t R = row_base + i * 8 + g;
                            const int n0 = j * 8 + t4 * 2;
                            if (zk[n0] >= 0 && a.coef[gemm_coef_offset(R, zk[n0], a.kpad)] != 0.0)
                                v0 = CUDART_NAN;
                            if (zk[n0 + 1] >= 0 && a.coef[gemm_coef_offset(R, zk[n0 + 1], a.kpad)] != 0.0)
                                v1 = CUDART_NAN;
                        }
                        if (a.epi == SPX_EPI_AUX) {
                            double* dstp = a.aux + (int64_t)dst * a.n_cells + c;
                            if (has2 && ((a.n_cells & 1) == 0)) {
                                *reinterpret_cast<double2*>(dstp) = make_double2(v0, v1);
                            } else {
                                dstp[0] = v0;
                                if (has2) dstp[1] = v1;
                            }
                            continue;
                        }
                        if (a.epi == SPX_EPI_FIELD_DIV) {
                            v0 = v0 / a.aux[aux_row + c];
                            if (has2) v1 = v1 / a.aux[aux_row + c + 1];
                        }
                        v0 = clampd(v0, a.has_lo, a.has_hi, a.lo, a.hi);
                        v1 = clampd(v1, a.has_lo, a.has_hi, a.lo, a.hi);
                        if (pair_ok && has2) {
                            const int64_t o = (int64_t)dst * a.out_ld + c;
                            if (a.out_f64)
                                *reinterpret_cast<double2*>(reinterpret_cast<double*>(a.out) + o) =
                                    make_double2(v0, v1);
                            else
                                *reinterpret_cast<float2*>(reinterpret_cast<float*>(a.out) + o) =
                                    make_float2((float)v0, (float)v1);
                        } else {
                            const int64_t col0 = a.cell_pos ? (int64_t)a.cell_pos[c] : c;
                            store_out(a.out, (int64_t)dst * a.out_ld + col0, v0, a.out_f64);
                            if (has2) {
                                const int64_t col1 = a.cell_pos ? (int64_t)a.cell_pos[c + 1] : c + 1;
                                store_out(a.out, (int64_t)dst * a.out_ld + col1, v1, a.out_f64);
                            }
                        }
                    }
                }
            }
            if (QUAD) {
                // sum over this warp's rows (lanes with equal t4), then over warps
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double v = quad[QUAD ? j : 0][e];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (g == 0) atomicAdd(&qs[j * 8 + t4 * 2 + e], v);
                    }
            }
        }
        // advance the ring position by the stages this tile consumed
        {
            const uint32_t adv = (uint32_t)((n_mtiles * n_ksteps) % (2 * n_stages));
            uint32_t pos = (uint32_t)stage + (phase ? n_stages : 0) + adv;
            pos %= (uint32_t)(2 * n_stages);
            phase = pos >= (uint32_t)n_stages;
            stage = (int)(pos - (phase ? n_stages : 0));
        }
    }
    if (QUAD) {
        // result of the last tile this block processed
        __syncthreads();
        const int64_t n_mine = (n_ctiles > (int64_t)blockIdx.x)
                                   ? (n_ctiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        if (n_mine > 0 && tid < BN) {
            const int64_t last_ct = blockIdx.x + (n_mine - 1) * (int64_t)gridDim.x;
            const int64_t pc = last_ct * BN + tid;
            if (pc < a.n_cells) a.aux[(int64_t)a.quad_slot * a.n_cells + pc] = qs[tid];
        }
    }
}

// kpad: the K extent held by one launch (all of it, or one pass of a K split)
static size_t gemm_smem_bytes(int kpad, int nt, int n_stages) {
    const size_t dbl = (size_t)(kpad / 4) * nt * 32 + (size_t)n_stages * STAGE_DOUBLES +
                       2 * (size_t)kpad + 3 * (size_t)nt * 8;
    return dbl * 8 + 2 * (size_t)n_stages * 8 + (size_t)nt * 8 * sizeof(int);
}

struct GemmCfg {
    int nt, n_stages, grid;
    size_t smem;
};

static int pick_config(const spx_gemm* g, GemmCfg* cfg, int k_extent = 0) {
    const int kpad = k_extent > 0 ? k_extent : g->kpad;
    int dev = 0, max_smem = 0, n_sm = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    SPX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SPX_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    static const int nts[] = {8, 6, 5, 4, 3, 2, 1};
    // tuning knobs (benchmark experiments only)
    const int force_nt = getenv("SPX_GEMM_NT") ? atoi(getenv("SPX_GEMM_NT")) : 0;
    const int max_st = getenv("SPX_GEMM_STAGES") ? atoi(getenv("SPX_GEMM_STAGES")) : 4;
    for (int nt : nts) {
        if (force_nt > 0 && nt > force_nt) continue;
        for (int st = max_st; st >= 2; --st) {
            if (st == 2 && nt > 1) continue;  // prefer fewer cells over a 2-deep ring
            const size_t sm = gemm_smem_bytes(kpad, nt, st);
            if (sm <= (size_t)max_smem) {
                // do not use tiles wider than the problem
                int use_nt = nt;
                while (use_nt > 1 && (int64_t)(use_nt - 1) * 8 >= g->n_cells) --use_nt;
                if (use_nt != nt) continue;
                cfg->nt = nt;
                cfg->n_stages = st;
                cfg->smem = sm;
                const int64_t tiles = (g->n_cells + nt * 8 - 1) / (nt * 8);
                cfg->grid = (int)(tiles < n_sm ? tiles : n_sm);
                return SPX_OK;
            }
        }
    }
    set_error("estimate_gemm: kpad=%d does not fit in %d bytes of shared memory", g->kpad,
              max_smem);
    return SPX_ENOMEM;
}

static int g_consumer_warps = 16;  // SPX_GEMM_WARPS=8|16 overrides (tuning knob)
static int g_ksplit = -1;          // largest number of K passes (-1: SPX_GEMM_KSPLIT, default 4; 0 / 1: off)

template <int NT>
static int launch(const GemmArgs& a, const GemmCfg& cfg, cudaStream_t st) {
    if (a.epi == SPX_EPI_QUADFORM) {
        SPX_CUDA(cudaFuncSetAttribute(k_estimate_gemm<NT, 8, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        k_estimate_gemm<NT, 8, true><<<cfg.grid, 9 * 32, cfg.smem, st>>>(a);
    } else if (g_consumer_warps == 8) {
        SPX_CUDA(cudaFuncSetAttribute(k_estimate_gemm<NT, 8, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        k_estimate_gemm<NT, 8, false><<<cfg.grid, 9 * 32, cfg.smem, st>>>(a);
    } else {
        SPX_CUDA(cudaFuncSetAttribute(k_estimate_gemm<NT, 16, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        k_estimate_gemm<NT, 16, false><<<cfg.grid, 17 * 32, cfg.smem, st>>>(a);
    }
    SPX_CHECK_LAUNCH("k_estimate_gemm");
    return SPX_OK;
}

static int validate(const spx_gemm* g) {
    if (!g) {
        set_error("estimate_gemm: null argument");
        return SPX_EINVAL;
    }
    if (g->kpad <= 0 || g->kpad % BK != 0 || g->kpad < g->n_stn + g->n_border) {
        set_error("estimate_gemm: kpad=%d must be a multiple of %d and >= n_stn + n_border = %d",
                  g->kpad, BK, g->n_stn + g->n_border);
        return SPX_EINVAL;
    }
    if (g->n_border > 1 && !g->cell_drift) {
        set_error("estimate_gemm: drift rows requested without cell_drift");
        return SPX_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(g->coef) & 15) != 0) {
        set_error("estimate_gemm: coef must be 16-byte aligned");
        return SPX_EINVAL;
    }
    if (g->gen == SPX_GEN_VG && (g->vg.n_terms < 0 || g->vg.n_terms > SPX_VG_MAX_TERMS)) {
        set_error("estimate_gemm: bad variogram");
        return SPX_EINVAL;
    }
    if (g->gen == SPX_GEN_IDW && !(g->dist_scale > 0)) {
        set_error("estimate_gemm: dist_scale must be > 0");
        return SPX_EINVAL;
    }
    return SPX_OK;
}

}  // namespace spx

using namespace spx;

extern "C" {

int spx_gemm_set_ksplit(int max_passes) {
    const int prev = g_ksplit;
    g_ksplit = max_passes;
    return prev;
}

int spx_estimate_gemm_config(const spx_gemm* g, int* cells_per_block, int* n_stages,
                             int* smem_bytes, int* grid) {
    int rc = validate(g);
    if (rc) return rc;
    GemmCfg cfg;
    if ((rc = pick_config(g, &cfg))) return rc;
    if (cells_per_block) *cells_per_block = cfg.nt * 8;
    if (n_stages) *n_stages = cfg.n_stages;
    if (smem_bytes) *smem_bytes = (int)cfg.smem;
    if (grid) *grid = cfg.grid;
    return SPX_OK;
}

int spx_estimate_gemm_dev(const spx_gemm* g, void* stream) {
    int rc = validate(g);
    if (rc) return rc;
    if (g->n_rows == 0 || g->n_cells == 0) return SPX_OK;
    GemmCfg cfg;
    if ((rc = pick_config(g, &cfg))) return rc;
    if (const char* e = getenv("SPX_GEMM_WARPS")) g_consumer_warps = (atoi(e) == 8) ? 8 : 16;

    GemmArgs a;
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
    a.gen = g->gen;
    a.covar_flag = g->covar_flag;
    a.vg = to_dev(g->vg);
    a.vgf = make_vg_fast(g->vg.n_terms, g->vg.types, g->vg.sills, g->vg.ranges);
    a.min_vg_val = g->min_vg_val;
    a.idw_exp = g->idw_exp;
    a.inv_scale = (g->gen == SPX_GEN_IDW) ? 1.0 / g->dist_scale : 1.0;
    a.epi = g->epi;
    a.row_dst = g->row_dst;
    a.row_aux = g->row_aux;
    a.out = g->out;
    a.out_ld = g->out_ld;
    a.out_f64 = g->out_f64;
    a.cell_pos = g->cell_pos;
    a.aux = g->aux;
    a.has_lo = g->has_lo;
    a.has_hi = g->has_hi;
    a.lo = g->lo;
    a.hi = g->hi;
    a.n_stages = cfg.n_stages;
    a.quad_slot = g->quad_slot;
    a.kc_lo = 0;
    a.kc_n = g->kpad / 4;
    a.part_in = nullptr;
    a.part_out = nullptr;
    a.zk_glob = nullptr;

    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](const GemmCfg& c) -> int {
        a.n_stages = c.n_stages;
        switch (c.nt) {
            case 8: return launch<8>(a, c, st);
            case 6: return launch<6>(a, c, st);
            case 5: return launch<5>(a, c, st);
            case 4: return launch<4>(a, c, st);
            case 3: return launch<3>(a, c, st);
            case 2: return launch<2>(a, c, st);
            default: return launch<1>(a, c, st);
        }
    };
    // ---- large K: split it over passes --------------------------------------------------
    // With a full-K B tile only 8-16 cells fit next to it at kpad ~ 2000 and every coefficient
    // fragment feeds one or two DMMAs: the kernel is then bound by shared-memory traffic
    // (0.53 of the FP64 tensor peak).  P passes over K / P columns each hold 4-5 cell groups
    // like the kpad ~ 500 case (0.84); the partial sums travel through HBM as f64
    // [n_rows, n_cells] (16 bytes per row and cell and extra pass: ~3 ms per 10 GB against
    // hundreds of ms of contraction), the epilogue runs in the last pass only.
    if (g_ksplit < 0) g_ksplit = getenv("SPX_GEMM_KSPLIT") ? atoi(getenv("SPX_GEMM_KSPLIT")) : 4;
    const int split_knob = g_ksplit;
    const int64_t part_bytes = g->n_rows * g->n_cells * (int64_t)sizeof(double);
    if (cfg.nt <= 2 && split_knob >= 2 && g->epi != SPX_EPI_QUADFORM &&
        part_bytes <= (int64_t)24 << 30 && g->n_cells >= 64) {
        const int kc_all = g->kpad / 4;
        int best_p = 1;
        GemmCfg best = cfg;
        for (int P = 2; P <= split_knob && P <= 8; ++P) {
            int kc_pass = (kc_all + P - 1) / P;
            kc_pass += kc_pass & 1;                        // stages hold two k-chunks
            GemmCfg c;
            if (pick_config(g, &c, kc_pass * 4) != SPX_OK) continue;
            if (c.nt > best.nt) { best = c; best_p = P; }
            if (c.nt >= 5) break;
        }
        if (best_p > 1) {
            static bool pool_set = false;
            if (!pool_set) {                               // keep freed blocks for the next call
                int dev = 0;
                cudaMemPool_t pool;
                SPX_CUDA(cudaGetDevice(&dev));
                SPX_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
                uint64_t keep = ~0ull;
                SPX_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
                pool_set = true;
            }
            double* part = nullptr;
            int* zkg = nullptr;
            if (cudaMallocAsync(reinterpret_cast<void**>(&part), (size_t)part_bytes, st) !=
                cudaSuccess) {
                cudaGetLastError();                        // no room for the partial sums:
                return run(cfg);                           // one pass over the full K
            }
            if (g->gen == SPX_GEN_IDW) {
                if (cudaMallocAsync(reinterpret_cast<void**>(&zkg),
                                    sizeof(int) * (size_t)g->n_cells, st) != cudaSuccess) {
                    cudaGetLastError();
                    cudaFreeAsync(part, st);
                    return run(cfg);
                }
                SPX_CUDA(cudaMemsetAsync(zkg, 0xFF, sizeof(int) * (size_t)g->n_cells, st));
            }
            int kc_pass = (kc_all + best_p - 1) / best_p;
            kc_pass += kc_pass & 1;
            int rc2 = SPX_OK;
            for (int pss = 0, lo = 0; lo < kc_all && rc2 == SPX_OK; ++pss, lo += kc_pass) {
                const int n = (kc_all - lo < kc_pass) ? kc_all - lo : kc_pass;   // even: kpad % 8 == 0
                const bool last = lo + n >= kc_all;
                a.kc_lo = lo;
                a.kc_n = n;
                a.part_in = pss ? part : nullptr;
                a.part_out = last ? nullptr : part;
                a.zk_glob = zkg;
                GemmCfg c = best;
                if (n != kc_pass && pick_config(g, &c, n * 4) != SPX_OK) c = best;
                c.smem = gemm_smem_bytes(n * 4, c.nt, c.n_stages);
                rc2 = run(c);
            }
            cudaFreeAsync(part, st);
            if (zkg) cudaFreeAsync(zkg, st);
            return rc2;
        }
    }
    return run(cfg);
}

}  // extern "C"
