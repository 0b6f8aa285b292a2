// Lossless 2-byte transport of rounded f32 fields (interp/steps.py:907-945 writes
// np.round(fld, nmrl_prcn) in float32 to the netCDF file).
//
// A field rounded to d decimals holds v = fdiv(q, p) with p = float(10^d) and q an
// integer-valued float.  Per row (time step) the kernel recovers q' = rint(v * p) as
// int32, VERIFIES fdiv(float(q'), p) == v bit for bit on every element, and -- if the
// whole row round-trips and max(q') - min(q') <= 65533 -- emits 16-bit codes q' - qmin
// (0xFFFF = NaN, 0xFFFE = -0.0) instead of 32-bit floats.  The host rebuilds exactly the
// same floats (spx_unpack_field_host: cvt + IEEE division, AVX2, threaded).  Rows that do not qualify
// (not rounded, huge values, infinities, too wide a range) are flagged "raw" and keep
// their float representation: nothing is ever approximated.
// The device -> host copy of a chunk shrinks from 4 to 2 bytes per cell-step.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <immintrin.h>

#include "spx_b200.h"
#include "spx_common.cuh"
#include "spx_host_pool.h"
#include "spx_stats.cuh"

namespace spx {

constexpr int PK_THREADS = 256;

constexpr uint32_t PK_NAN = 0xFFFFu;       // code of NaN
constexpr uint32_t PK_NEGZERO = 0xFFFEu;   // code of -0.0 (np.round(-0.001, 2) is -0.0)
constexpr int PK_MAX_RANGE = 65533;        // largest qmax - qmin of a 16-bit row

__device__ __forceinline__ bool pack_q(float v, float p, int& q) {
    // q' = rint(v * p); true if float(q') / p reproduces v bit for bit (the host decodes
    // from the INTEGER q', so -0.0 has its own code and is not handled here)
    const float qf = rintf(__fmul_rn(v, p));
    if (!(fabsf(qf) < 2147483520.0f)) return false;        // inf, or beyond int32
    q = (int)qf;
    return __float_as_uint(__fdiv_rn((float)q, p)) == __float_as_uint(v);
}

__global__ void k_pack_init(spx_pack_row* hdr, int64_t n_rows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) {
        hdr[r].mode = SPX_PACK_U16;
        hdr[r].qmin = INT32_MAX;
        hdr[r].qmax = INT32_MIN;
        hdr[r].n_nan = 0;
    }
}

// pass 1: per-row range of q' and the round-trip verdict
__global__ void __launch_bounds__(PK_THREADS) k_pack_scan(const float* __restrict__ fld,
                                                          int64_t row_len, int64_t ld,
                                                          int64_t seg_len, float p,
                                                          spx_pack_row* __restrict__ hdr) {
    const int64_t row = blockIdx.y;
    const float* __restrict__ base = fld + row * ld;
    const int64_t beg = (int64_t)blockIdx.x * seg_len;
    const int64_t end = min(row_len, beg + seg_len);
    int qmin = INT32_MAX, qmax = INT32_MIN, n_nan = 0, bad = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    auto take = [&](float v) {
        if (v != v) {
            ++n_nan;
        } else if (__float_as_uint(v) == 0x80000000u) {
            // -0.0: its own code, outside the integer range
        } else {
            int q = 0;
            if (pack_q(v, p, q)) {
                qmin = min(qmin, q);
                qmax = max(qmax, q);
            } else {
                bad = 1;
            }
        }
    };
    if (vec) {
        // four independent 16-byte loads in flight per thread (one was latency-bound:
        // 3.0 TB/s)
        constexpr int64_t STEP = (int64_t)PK_THREADS * 4;
        int64_t i = beg + (int64_t)threadIdx.x * 4;
        for (; i + 3 * STEP + 4 <= end; i += 4 * STEP) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = __ldcs(reinterpret_cast<const float4*>(base + i + u * STEP));
#pragma unroll
            for (int u = 0; u < 4; ++u) { take(v[u].x); take(v[u].y); take(v[u].z); take(v[u].w); }
        }
        for (; i < end; i += STEP) {
            if (i + 4 <= end) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(base + i));
                take(v.x); take(v.y); take(v.z); take(v.w);
            } else {
                for (int64_t j = i; j < end; ++j) take(base[j]);
            }
        }
    } else {
        for (int64_t i = beg + threadIdx.x; i < end; i += PK_THREADS) take(base[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qmin = min(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
        qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
        n_nan += __shfl_xor_sync(0xffffffffu, n_nan, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (qmin <= qmax) {
            atomicMin(&hdr[row].qmin, qmin);
            atomicMax(&hdr[row].qmax, qmax);
        }
        if (n_nan) atomicAdd(&hdr[row].n_nan, n_nan);
        if (bad) atomicExch(&hdr[row].mode, SPX_PACK_RAW);
    }
}

__global__ void k_pack_modes(spx_pack_row* hdr, int64_t n_rows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    spx_pack_row h = hdr[r];
    if (h.qmin > h.qmax) {                       // no value at all: every code is NaN
        h.qmin = 0;
        h.qmax = 0;
    }
    if (h.mode != SPX_PACK_RAW && (int64_t)h.qmax - (int64_t)h.qmin > PK_MAX_RANGE)
        h.mode = SPX_PACK_RAW;
    hdr[r] = h;
}

// pass 2: codes of the rows that qualify (8 bytes written per 16 read)
__global__ void __launch_bounds__(PK_THREADS) k_pack_encode(const float* __restrict__ fld,
                                                            int64_t row_len, int64_t ld,
                                                            int64_t seg_len, float p,
                                                            const spx_pack_row* __restrict__ hdr,
                                                            uint16_t* __restrict__ codes,
                                                            int64_t stride) {
    const int64_t row = blockIdx.y;
    const spx_pack_row h = hdr[row];
    if (h.mode != SPX_PACK_U16) return;
    const float* __restrict__ base = fld + row * ld;
    uint16_t* __restrict__ dst = codes + row * stride;
    const int64_t beg = (int64_t)blockIdx.x * seg_len;
    const int64_t end = min(row_len, beg + seg_len);
    auto code = [&](float v) -> uint32_t {
        if (v != v) return PK_NAN;
        if (__float_as_uint(v) == 0x80000000u) return PK_NEGZERO;
        return (uint32_t)((int)rintf(__fmul_rn(v, p)) - h.qmin);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    if (vec) {
        constexpr int64_t STEP = (int64_t)PK_THREADS * 4;
        int64_t i = beg + (int64_t)threadIdx.x * 4;
        for (; i + 3 * STEP + 4 <= end; i += 4 * STEP) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = __ldcs(reinterpret_cast<const float4*>(base + i + u * STEP));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint2 o;
                o.x = code(v[u].x) | (code(v[u].y) << 16);
                o.y = code(v[u].z) | (code(v[u].w) << 16);
                __stcs(reinterpret_cast<uint2*>(dst + i + u * STEP), o);
            }
        }
        for (; i < end; i += STEP) {
            if (i + 4 <= end) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(base + i));
                uint2 o;
                o.x = code(v.x) | (code(v.y) << 16);
                o.y = code(v.z) | (code(v.w) << 16);
                __stcs(reinterpret_cast<uint2*>(dst + i), o);
            } else {
                for (int64_t j = i; j < end; ++j) dst[j] = (uint16_t)code(base[j]);
            }
        }
    } else {
        for (int64_t i = beg + threadIdx.x; i < end; i += PK_THREADS)
            dst[i] = (uint16_t)code(base[i]);
    }
}

static float pack_pow10(int decimals) {
    double pw = 1.0;
    for (int i = 0; i < decimals; ++i) pw *= 10.0;
    return (float)pw;                              // like spx_round_stats_dev
}

// ------------------------------------------------------------------ host decode
__attribute__((target("avx2"))) static void unpack_row_avx2(const uint16_t* __restrict__ src,
                                                            int64_t n, int32_t qmin, float p,
                                                            float* __restrict__ dst) {
    const __m256 vp = _mm256_set1_ps(p);
    const __m256i vq = _mm256_set1_epi32(qmin);
    const __m256i vnan_code = _mm256_set1_epi32(0xFFFF);
    const __m256i vnz_code = _mm256_set1_epi32(0xFFFE);
    const __m256 vnan = _mm256_castsi256_ps(_mm256_set1_epi32(0x7FC00000));
    const __m256 vnz = _mm256_castsi256_ps(_mm256_set1_epi32((int)0x80000000u));
    const bool nt = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128i c16 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m256i c = _mm256_cvtepu16_epi32(c16);
        const __m256 f = _mm256_div_ps(_mm256_cvtepi32_ps(_mm256_add_epi32(c, vq)), vp);
        const __m256 isn = _mm256_castsi256_ps(_mm256_cmpeq_epi32(c, vnan_code));
        const __m256 isz = _mm256_castsi256_ps(_mm256_cmpeq_epi32(c, vnz_code));
        const __m256 r = _mm256_blendv_ps(_mm256_blendv_ps(f, vnz, isz), vnan, isn);
        if (nt) _mm256_stream_ps(dst + i, r);
        else _mm256_storeu_ps(dst + i, r);
    }
    for (; i < n; ++i)
        dst[i] = (src[i] == 0xFFFFu) ? __builtin_nanf("")
                 : (src[i] == 0xFFFEu) ? -0.0f : (float)(qmin + (int32_t)src[i]) / p;
}

static void unpack_row_scalar(const uint16_t* src, int64_t n, int32_t qmin, float p, float* dst) {
    for (int64_t i = 0; i < n; ++i)
        dst[i] = (src[i] == 0xFFFFu) ? __builtin_nanf("")
                 : (src[i] == 0xFFFEu) ? -0.0f : (float)(qmin + (int32_t)src[i]) / p;
}

// =====================================================================================
// Delta transport ("dpack"): the same integer lattice, but variable rate.
//
// A kriged / IDW field is smooth, so neighbouring cells of a row differ by few lattice
// steps: q[c] - q[c-1] needs 0..6 bits where the 16-bit code above spends 16.  A row is cut
// into tiles of 256 cells (one warp, 8 consecutive cells per lane) and segments of 32 tiles
// (one thread block).  Inside a tile the chain f[c] = q[c] (valid cell) / f[c-1] (NaN cell),
// f[-1] = base = q of the first valid cell, is delta coded (first differences, or second
// differences when every cell of the tile is valid and that is smaller; arithmetic modulo
// 2^32, zigzag), every group of 8 cells (a lane) is bit-packed with its own width w in
// {0..12, 14, 16, 32} -- 8 values of w bits are exactly w bytes.  The records of a segment
// follow one another in tile order, each padded to 4 bytes; the block builds them in shared
// memory, reserves the words with ONE atomicAdd and copies them out compacted;
// seg_off[row, segment] points at them (units of 4 bytes).  Tile record:
//   byte 0          mode: bits 0-1 kind (0 every cell NaN, 1 packed, 2 raw f32, 3 constant),
//                   bit 2 NaN bitmap present, bit 3 -0.0 bitmap present, bit 4 second differences
//   kind 1          base int32 | group bitmap uint32 (bit l: lane l has w > 0) | width codes of
//                   those groups, 4 bits each in group order (low nibble first) | [32 B NaN
//                   bitmap: byte l = cells of lane l] | [32 B -0.0 bitmap] | payload, sum of the
//                   widths bytes
//   kind 2          256 raw floats (a value is not on the lattice / beyond int32, or the packed
//                   record would not be smaller)
//   kind 3          base int32 (all 256 cells equal)
// The host decode (spx_dunpack_rows_host) rebuilds float(f) / 10^d: the identical bytes.
//
// The same kernel is the writer's whole output stage when asked (ROUND / STATS): it reads
// the UNROUNDED field once, q = rint(x * 10^d) is np.round's own intermediate
// (interp/steps.py:907-912), so no round-trip check is needed, the per-step statistics of
// interp/main.py:474-525 are reduced from the rounded values on the way, and the rounded
// field is written back only if the caller wants it in HBM as well.
constexpr int DP_TILE = SPX_DPACK_TILE;
constexpr int DP_WARPS = 8;
constexpr int DP_TPW = 4;                                 // tiles per warp
constexpr int DP_SEG_TILES = DP_WARPS * DP_TPW;           // 32
constexpr int DP_SEG_CELLS = DP_SEG_TILES * DP_TILE;      // 8192 == SPX_DPACK_SEGMENT
constexpr int DP_RAW_BYTES = 1 + 4 * DP_TILE;             // mode + 256 floats
constexpr int DP_SLOT_BYTES = (DP_RAW_BYTES + 3) / 4 * 4;   // 1028: a record padded to 4 bytes
static_assert(DP_SEG_CELLS == SPX_DPACK_SEGMENT, "segment size");

__host__ __device__ __forceinline__ int dp_width_of_code(int code) {
    return code < 13 ? code : (code == 13 ? 14 : (code == 14 ? 16 : 32));
}
__device__ __forceinline__ int dp_code_of_width(int w) {
    return w <= 12 ? w : (w <= 14 ? 13 : (w <= 16 ? 14 : 15));
}
__device__ __forceinline__ uint32_t dp_zigzag(int d) {
    return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31);
}
__device__ __forceinline__ void dp_store_u32(uint8_t* at, uint32_t v) {
    at[0] = (uint8_t)v; at[1] = (uint8_t)(v >> 8); at[2] = (uint8_t)(v >> 16); at[3] = (uint8_t)(v >> 24);
}

// Results of the general per-tile path (NaN / -0.0 / off-lattice cells, row tails, input
// that is already rounded): through local memory, these tiles are the minority.
struct DpGen {
    uint32_t z[8];
    float r[8];                      // the (rounded) values of the lane's cells
    int code, wq, pay, base, kind;   // kind: 0 no value at all, 2 raw, 1 otherwise
    uint32_t nzg, m_nan, m_nz, any_nan, any_nz, order2;
    int st_n, st_fin;
    double st_s, st_q;
    float st_mn, st_mx;
};

// first / second differences of a tile whose 256 cells are all valid: fills z / code / wq /
// pay / nzg with the smaller of the two, returns 1 for second differences
__device__ __forceinline__ uint32_t dp_deltas_all_valid(const int (&q)[8], int lane, uint32_t (&z)[8],
                                                        int& base, int& code, int& wq, int& pay,
                                                        uint32_t& nzg) {
    const uint32_t FULL = 0xffffffffu;
    base = __shfl_sync(FULL, q[0], 0);
    int p1 = __shfl_up_sync(FULL, q[7], 1), p2 = __shfl_up_sync(FULL, q[6], 1);
    if (lane == 0) { p1 = base; p2 = base; }
    uint32_t z2[8], or1 = 0, or2 = 0;
    int f = p1, dprev = (int)((uint32_t)p1 - (uint32_t)p2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d1 = (int)((uint32_t)q[j] - (uint32_t)f);
        const int d2 = (int)((uint32_t)d1 - (uint32_t)dprev);
        f = q[j];
        dprev = d1;
        z[j] = dp_zigzag(d1);
        z2[j] = dp_zigzag(d2);
        or1 |= z[j];
        or2 |= z2[j];
    }
    const int c1 = dp_code_of_width(32 - __clz(or1)), c2 = dp_code_of_width(32 - __clz(or2));
    const int w1 = dp_width_of_code(c1), w2 = dp_width_of_code(c2);
    const int pay1 = __reduce_add_sync(FULL, w1), pay2 = __reduce_add_sync(FULL, w2);
    const uint32_t g1 = __ballot_sync(FULL, w1 > 0), g2 = __ballot_sync(FULL, w2 > 0);
    const int s1 = pay1 + ((__popc(g1) + 1) >> 1), s2 = pay2 + ((__popc(g2) + 1) >> 1);
    if (s2 < s1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = z2[j];
        code = c2; wq = w2; pay = pay2; nzg = g2;
        return 1u;
    }
    code = c1; wq = w1; pay = pay1; nzg = g1;
    return 0u;
}

__device__ __noinline__ void dp_tile_general(const float* v, uint32_t in_rng, float p, int round,
                                             int stats, double K, DpGen* g) {
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int q[8];
    uint32_t m_val = 0, m_nan = 0, m_nz = 0, bad = 0;
    int st_n = 0, st_fin = 0;
    double st_s = 0.0, st_q = 0.0;
    float st_mn = CUDART_INF_F, st_mx = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        q[j] = 0;
        g->r[j] = 0.0f;
        if (!((in_rng >> j) & 1u)) continue;
        const float x = v[j];
        float rr = x;
        if (round) {
            const float qf = rintf(__fmul_rn(x, p));
            rr = __fdiv_rn(qf, p);                          // np.round(x, d) in float32
            if (x != x) {
                m_nan |= 1u << j;
            } else if (fabsf(qf) < 2147483520.0f) {
                q[j] = (int)qf;
                m_val |= 1u << j;
                if (__float_as_uint(qf) == 0x80000000u) m_nz |= 1u << j;
            } else {
                bad = 1;                                    // inf, or beyond int32: raw tile
            }
        } else {
            if (x != x) {
                m_nan |= 1u << j;
            } else if (__float_as_uint(x) == 0x80000000u) {
                m_nz |= 1u << j;                            // -0.0: q = 0 in the chain
                m_val |= 1u << j;
            } else {
                int qq = 0;
                if (pack_q(x, p, qq)) { q[j] = qq; m_val |= 1u << j; }
                else bad = 1;
            }
        }
        g->r[j] = rr;
        if (stats && rr == rr) {
            const double d = (double)rr - K;
            st_n += 1;
            st_s += d;
            st_q = fma(d, d, st_q);
            st_mn = fminf(st_mn, rr);
            st_mx = fmaxf(st_mx, rr);
            st_fin += (fabsf(rr) <= 3.402823466e38f) ? 1 : 0;
        }
    }
    g->st_n = st_n; g->st_fin = st_fin; g->st_s = st_s; g->st_q = st_q;
    g->st_mn = st_mn; g->st_mx = st_mx;
    const uint32_t any_bad = __ballot_sync(FULL, bad != 0);
    const uint32_t has_val = __ballot_sync(FULL, m_val != 0);
    g->any_nan = __ballot_sync(FULL, m_nan != 0);
    g->any_nz = __ballot_sync(FULL, m_nz != 0);
    const uint32_t all_val = __ballot_sync(FULL, m_val == 0xFFu);
    g->m_nan = m_nan;
    g->m_nz = m_nz;
    g->kind = any_bad ? 2 : (has_val ? 1 : 0);
    uint32_t z[8];
    int base, code, wq, pay;
    uint32_t nzg;
    g->order2 = 0;
    if (all_val == FULL) {
        g->order2 = dp_deltas_all_valid(q, lane, z, base, code, wq, pay, nzg);
    } else {
        // value in front of this lane's first cell: last valid q of the lower lanes
        int lastq = 0, firstq = 0;
#pragma unroll
        for (int j = 7; j >= 0; --j)
            if ((m_val >> j) & 1u) firstq = q[j];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if ((m_val >> j) & 1u) lastq = q[j];
        int has = m_val != 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int h2 = __shfl_up_sync(FULL, has, o);
            const int q2 = __shfl_up_sync(FULL, lastq, o);
            if (lane >= o && !has) { has = h2; lastq = q2; }
        }
        const int hx = __shfl_up_sync(FULL, has, 1);
        const int qx = __shfl_up_sync(FULL, lastq, 1);
        base = __shfl_sync(FULL, firstq, has_val ? __ffs(has_val) - 1 : 0);
        int f = (lane > 0 && hx) ? qx : base;
        uint32_t orz = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int d = 0;
            if ((m_val >> j) & 1u) {
                d = (int)((uint32_t)q[j] - (uint32_t)f);
                f = q[j];
            }
            z[j] = dp_zigzag(d);
            orz |= z[j];
        }
        code = dp_code_of_width(32 - __clz(orz));
        wq = dp_width_of_code(code);
        pay = __reduce_add_sync(FULL, wq);
        nzg = __ballot_sync(FULL, wq > 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g->z[j] = z[j];
    g->code = code; g->wq = wq; g->pay = pay; g->base = base; g->nzg = nzg;
}

__device__ __forceinline__ void dp_load_tile(const float* rowp, int64_t c0, int64_t row_len,
                                             int vec_ok, float (&v)[8], uint32_t& in_rng) {
    if (vec_ok && c0 + 8 <= row_len) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(rowp + c0));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(rowp + c0) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        in_rng = 0xFFu;
    } else {
        in_rng = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = c0 + j < row_len;
            v[j] = in ? rowp[c0 + j] : 0.0f;
            in_rng |= (uint32_t)in << j;
        }
    }
}

template <bool ROUND, bool STATS>
__global__ void __launch_bounds__(DP_WARPS * 32, 3) k_dpack(
    const float* fld, float* wb, int64_t row_len, int64_t ld,
    int64_t segs_per_row, float p, int vec_ok, uint32_t* __restrict__ seg_off,
    uint32_t* __restrict__ payload, unsigned long long cap_words,
    unsigned long long* __restrict__ counters, StatPart* __restrict__ parts) {
    // one 4-byte aligned slot per tile of the segment; the records are compacted on the way out
    __shared__ __align__(16) uint8_t stage[DP_SEG_TILES * DP_SLOT_BYTES];
    __shared__ int rec_words[DP_SEG_TILES];
    __shared__ int rec_pos[DP_SEG_TILES + 1];
    __shared__ unsigned long long blk_base;
    __shared__ StatPart sp[DP_WARPS];
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = blockIdx.x / segs_per_row;
    const int64_t seg = blockIdx.x - row * segs_per_row;
    const float* rowp = fld + row * ld;
    const int64_t seg_c0 = seg * DP_SEG_CELLS;

    // statistics: shifted sums against the segment's first (rounded) value
    double K = 0.0, st_s = 0.0, st_q = 0.0;
    float st_mn = CUDART_INF_F, st_mx = -CUDART_INF_F;
    int st_n = 0, st_fin = 0;
    if (STATS) {
        float x0 = rowp[seg_c0];
        if (ROUND) x0 = __fdiv_rn(rintf(__fmul_rn(x0, p)), p);
        K = (double)x0;
        if (!(fabs(K) < 1.0e300)) K = 0.0;
    }

    float v[8], vn[8];
    uint32_t in_rng, in_rng_n = 0;
    dp_load_tile(rowp, seg_c0 + (int64_t)warp * DP_TILE + lane * 8, row_len, vec_ok, v, in_rng);
#pragma unroll 1
    for (int i = 0; i < DP_TPW; ++i) {
        const int tile = i * DP_WARPS + warp;
        const int64_t tile_c0 = seg_c0 + (int64_t)tile * DP_TILE;
        if (i + 1 < DP_TPW)                        // the next tile's loads fly during this one
            dp_load_tile(rowp, tile_c0 + (int64_t)DP_WARPS * DP_TILE + lane * 8, row_len, vec_ok,
                         vn, in_rng_n);
        if (tile_c0 >= row_len) {                  // past the end of the row: no record
            if (lane == 0) rec_words[tile] = 0;
        } else {
            uint32_t z[8];
            int base = 0, code = 0, wq = 0, pay = 0, kind = 1;
            uint32_t nzg = 0, any_nan = 0, any_nz = 0, order2 = 0, m_nan = 0, m_nz = 0;
            bool fast = false;
            if (ROUND) {
                // fast path: every cell of the tile exists, is finite, on the int32 lattice
                // and not -0.0 -- then q = rint(x * 10^d) needs no further check
                float qf[8];
                uint32_t amax = 0, nzmin = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    qf[j] = rintf(__fmul_rn(v[j], p));
                    const uint32_t u = __float_as_uint(qf[j]);
                    amax = max(amax, u & 0x7fffffffu);
                    nzmin = min(nzmin, u ^ 0x80000000u);
                }
                fast = __all_sync(FULL, in_rng == 0xFFu && amax < 0x4f000000u && nzmin != 0u);
                if (fast) {
                    int q[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) q[j] = (int)qf[j];
                    if (STATS || wb != nullptr) {
                        float r[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) r[j] = __fdiv_rn(qf[j], p);
                        if (STATS) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const double d = (double)r[j] - K;
                                st_s += d;
                                st_q = fma(d, d, st_q);
                                st_mn = fminf(st_mn, r[j]);
                                st_mx = fmaxf(st_mx, r[j]);
                            }
                            st_n += 8;
                            st_fin += 8;
                        }
                        if (wb != nullptr) {
                            float* o = wb + row * ld + tile_c0 + lane * 8;
                            if (vec_ok) {
                                __stcs(reinterpret_cast<float4*>(o),
                                       make_float4(r[0], r[1], r[2], r[3]));
                                __stcs(reinterpret_cast<float4*>(o) + 1,
                                       make_float4(r[4], r[5], r[6], r[7]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j] = r[j];
                            }
                        }
                    }
                    order2 = dp_deltas_all_valid(q, lane, z, base, code, wq, pay, nzg);
                }
            }
            if (!fast) {
                float tmp[8];
                DpGen g;
#pragma unroll
                for (int j = 0; j < 8; ++j) tmp[j] = v[j];
                dp_tile_general(tmp, in_rng, p, ROUND ? 1 : 0, STATS ? 1 : 0, K, &g);
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] = g.z[j];
                base = g.base; code = g.code; wq = g.wq; pay = g.pay; kind = g.kind;
                nzg = g.nzg; any_nan = g.any_nan; any_nz = g.any_nz; order2 = g.order2;
                m_nan = g.m_nan;
                m_nz = g.m_nz;
                if (STATS) {
                    st_n += g.st_n; st_fin += g.st_fin; st_s += g.st_s; st_q += g.st_q;
                    st_mn = fminf(st_mn, g.st_mn); st_mx = fmaxf(st_mx, g.st_mx);
                }
                if (ROUND && wb != nullptr) {
                    float* o = wb + row * ld + tile_c0 + lane * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((in_rng >> j) & 1u) o[j] = g.r[j];
                }
                if (ROUND) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = g.r[j];       // raw tiles carry rounded values
                }
            }
            int bytes;
            if (kind == 2) { bytes = DP_RAW_BYTES; }
            else if (kind == 0) { bytes = 1; }
            else if (pay == 0 && !any_nan && !any_nz) { kind = 3; bytes = 5; }
            else {
                bytes = 9 + ((__popc(nzg) + 1) >> 1) + (any_nan ? 32 : 0) + (any_nz ? 32 : 0) + pay;
                if (bytes >= DP_RAW_BYTES) { kind = 2; bytes = DP_RAW_BYTES; }
            }
            const int words = (bytes + 3) >> 2;
            uint8_t* rb = stage + tile * DP_SLOT_BYTES;
            if (lane == 0) {
                rec_words[tile] = words;
                rb[0] = (kind != 1) ? (uint8_t)kind
                                    : (uint8_t)(1u | (any_nan ? 4u : 0u) | (any_nz ? 8u : 0u) |
                                                (order2 << 4));
                if (kind == 1 || kind == 3) dp_store_u32(rb + 1, (uint32_t)base);
                if (kind == 1) dp_store_u32(rb + 5, nzg);
                for (int k = bytes; k < words * 4; ++k) rb[k] = 0;     // padding to 4 bytes
            }
            if (kind == 2) {
                uint8_t* o = rb + 1 + lane * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) dp_store_u32(o + 4 * j, __float_as_uint(v[j]));
            } else if (kind == 1) {
                {   // width codes of the non-empty groups, two per byte
                    const int idx = __popc(nzg & ((1u << lane) - 1u));
                    const uint32_t above = (lane < 31) ? (nzg & ~((2u << lane) - 1u)) : 0u;
                    const int nxt = above ? __ffs(above) - 1 : 0;
                    const int c_hi = __shfl_sync(FULL, code, nxt);
                    if (wq > 0 && (idx & 1) == 0)
                        rb[9 + (idx >> 1)] = (uint8_t)(code | ((above ? c_hi : 0) << 4));
                }
                int off = wq;                      // exclusive prefix sum of the widths
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, off, o);
                    if (lane >= o) off += t;
                }
                off -= wq;
                int pos = 9 + ((__popc(nzg) + 1) >> 1);
                if (any_nan) { rb[pos + lane] = (uint8_t)m_nan; pos += 32; }
                if (any_nz) { rb[pos + lane] = (uint8_t)m_nz; pos += 32; }
                uint8_t* out = rb + pos + off;
                if (wq == 32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) dp_store_u32(out + 4 * j, z[j]);
                } else if (wq > 0) {
                    const int w = wq;                          // 1..16
                    const uint32_t a0 = z[0] | (z[1] << w), a1 = z[2] | (z[3] << w);
                    const uint32_t a2 = z[4] | (z[5] << w), a3 = z[6] | (z[7] << w);
                    const uint64_t b0 = (uint64_t)a0 | ((uint64_t)a1 << (2 * w));
                    const uint64_t b1 = (uint64_t)a2 | ((uint64_t)a3 << (2 * w));
                    uint64_t lo, hi;
                    if (w == 16) { lo = b0; hi = b1; }
                    else { lo = b0 | (b1 << (4 * w)); hi = b1 >> (64 - 4 * w); }
                    const uint32_t l0 = (uint32_t)lo, l1 = (uint32_t)(lo >> 32);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < w) out[k] = (uint8_t)(l0 >> (8 * k));
                    if (w > 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k + 4 < w) out[k + 4] = (uint8_t)(l1 >> (8 * k));
                        for (int k = 8; k < w; ++k) { out[k] = (uint8_t)hi; hi >>= 8; }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = vn[j];
        in_rng = in_rng_n;
    }
    if (STATS) {
        // one shift K for the whole block: the sums simply add up
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_n += __shfl_xor_sync(FULL, st_n, o);
            st_fin += __shfl_xor_sync(FULL, st_fin, o);
            st_s += __shfl_xor_sync(FULL, st_s, o);
            st_q += __shfl_xor_sync(FULL, st_q, o);
            st_mn = fminf(st_mn, __shfl_xor_sync(FULL, st_mn, o));
            st_mx = fmaxf(st_mx, __shfl_xor_sync(FULL, st_mx, o));
        }
        if (lane == 0) {
            StatPart a;
            a.n = (double)st_n; a.mean = st_s; a.m2 = st_q;      // raw sums, finished below
            a.mn = (double)st_mn; a.mx = (double)st_mx; a.nfin = (double)st_fin;
            sp[warp] = a;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const int b = rec_words[lane];
        int inc = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        rec_pos[lane] = inc - b;
        if (lane == 31) {
            rec_pos[32] = inc;
            blk_base = atomicAdd(&counters[0], (unsigned long long)inc);
        }
    }
    if (STATS && threadIdx.x == 32) {
        StatPart t = sp[0];
        for (int w = 1; w < DP_WARPS; ++w) {
            t.n += sp[w].n; t.mean += sp[w].mean; t.m2 += sp[w].m2; t.nfin += sp[w].nfin;
            t.mn = fmin(t.mn, sp[w].mn); t.mx = fmax(t.mx, sp[w].mx);
        }
        const double sum = t.mean, sq = t.m2;
        t.mean = (t.n > 0.0) ? K + sum / t.n : 0.0;
        t.m2 = (t.n > 0.0) ? sq - sum * (sum / t.n) : 0.0;
        parts[blockIdx.x] = t;
    }
    __syncthreads();
    const unsigned long long at = blk_base;
    if (at + (unsigned long long)rec_pos[DP_SEG_TILES] > cap_words) {   // does not fit: flagged
        if (threadIdx.x == 0) {
            seg_off[blockIdx.x] = 0xFFFFFFFFu;
            atomicExch(&counters[1], 1ull);
        }
        return;
    }
    if (threadIdx.x == 0) seg_off[blockIdx.x] = (uint32_t)at;
    // the records, compacted in tile order: every warp moves the four it wrote
#pragma unroll 1
    for (int i = 0; i < DP_TPW; ++i) {
        const int tile = i * DP_WARPS + warp;
        const int words = rec_words[tile];
        const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(stage + tile * DP_SLOT_BYTES);
        uint32_t* __restrict__ dst = payload + at + rec_pos[tile];
        for (int k = lane; k < words; k += 32) dst[k] = src[k];
    }
}

// ------------------------------------------------------------------ host decode (dpack)
// one tile record at rec (n cells of it exist); returns the next record, nullptr if malformed
static const uint8_t* dunpack_tile(const uint8_t* rec, const uint8_t* end, int n, float p,
                                   float* out) {
    if (rec >= end) return nullptr;
    const uint32_t mode = rec[0];
    const int kind = (int)(mode & 3u);
    const float nanv = __builtin_nanf("");
    if (kind == 0) {
        for (int c = 0; c < n; ++c) out[c] = nanv;
        return rec + 4;
    }
    if (kind == 2) {
        if (rec + DP_SLOT_BYTES > end) return nullptr;
        memcpy(out, rec + 1, sizeof(float) * (size_t)n);
        return rec + DP_SLOT_BYTES;
    }
    if (rec + 5 > end) return nullptr;
    int32_t f;
    memcpy(&f, rec + 1, 4);
    if (kind == 3) {
        const float x = (float)f / p;
        for (int c = 0; c < n; ++c) out[c] = x;
        return rec + 8;
    }
    if (rec + 9 > end) return nullptr;
    uint32_t nzg;
    memcpy(&nzg, rec + 5, 4);
    const uint8_t* nib = rec + 9;
    const uint8_t* pay = nib + ((__builtin_popcount(nzg) + 1) >> 1);
    const uint8_t* bm_nan = nullptr;
    const uint8_t* bm_nz = nullptr;
    if (mode & 4u) { bm_nan = pay; pay += 32; }
    if (mode & 8u) { bm_nz = pay; pay += 32; }
    if (pay > end) return nullptr;
    const bool order2 = (mode & 16u) != 0;
    int32_t d1 = 0;
    int k = 0;                                     // index among the non-empty groups
    for (int l = 0; l < 32; ++l) {
        int w = 0;
        if ((nzg >> l) & 1u) {
            w = dp_width_of_code((nib[k >> 1] >> ((k & 1) * 4)) & 15);
            ++k;
        }
        if (pay + w > end) return nullptr;
        uint32_t z[8];
        if (w == 32) {
            memcpy(z, pay, 32);
        } else if (w == 0) {
            for (int j = 0; j < 8; ++j) z[j] = 0;
        } else {
            unsigned __int128 bits = 0;
            memcpy(&bits, pay, (size_t)w);
            const uint32_t mask = (1u << w) - 1u;
            for (int j = 0; j < 8; ++j) {
                z[j] = (uint32_t)bits & mask;
                bits >>= w;
            }
        }
        pay += w;
        const int cnt = (n - l * 8) < 8 ? (n - l * 8) : 8;
        if (cnt <= 0) continue;                    // groups past the end of the row: zero deltas
        const uint32_t mn = bm_nan ? bm_nan[l] : 0u, mz = bm_nz ? bm_nz[l] : 0u;
        float* o = out + l * 8;
        for (int j = 0; j < cnt; ++j) {
            const int32_t d = (int32_t)(z[j] >> 1) ^ -(int32_t)(z[j] & 1u);
            if (order2) {
                d1 = (int32_t)((uint32_t)d1 + (uint32_t)d);
                f = (int32_t)((uint32_t)f + (uint32_t)d1);
            } else {
                f = (int32_t)((uint32_t)f + (uint32_t)d);
            }
            o[j] = ((mn >> j) & 1u) ? nanv : (((mz >> j) & 1u) ? -0.0f : (float)f / p);
        }
    }
    return rec + (((pay - rec) + 3) & ~(ptrdiff_t)3);          // records are padded to 4 bytes
}

// The same decoder with AVX2 + BMI2: a group of 8 cells is unpacked with PDEP, zigzag-decoded,
// prefix-summed (twice for second differences) and divided 8 lanes at a time.
#define SPX_AVX2_BMI2 __attribute__((target("avx2,bmi2")))

SPX_AVX2_BMI2 static inline __m256i dp_prefix8(__m256i d) {
    __m256i s = _mm256_add_epi32(d, _mm256_slli_si256(d, 4));
    s = _mm256_add_epi32(s, _mm256_slli_si256(s, 8));                  // prefix inside each half
    const __m256i lo_tot = _mm256_permutevar8x32_epi32(s, _mm256_set1_epi32(3));
    return _mm256_add_epi32(s, _mm256_blend_epi32(_mm256_setzero_si256(), lo_tot, 0xF0));
}

SPX_AVX2_BMI2 static const uint8_t* dunpack_tile_avx2(const uint8_t* rec, const uint8_t* end, int n,
                                                      float p, float* out) {
    if (rec >= end) return nullptr;
    const uint32_t mode = rec[0];
    const int kind = (int)(mode & 3u);
    if (kind != 1 && kind != 3) return dunpack_tile(rec, end, n, p, out);
    if (rec + 5 > end) return nullptr;
    int32_t f;
    memcpy(&f, rec + 1, 4);
    const __m256 vp = _mm256_set1_ps(p);
    if (kind == 3) {
        const __m256 x = _mm256_set1_ps((float)f / p);
        int c = 0;
        for (; c + 8 <= n; c += 8) _mm256_storeu_ps(out + c, x);
        for (; c < n; ++c) out[c] = (float)f / p;
        return rec + 8;
    }
    if (rec + 9 > end) return nullptr;
    uint32_t nzg;
    memcpy(&nzg, rec + 5, 4);
    const uint8_t* nib = rec + 9;
    const uint8_t* pay = nib + ((__builtin_popcount(nzg) + 1) >> 1);
    const uint8_t* bm_nan = nullptr;
    const uint8_t* bm_nz = nullptr;
    if (mode & 4u) { bm_nan = pay; pay += 32; }
    if (mode & 8u) { bm_nz = pay; pay += 32; }
    if (pay > end) return nullptr;
    const bool order2 = (mode & 16u) != 0;
    const float nanv = __builtin_nanf("");
    const __m256i one = _mm256_set1_epi32(1);
    const __m256i ramp = _mm256_setr_epi32(1, 2, 3, 4, 5, 6, 7, 8);
    int32_t d1 = 0;
    int k = 0;
    for (int l = 0; l < 32; ++l) {
        int w = 0;
        if ((nzg >> l) & 1u) {
            w = dp_width_of_code((nib[k >> 1] >> ((k & 1) * 4)) & 15);
            ++k;
        }
        if (pay + w > end) return nullptr;
        const int cnt = (n - l * 8) < 8 ? (n - l * 8) : 8;
        __m256i fv;
        if (w == 0) {
            if (cnt <= 0) continue;
            if (!order2 || d1 == 0) {
                fv = _mm256_set1_epi32(f);
            } else {
                fv = _mm256_add_epi32(_mm256_set1_epi32(f),
                                      _mm256_mullo_epi32(_mm256_set1_epi32(d1), ramp));
                f = (int32_t)((uint32_t)f + 8u * (uint32_t)d1);
            }
        } else {
            __m256i z;
            if (w == 32) {
                z = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(pay));
            } else {
                uint64_t lo = 0, hi = 0;
                if (pay + 16 <= end) {
                    memcpy(&lo, pay, 8);
                    memcpy(&hi, pay + 8, 8);
                } else {
                    uint8_t tmp[16] = {0};
                    memcpy(tmp, pay, (size_t)w);
                    memcpy(&lo, tmp, 8);
                    memcpy(&hi, tmp + 8, 8);
                }
                if (w <= 8) {
                    const uint64_t m = 0x0101010101010101ull * ((1ull << w) - 1ull);
                    z = _mm256_cvtepu8_epi32(_mm_cvtsi64_si128((long long)_pdep_u64(lo, m)));
                } else {
                    const uint64_t m = 0x0001000100010001ull * ((1ull << w) - 1ull);
                    const uint64_t a = lo;
                    const uint64_t b = (w == 16) ? hi : ((lo >> (4 * w)) | (hi << (64 - 4 * w)));
                    const __m128i x = _mm_set_epi64x((long long)_pdep_u64(b, m),
                                                     (long long)_pdep_u64(a, m));
                    z = _mm256_cvtepu16_epi32(x);
                }
            }
            pay += w;
            if (cnt <= 0) continue;
            const __m256i d = _mm256_xor_si256(
                _mm256_srli_epi32(z, 1),
                _mm256_sub_epi32(_mm256_setzero_si256(), _mm256_and_si256(z, one)));
            __m256i s = dp_prefix8(d);
            if (order2) {
                const __m256i d1v = _mm256_add_epi32(_mm256_set1_epi32(d1), s);
                d1 = _mm256_extract_epi32(d1v, 7);
                s = dp_prefix8(d1v);
            }
            fv = _mm256_add_epi32(_mm256_set1_epi32(f), s);
            f = _mm256_extract_epi32(fv, 7);
        }
        const __m256 x = _mm256_div_ps(_mm256_cvtepi32_ps(fv), vp);
        float* o = out + l * 8;
        const uint32_t mn = bm_nan ? bm_nan[l] : 0u, mz = bm_nz ? bm_nz[l] : 0u;
        if (cnt == 8 && !(mn | mz)) {
            _mm256_storeu_ps(o, x);
        } else {
            float tmp[8];
            _mm256_storeu_ps(tmp, x);
            for (int j = 0; j < cnt; ++j)
                o[j] = ((mn >> j) & 1u) ? nanv : (((mz >> j) & 1u) ? -0.0f : tmp[j]);
        }
    }
    return rec + (((pay - rec) + 3) & ~(ptrdiff_t)3);
}

}  // namespace spx

using namespace spx;

extern "C" {

int64_t spx_pack_stride(int64_t row_len) { return row_len < 0 ? 0 : (row_len + 7) / 8 * 8; }

int spx_pack_field_dev(const float* fld, int64_t n_rows, int64_t row_len, int64_t ld,
                       int32_t decimals, spx_pack_row* hdr, uint16_t* codes, void* stream) {
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    if (!fld || !hdr || !codes || ld < row_len || decimals < 0 || decimals > 9) {
        set_error("pack_field: bad argument (decimals must be 0..9)");
        return SPX_EINVAL;
    }
    if (n_rows > 65535) {
        set_error("pack_field: more than 65535 rows in one call");
        return SPX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const float p = pack_pow10(decimals);
    // enough blocks to fill the GPU a few times over, at least 4096 elements each
    int64_t n_seg = (148 * 64 + n_rows - 1) / n_rows;     // many more blocks than slots: no tail
    const int64_t max_seg = (row_len + 4095) / 4096;
    if (n_seg > max_seg) n_seg = max_seg;
    if (n_seg < 1) n_seg = 1;
    int64_t seg_len = (row_len + n_seg - 1) / n_seg;
    seg_len = (seg_len + 3) / 4 * 4;
    const unsigned rb = (unsigned)((n_rows + 255) / 256);
    dim3 grid((unsigned)n_seg, (unsigned)n_rows);
    k_pack_init<<<rb, 256, 0, st>>>(hdr, n_rows);
    k_pack_scan<<<grid, PK_THREADS, 0, st>>>(fld, row_len, ld, seg_len, p, hdr);
    k_pack_modes<<<rb, 256, 0, st>>>(hdr, n_rows);
    k_pack_encode<<<grid, PK_THREADS, 0, st>>>(fld, row_len, ld, seg_len, p, hdr, codes,
                                              spx_pack_stride(row_len));
    SPX_CHECK_LAUNCH("k_pack_*");
    return SPX_OK;
}

int spx_unpack_field_host(const spx_pack_row* hdr, const uint16_t* codes, int64_t n_rows,
                          int64_t row_len, int32_t decimals, float* out, int64_t out_ld,
                          int32_t n_threads) {
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    if (!hdr || !codes || !out || out_ld < row_len || decimals < 0 || decimals > 9) {
        set_error("unpack_field: bad argument");
        return SPX_EINVAL;
    }
    const float p = pack_pow10(decimals);
    const int64_t stride = spx_pack_stride(row_len);
    const bool avx2 = __builtin_cpu_supports("avx2") != 0;
    auto work = [&](int part, int n_parts) {
        // interleaved rows: neighbouring threads write neighbouring rows
        for (int64_t r = part; r < n_rows; r += n_parts) {
            if (hdr[r].mode != SPX_PACK_U16) continue;       // raw rows are copied by the caller
            if (avx2)
                unpack_row_avx2(codes + r * stride, row_len, hdr[r].qmin, p, out + r * out_ld);
            else
                unpack_row_scalar(codes + r * stride, row_len, hdr[r].qmin, p, out + r * out_ld);
        }
        if (avx2) _mm_sfence();
    };
    HostPool::get().run(work, n_threads);
    return SPX_OK;
}

int64_t spx_dpack_segments(int64_t row_len) {
    return row_len <= 0 ? 0 : (row_len + DP_SEG_CELLS - 1) / DP_SEG_CELLS;
}

int64_t spx_dpack_capacity(int64_t n_rows, int64_t row_len) {
    if (n_rows <= 0 || row_len <= 0) return 0;
    return n_rows * spx_dpack_segments(row_len) * (int64_t)(DP_SEG_TILES * DP_SLOT_BYTES);
}

int64_t spx_dpack_stats_workspace(int64_t n_rows, int64_t row_len) {
    if (n_rows <= 0 || row_len <= 0) return 0;
    return n_rows * spx_dpack_segments(row_len) * (int64_t)sizeof(StatPart);
}

int spx_dpack_field_dev(float* fld, int64_t n_rows, int64_t row_len, int64_t ld,
                        int32_t decimals, int32_t flags, double* stats, void* workspace,
                        uint32_t* seg_off, void* payload, int64_t capacity_bytes,
                        uint64_t* counters, void* stream) {
    if (n_rows < 0 || row_len < 0 || !counters) {
        set_error("dpack_field: bad argument");
        return SPX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(uint64_t), st);
    if (e != cudaSuccess) {
        set_error(cudaGetErrorString(e));
        return SPX_ECUDA;
    }
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    const bool do_round = (flags & SPX_DPACK_ROUND) != 0;
    const bool write_back = (flags & SPX_DPACK_WRITE_BACK) != 0;
    if (!fld || !seg_off || !payload || ld < row_len || decimals < 0 || decimals > 9 ||
        capacity_bytes < 0 || (reinterpret_cast<uintptr_t>(payload) & 3) ||
        (stats && !workspace) || (write_back && !do_round)) {
        set_error("dpack_field: bad argument (decimals 0..9, payload 4-byte aligned, workspace "
                  "with stats, write-back only with rounding)");
        return SPX_EINVAL;
    }
    unsigned long long cap_words = (unsigned long long)capacity_bytes / 4;
    if (cap_words > 0xFFFFFFFEull) cap_words = 0xFFFFFFFEull;      // offsets are 32-bit words
    const int64_t segs = spx_dpack_segments(row_len);
    const int64_t n_blk = n_rows * segs;
    if (n_blk > 0x7FFFFFFFll) {
        set_error("dpack_field: field too large for one call");
        return SPX_EINVAL;
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(fld) & 15) == 0) && (ld % 4 == 0);
    const float p = pack_pow10(decimals);
    uint32_t* pay = reinterpret_cast<uint32_t*>(payload);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counters);
    StatPart* parts = reinterpret_cast<StatPart*>(workspace);
    float* wb = write_back ? fld : nullptr;
    const dim3 grid((unsigned)n_blk), block(DP_WARPS * 32);
#define SPX_DPACK_LAUNCH(R, S)                                                                   \
    k_dpack<R, S><<<grid, block, 0, st>>>(fld, wb, row_len, ld, segs, p, vec_ok, seg_off, pay,   \
                                          cap_words, cnt, parts)
    if (do_round) {
        if (stats) SPX_DPACK_LAUNCH(true, true);
        else SPX_DPACK_LAUNCH(true, false);
    } else {
        if (stats) SPX_DPACK_LAUNCH(false, true);
        else SPX_DPACK_LAUNCH(false, false);
    }
#undef SPX_DPACK_LAUNCH
    SPX_CHECK_LAUNCH("k_dpack");
    if (stats) {
        launch_stats_final(parts, (int)segs, n_rows, stats, st);
        SPX_CHECK_LAUNCH("k_stats_final");
    }
    return SPX_OK;
}

int spx_dunpack_rows_host(const uint32_t* seg_off, const void* payload, int64_t payload_bytes,
                          int64_t n_rows, int64_t row_len, int32_t decimals, float* out,
                          int64_t out_ld, int32_t n_threads) {
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    if (!seg_off || !payload || !out || out_ld < row_len || decimals < 0 || decimals > 9 ||
        payload_bytes < 0) {
        set_error("dunpack_rows: bad argument");
        return SPX_EINVAL;
    }
    const float p = pack_pow10(decimals);
    const int64_t segs = spx_dpack_segments(row_len);
    const uint8_t* pay = reinterpret_cast<const uint8_t*>(payload);
    const uint8_t* pay_end = pay + payload_bytes;
    int bad = 0;
    static const bool scalar_only = []() {
        const char* e = std::getenv("SPX_DUNPACK_SCALAR");
        return e && e[0] == '1';
    }();
    const bool fast = !scalar_only && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    auto work = [&](int part, int n_parts) {
        for (int64_t r = part; r < n_rows; r += n_parts) {
            float* o = out + r * out_ld;
            for (int64_t sg = 0; sg < segs; ++sg) {
                const uint32_t at = seg_off[r * segs + sg];
                const uint8_t* rec = (at == 0xFFFFFFFFu || (int64_t)at * 4 > payload_bytes)
                                         ? nullptr : pay + (int64_t)at * 4;
                for (int64_t c0 = sg * DP_SEG_CELLS;
                     rec && c0 < row_len && c0 < (sg + 1) * DP_SEG_CELLS; c0 += DP_TILE) {
                    const int64_t rest = row_len - c0;
                    const int nc = rest < DP_TILE ? (int)rest : DP_TILE;
                    rec = fast ? dunpack_tile_avx2(rec, pay_end, nc, p, o + c0)
                               : dunpack_tile(rec, pay_end, nc, p, o + c0);
                }
                if (!rec) __atomic_store_n(&bad, 1, __ATOMIC_RELAXED);
            }
        }
    };
    if (n_threads == 1 || n_rows == 1) work(0, 1);
    else HostPool::get().run(work, n_threads);
    if (bad) {
        set_error("dunpack_rows: malformed or overflowed record");
        return SPX_EINVAL;
    }
    return SPX_OK;
}

}  // extern "C"
