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
// into tiles of 256 cells (one warp, 8 consecutive cells per lane).  Inside a tile the
// chain f[c] = q[c] (valid cell) / f[c-1] (NaN cell), f[-1] = base = q of the first valid
// cell, is delta coded (d = f[c] - f[c-1] modulo 2^32, zigzag), every group of 8 cells
// (a lane) is bit-packed with its own width w in {0..12, 14, 16, 32} -- 8 values of w bits
// are exactly w bytes -- and the tile becomes ONE variable-size record in a payload
// buffer; records are allocated by one atomicAdd per block, tile_off[row, tile] points at
// them (units of 4 bytes).  Record:
//   word 0          mode: bits 0-1 kind (0 every cell NaN, 1 packed, 2 raw f32, 3 constant),
//                   bit 2 NaN bitmap present, bit 3 -0.0 bitmap present
//   kind 1          base int32 | 16 B width nibbles (lane l: byte l/2, low nibble = even
//                   lane) | [32 B NaN bitmap: byte l = cells of lane l] | [32 B -0.0
//                   bitmap] | payload, sum of the widths bytes | zero padding to 4 B
//   kind 2          256 raw floats (a cell failed the bit-exact round-trip check, or the
//                   packed record would not be smaller)
//   kind 3          base int32 (all 256 cells equal)
// The host decode (spx_dunpack_rows_host) rebuilds float(f) / 10^d: the identical bytes.
constexpr int DP_TILE = SPX_DPACK_TILE;
constexpr int DP_WARPS = 8;
constexpr int DP_RAW_WORDS = 1 + DP_TILE;                 // mode + 256 floats
constexpr int DP_STAGE_WORDS = 280;                       // >= 2 + 4 + 8 + 8 + 256 + slack

__host__ __device__ __forceinline__ int dp_width_of_code(int code) {
    return code < 13 ? code : (code == 13 ? 14 : (code == 14 ? 16 : 32));
}
__device__ __forceinline__ int dp_code_of_width(int w) {
    return w <= 12 ? w : (w <= 14 ? 13 : (w <= 16 ? 14 : 15));
}

__global__ void __launch_bounds__(DP_WARPS * 32) k_dpack(
    const float* __restrict__ fld, int64_t row_len, int64_t ld, int64_t tiles_per_row,
    int64_t n_tiles, float p, int vec_ok, uint32_t* __restrict__ tile_off,
    uint32_t* __restrict__ payload, unsigned long long cap_words,
    unsigned long long* __restrict__ counters) {
    __shared__ uint32_t stage[DP_WARPS][DP_STAGE_WORDS];
    __shared__ uint32_t rec_words[DP_WARPS];
    __shared__ unsigned long long blk_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * DP_WARPS + warp;
    const bool live = g < n_tiles;
    const int64_t row = live ? g / tiles_per_row : 0;
    const int64_t c0 = live ? (g - row * tiles_per_row) * DP_TILE + lane * 8 : 0;
    const float* __restrict__ src = fld + row * ld + c0;

    float v[8];
    uint32_t in_rng = 0;                     // bit j: cell c0 + j exists
    if (live) {
        if (vec_ok && c0 + 8 <= row_len) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
            const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            in_rng = 0xFFu;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool in = c0 + j < row_len;
                v[j] = in ? src[j] : 0.0f;
                in_rng |= (uint32_t)in << j;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    }
    int q[8];
    uint32_t m_val = 0, m_nan = 0, m_nz = 0, bad = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        q[j] = 0;
        if (!((in_rng >> j) & 1u)) continue;
        const float x = v[j];
        if (x != x) {
            m_nan |= 1u << j;
        } else if (__float_as_uint(x) == 0x80000000u) {
            m_nz |= 1u << j;                 // -0.0: q = 0 in the chain, sign from the bitmap
            m_val |= 1u << j;
        } else {
            int qq = 0;
            if (pack_q(x, p, qq)) {
                q[j] = qq;
                m_val |= 1u << j;
            } else {
                bad = 1;
            }
        }
    }
    const uint32_t FULL = 0xffffffffu;
    const uint32_t any_bad = __ballot_sync(FULL, bad != 0);
    const uint32_t has_val = __ballot_sync(FULL, m_val != 0);
    const uint32_t any_nan = __ballot_sync(FULL, m_nan != 0);
    const uint32_t any_nz = __ballot_sync(FULL, m_nz != 0);
    const uint32_t all_val = __ballot_sync(FULL, m_val == 0xFFu);

    // value in front of this lane's first cell: last valid q of the lower lanes, else base
    int prev, base;
    if (all_val == FULL) {
        base = __shfl_sync(FULL, q[0], 0);
        prev = __shfl_up_sync(FULL, q[7], 1);
        if (lane == 0) prev = base;
    } else {
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
        prev = (lane > 0 && hx) ? qx : base;
    }
    uint32_t z[8], orz = 0;
    {
        int f = prev;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int d = 0;
            if ((m_val >> j) & 1u) {
                d = (int)((uint32_t)q[j] - (uint32_t)f);
                f = q[j];
            }
            z[j] = ((uint32_t)d << 1) ^ (uint32_t)(d >> 31);
            orz |= z[j];
        }
    }
    const int code = dp_code_of_width(32 - __clz(orz));
    const int wq = dp_width_of_code(code);
    int off = wq;                                  // inclusive prefix sum of the widths
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, off, o);
        if (lane >= o) off += t;
    }
    const int pay = __shfl_sync(FULL, off, 31);
    off -= wq;
    int kind, words;
    const int hdr_bytes = 8 + 16 + (any_nan ? 32 : 0) + (any_nz ? 32 : 0);
    if (!has_val && !any_bad) { kind = 0; words = 1; }
    else if (any_bad) { kind = 2; words = DP_RAW_WORDS; }
    else if (pay == 0 && !any_nan && !any_nz) { kind = 3; words = 2; }
    else {
        kind = 1;
        words = (hdr_bytes + pay + 3) >> 2;
        if (words >= DP_RAW_WORDS) { kind = 2; words = DP_RAW_WORDS; }
    }
    if (!live) words = 0;
    if (lane == 0) rec_words[warp] = (uint32_t)words;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < DP_WARPS; ++w) tot += rec_words[w];
        blk_base = atomicAdd(&counters[0], (unsigned long long)tot);
    }
    __syncthreads();
    if (!live) return;
    unsigned long long at = blk_base;
    for (int w = 0; w < warp; ++w) at += rec_words[w];
    if (at + (unsigned long long)words > cap_words) {          // does not fit: flagged, not written
        if (lane == 0) {
            tile_off[g] = 0xFFFFFFFFu;
            atomicExch(&counters[1], 1ull);
        }
        return;
    }
    uint32_t* __restrict__ dst = payload + at;
    if (lane == 0) tile_off[g] = (uint32_t)at;
    const uint32_t mode = (uint32_t)kind | (any_nan ? 4u : 0u) | (any_nz ? 8u : 0u);
    if (kind == 0) {
        if (lane == 0) dst[0] = 0u;
        return;
    }
    if (kind == 3) {
        if (lane == 0) { dst[0] = 3u; dst[1] = (uint32_t)base; }
        return;
    }
    if (kind == 2) {
        if (lane == 0) dst[0] = 2u;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[1 + lane * 8 + j] = __float_as_uint(v[j]);
        return;
    }
    // kind 1: build the record in shared memory, copy it out with coalesced words
    uint32_t* st = stage[warp];
    uint8_t* sb = reinterpret_cast<uint8_t*>(st);
    if (lane == 0) { st[0] = mode; st[1] = (uint32_t)base; }
    {
        const int c_hi = __shfl_down_sync(FULL, code, 1);
        if ((lane & 1) == 0) sb[8 + (lane >> 1)] = (uint8_t)(code | (c_hi << 4));
    }
    int pos = 24;
    if (any_nan) { sb[pos + lane] = (uint8_t)m_nan; pos += 32; }
    if (any_nz) { sb[pos + lane] = (uint8_t)m_nz; pos += 32; }
    uint8_t* out = sb + pos + off;
    if (wq == 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            out[4 * j + 0] = (uint8_t)z[j];
            out[4 * j + 1] = (uint8_t)(z[j] >> 8);
            out[4 * j + 2] = (uint8_t)(z[j] >> 16);
            out[4 * j + 3] = (uint8_t)(z[j] >> 24);
        }
    } else if (wq > 0) {
        const int w = wq;                          // 1..16
        const uint32_t a0 = z[0] | (z[1] << w), a1 = z[2] | (z[3] << w);
        const uint32_t a2 = z[4] | (z[5] << w), a3 = z[6] | (z[7] << w);
        const uint64_t b0 = (uint64_t)a0 | ((uint64_t)a1 << (2 * w));
        const uint64_t b1 = (uint64_t)a2 | ((uint64_t)a3 << (2 * w));
        uint64_t lo, hi;
        if (w == 16) { lo = b0; hi = b1; }
        else { lo = b0 | (b1 << (4 * w)); hi = b1 >> (64 - 4 * w); }
        for (int i = 0; i < w && i < 8; ++i) { out[i] = (uint8_t)lo; lo >>= 8; }
        for (int i = 8; i < w; ++i) { out[i] = (uint8_t)hi; hi >>= 8; }
    }
    const int end = pos + pay;
    if (lane < ((4 - (end & 3)) & 3)) sb[end + lane] = 0;
    __syncwarp();
    for (int i = lane; i < words; i += 32) dst[i] = st[i];
}

// ------------------------------------------------------------------ host decode (dpack)
static inline float dp_value(int32_t f, float p) { return (float)f / p; }

// one tile; returns false on a malformed record
static bool dunpack_tile(const uint32_t* rec, const uint32_t* pay_end, int n, float p, float* out) {
    if (rec >= pay_end) return false;
    const uint32_t mode = rec[0];
    const int kind = (int)(mode & 3u);
    const float nanv = __builtin_nanf("");
    if (kind == 0) {
        for (int c = 0; c < n; ++c) out[c] = nanv;
        return true;
    }
    if (kind == 2) {
        if (rec + DP_RAW_WORDS > pay_end) return false;
        memcpy(out, rec + 1, sizeof(float) * (size_t)n);
        return true;
    }
    if (rec + 2 > pay_end) return false;
    int32_t f = (int32_t)rec[1];
    if (kind == 3) {
        const float x = dp_value(f, p);
        for (int c = 0; c < n; ++c) out[c] = x;
        return true;
    }
    const uint8_t* b = reinterpret_cast<const uint8_t*>(rec);
    const uint8_t* bend = reinterpret_cast<const uint8_t*>(pay_end);
    const uint8_t* nib = b + 8;
    const uint8_t* bm_nan = nullptr;
    const uint8_t* bm_nz = nullptr;
    const uint8_t* pay = b + 24;
    if (mode & 4u) { bm_nan = pay; pay += 32; }
    if (mode & 8u) { bm_nz = pay; pay += 32; }
    if (pay > bend) return false;
    const int n_grp = (n + 7) >> 3;
    for (int l = 0; l < n_grp; ++l) {
        const int code = (nib[l >> 1] >> ((l & 1) * 4)) & 15;
        const int w = dp_width_of_code(code);
        if (pay + w > bend) return false;
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
        const uint32_t mn = bm_nan ? bm_nan[l] : 0u, mz = bm_nz ? bm_nz[l] : 0u;
        const int cnt = (n - l * 8) < 8 ? (n - l * 8) : 8;
        float* o = out + l * 8;
        for (int j = 0; j < cnt; ++j) {
            const int32_t d = (int32_t)(z[j] >> 1) ^ -(int32_t)(z[j] & 1u);
            f = (int32_t)((uint32_t)f + (uint32_t)d);
            o[j] = ((mn >> j) & 1u) ? nanv : (((mz >> j) & 1u) ? -0.0f : dp_value(f, p));
        }
        // cells past the end of the row carry zero deltas: nothing to add
    }
    return true;
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

int64_t spx_dpack_tiles(int64_t row_len) {
    return row_len <= 0 ? 0 : (row_len + DP_TILE - 1) / DP_TILE;
}

int64_t spx_dpack_capacity(int64_t n_rows, int64_t row_len) {
    if (n_rows <= 0 || row_len <= 0) return 0;
    return n_rows * spx_dpack_tiles(row_len) * (int64_t)DP_RAW_WORDS * 4;
}

int spx_dpack_field_dev(const float* fld, int64_t n_rows, int64_t row_len, int64_t ld,
                        int32_t decimals, uint32_t* tile_off, void* payload,
                        int64_t capacity_bytes, uint64_t* counters, void* stream) {
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
    if (!fld || !tile_off || !payload || ld < row_len || decimals < 0 || decimals > 9 ||
        capacity_bytes < 0 || (reinterpret_cast<uintptr_t>(payload) & 3)) {
        set_error("dpack_field: bad argument (decimals must be 0..9, payload 4-byte aligned)");
        return SPX_EINVAL;
    }
    unsigned long long cap_words = (unsigned long long)capacity_bytes / 4;
    if (cap_words > 0xFFFFFFFEull) cap_words = 0xFFFFFFFEull;      // offsets are 32-bit words
    const int64_t tiles = spx_dpack_tiles(row_len);
    const int64_t n_tiles = n_rows * tiles;
    const int64_t n_blk = (n_tiles + DP_WARPS - 1) / DP_WARPS;
    if (n_blk > 0x7FFFFFFFll) {
        set_error("dpack_field: field too large for one call");
        return SPX_EINVAL;
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(fld) & 15) == 0) && (ld % 4 == 0);
    k_dpack<<<(unsigned)n_blk, DP_WARPS * 32, 0, st>>>(
        fld, row_len, ld, tiles, n_tiles, pack_pow10(decimals), vec_ok, tile_off,
        reinterpret_cast<uint32_t*>(payload), cap_words,
        reinterpret_cast<unsigned long long*>(counters));
    SPX_CHECK_LAUNCH("k_dpack");
    return SPX_OK;
}

int spx_dunpack_rows_host(const uint32_t* tile_off, const void* payload, int64_t payload_bytes,
                          int64_t n_rows, int64_t row_len, int32_t decimals, float* out,
                          int64_t out_ld, int32_t n_threads) {
    if (n_rows == 0 || row_len == 0) return SPX_OK;
    if (!tile_off || !payload || !out || out_ld < row_len || decimals < 0 || decimals > 9 ||
        payload_bytes < 0) {
        set_error("dunpack_rows: bad argument");
        return SPX_EINVAL;
    }
    const float p = pack_pow10(decimals);
    const int64_t tiles = spx_dpack_tiles(row_len);
    const uint32_t* pay = reinterpret_cast<const uint32_t*>(payload);
    const uint32_t* pay_end = pay + payload_bytes / 4;
    int bad = 0;
    auto work = [&](int part, int n_parts) {
        for (int64_t r = part; r < n_rows; r += n_parts) {
            const uint32_t* offs = tile_off + r * tiles;
            float* o = out + r * out_ld;
            for (int64_t t = 0; t < tiles; ++t) {
                const int64_t rest = row_len - t * DP_TILE;
                const int n = rest < DP_TILE ? (int)rest : DP_TILE;
                if (offs[t] == 0xFFFFFFFFu ||
                    !dunpack_tile(pay + offs[t], pay_end, n, p, o + t * DP_TILE))
                    __atomic_store_n(&bad, 1, __ATOMIC_RELAXED);
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
