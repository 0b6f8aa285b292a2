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

}  // extern "C"
