// Per-step field statistics (interp/main.py:474-525) shared by the output-stage kernels:
// Chan et al. partials (n, mean, M2, min, max, finite count) and np.round in the field dtype.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace spx {

struct StatPart {
    double n, mean, m2, mn, mx, nfin;
};

__device__ __forceinline__ void stat_merge(StatPart& a, const StatPart& b) {
    // Chan et al. pairwise update of (n, mean, M2)
    if (b.n > 0.0) {
        if (a.n == 0.0) {
            a.n = b.n; a.mean = b.mean; a.m2 = b.m2;
        } else {
            const double n = a.n + b.n;
            const double dlt = b.mean - a.mean;
            a.mean += dlt * (b.n / n);
            a.m2 += b.m2 + dlt * dlt * (a.n * b.n / n);
            a.n = n;
        }
    }
    a.mn = fmin(a.mn, b.mn);     // fmin / fmax ignore NaN: nanmin / nanmax
    a.mx = fmax(a.mx, b.mx);
    a.nfin += b.nfin;
}

__device__ __forceinline__ StatPart stat_shfl(const StatPart& a, int o) {
    StatPart b;
    b.n = __shfl_xor_sync(0xffffffffu, a.n, o);
    b.mean = __shfl_xor_sync(0xffffffffu, a.mean, o);
    b.m2 = __shfl_xor_sync(0xffffffffu, a.m2, o);
    b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
    b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
    b.nfin = __shfl_xor_sync(0xffffffffu, a.nfin, o);
    return b;
}

template <typename T>
__device__ __forceinline__ T round_dec(T x, T p);
template <>
__device__ __forceinline__ float round_dec<float>(float x, float p) {
    return __fdiv_rn(rintf(__fmul_rn(x, p)), p);
}
template <>
__device__ __forceinline__ double round_dec<double>(double x, double p) {
    return __ddiv_rn(rint(__dmul_rn(x, p)), p);
}

// reduces parts[row * n_seg + s] over s into stats[0..4][row] (spx_misc.cu)
void launch_stats_final(const StatPart* parts, int n_seg, int64_t n_rows, double* stats,
                        cudaStream_t st);

}  // namespace spx
