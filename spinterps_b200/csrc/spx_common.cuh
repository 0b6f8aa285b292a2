// Shared device helpers for the spinterps B200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "spx_b200.h"

namespace spx {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SPX_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return spx::cuda_fail(e__, #call); \
    } while (0)

#define SPX_CHECK_LAUNCH(name)                                \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return spx::cuda_fail(e__, name); \
    } while (0)

// Variogram in kernel-parameter form (by value).
struct VgDev {
    int n_terms;
    int types[SPX_VG_MAX_TERMS];
    double sills[SPX_VG_MAX_TERMS];
    double ranges[SPX_VG_MAX_TERMS];
};

inline VgDev to_dev(const spx_vg& v) {
    VgDev d;
    d.n_terms = v.n_terms;
    for (int i = 0; i < SPX_VG_MAX_TERMS; ++i) {
        d.types[i] = v.types[i];
        d.sills[i] = v.sills[i];
        d.ranges[i] = v.ranges[i];
    }
    return d;
}

// IEEE distance without FMA contraction so that index decisions taken on it
// (nearest neighbour) reproduce NumPy's ((dx**2) + (dy**2)) ** 0.5 bit for bit
// (interp/grps.py:158-160, cyth/interpmthds.pyx:141).
__device__ __forceinline__ double dist_rn(double x1, double y1, double x2, double y2) {
    const double dx = __dsub_rn(x1, x2);
    const double dy = __dsub_rn(y1, y2);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// One nested-variogram term, cyth/interpmthds.pyx:38-83.  Same expression
// structure as the reference (quirks Q1, Q3 of SURVEY.md section 8a).
__device__ __forceinline__ double vg_term(int type, double h, double r, double s) {
    switch (type) {
        case SPX_VG_RNG:
            return h;
        case SPX_VG_NUG:
            return s;  // for every h, including 0
        case SPX_VG_SPH: {
            if (h >= r) return s;
            const double a = (1.5 * h) / r;
            const double b = (h * h * h) / (2 * (r * r * r));
            return s * (a - b);
        }
        case SPX_VG_EXP:
            return s * (1 - exp(-3 * h / r));
        case SPX_VG_LIN:
            return (h > r) ? s : s * (h / r);
        case SPX_VG_GAU:
            return s * (1 - exp(-3 * ((h * h) / (r * r))));
        case SPX_VG_POW:
            return s * pow(h, r);
        case SPX_VG_HOL: {
            if (h == 0) return 0.0;
            const double a = (CUDART_PI * h) / r;
            return s * (1 - (sin(a) / a));
        }
        default:
            return CUDART_NAN;
    }
}

// Sum over nested terms + covariance flip + min_vg_val cut,
// cyth/interpmthds.pyx:162-216.
__device__ __forceinline__ double vg_eval(const VgDev& vg, double h, int covar_flag,
                                          double min_vg_val) {
    double v = 0.0;
    if (covar_flag) {
        for (int t = 0; t < vg.n_terms; ++t)
            v += vg.sills[t] - vg_term(vg.types[t], h, vg.ranges[t], vg.sills[t]);
    } else {
        for (int t = 0; t < vg.n_terms; ++t)
            v += vg_term(vg.types[t], h, vg.ranges[t], vg.sills[t]);
    }
    if (v <= min_vg_val) v = 0.0;
    return v;
}

// Per-term constants precomputed on the host so that the B-tile generator needs
// no FP64 division: Sph s*(h*ca - h^3*cb), Exp/Gau s*(1 - exp(ca * h or h^2)),
// Lin s*h*ca.  Same functions as cyth/interpmthds.pyx:46-70, operations
// re-associated: (1.5*h)/r becomes h*(1.5/r) etc., a difference of a few ulp
// (tests compare the drop-in fill at 1e-13, the estimates at 1e-9).  Pow / Hol /
// Rng and the symmetric (diag) fill take the exact-expression path vg_eval().
struct VgFast {
    int n_terms;
    int all_fast;  // every term is one of Nug / Sph / Exp / Lin / Gau
    int types[SPX_VG_MAX_TERMS];
    double sills[SPX_VG_MAX_TERMS];
    double ranges[SPX_VG_MAX_TERMS];
    double ca[SPX_VG_MAX_TERMS];
    double cb[SPX_VG_MAX_TERMS];
};

__device__ __forceinline__ double vg_eval_fast(const VgFast& v, double h, int covar_flag,
                                               double min_vg_val) {
    double acc = 0.0;
#pragma unroll 1
    for (int t = 0; t < v.n_terms; ++t) {
        const int ty = v.types[t];
        const double s = v.sills[t];
        double g;
        if (ty == SPX_VG_NUG) {
            g = s;
        } else if (ty == SPX_VG_SPH) {
            const double h2 = h * h;
            g = (h >= v.ranges[t]) ? s : s * (h * v.ca[t] - h2 * h * v.cb[t]);
        } else if (ty == SPX_VG_EXP) {
            g = s * (1.0 - exp(v.ca[t] * h));
        } else if (ty == SPX_VG_GAU) {
            g = s * (1.0 - exp(v.ca[t] * (h * h)));
        } else {  // SPX_VG_LIN
            g = (h > v.ranges[t]) ? s : s * (h * v.ca[t]);
        }
        acc += covar_flag ? (s - g) : g;
    }
    if (acc <= min_vg_val) acc = 0.0;
    return acc;
}

inline VgFast make_vg_fast(int n_terms, const int* types, const double* sills,
                             const double* ranges) {
    VgFast f;
    f.n_terms = n_terms;
    f.all_fast = 1;
    for (int t = 0; t < SPX_VG_MAX_TERMS; ++t) {
        const int ty = (t < n_terms) ? types[t] : SPX_VG_NUG;
        const double r = (t < n_terms) ? ranges[t] : 1.0, sl = (t < n_terms) ? sills[t] : 0.0;
        f.types[t] = ty;
        f.sills[t] = sl;
        f.ranges[t] = r;
        f.ca[t] = f.cb[t] = 0.0;
        if (t >= n_terms) continue;
        switch (ty) {
            case SPX_VG_NUG: break;
            case SPX_VG_SPH: f.ca[t] = 1.5 / r; f.cb[t] = 1.0 / (2 * (r * r * r)); break;
            case SPX_VG_EXP: f.ca[t] = -3.0 / r; break;
            case SPX_VG_GAU: f.ca[t] = -3.0 / (r * r); break;
            case SPX_VG_LIN: f.ca[t] = 1.0 / r; break;
            default: f.all_fast = 0;
        }
    }
    return f;
}

__device__ __forceinline__ VgFast make_vg_fast_dev(const spx_vg& v) {
    VgFast f;
    f.n_terms = v.n_terms;
    f.all_fast = 1;
    for (int t = 0; t < SPX_VG_MAX_TERMS; ++t) {
        const int ty = (t < v.n_terms) ? v.types[t] : SPX_VG_NUG;
        const double r = (t < v.n_terms) ? v.ranges[t] : 1.0;
        f.types[t] = ty;
        f.sills[t] = (t < v.n_terms) ? v.sills[t] : 0.0;
        f.ranges[t] = r;
        f.ca[t] = f.cb[t] = 0.0;
        if (t >= v.n_terms) continue;
        if (ty == SPX_VG_SPH) { f.ca[t] = 1.5 / r; f.cb[t] = 1.0 / (2 * (r * r * r)); }
        else if (ty == SPX_VG_EXP) f.ca[t] = -3.0 / r;
        else if (ty == SPX_VG_GAU) f.ca[t] = -3.0 / (r * r);
        else if (ty == SPX_VG_LIN) f.ca[t] = 1.0 / r;
        else if (ty != SPX_VG_NUG) f.all_fast = 0;
    }
    return f;
}

// Angular sector of a reference point seen from a destination, cyth/interpmthds.pyx:849-866:
// atan of the slope plus quadrant fix-ups (xd == 0 -> 0; xd < 0, yd == 0 -> atan(-0.0) ->
// sector 0).  CUDA's atan may differ from libm's in the last bit, which can only matter
// for a point exactly on a sector edge.  The reference indexes out of range when the
// product rounds up to n_pies; clamped here.
__device__ __forceinline__ int pie_sector(double xd, double yd, int n_pies) {
    const double two_pi = 2.0 * CUDART_PI;
    double ang;
    if (xd == 0.0) {
        ang = 0.0;
    } else {
        ang = atan(__ddiv_rn(yd, xd));
        if (xd < 0.0 && yd > 0.0) ang = __dadd_rn(CUDART_PI, ang);
        else if (xd < 0.0 && yd < 0.0) ang = __dadd_rn(CUDART_PI, ang);
        else if (xd > 0.0 && yd < 0.0) ang = __dadd_rn(two_pi, ang);
    }
    const int p = (int)__ddiv_rn(__dmul_rn(ang, (double)n_pies), two_pi);
    return min(max(p, 0), n_pies - 1);
}

__device__ __forceinline__ double clampd(double v, int has_lo, int has_hi, double lo, double hi) {
    // NaN-safe like interp/steps.py:466-476 (comparisons with NaN are false)
    if (has_lo && v < lo) v = lo;
    if (has_hi && v > hi) v = hi;
    return v;
}

__device__ __forceinline__ void store_out(void* out, int64_t idx, double v, int out_f64) {
    if (out_f64)
        reinterpret_cast<double*>(out)[idx] = v;
    else
        reinterpret_cast<float*>(out)[idx] = static_cast<float>(v);
}

}  // namespace spx
