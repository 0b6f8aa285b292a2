// Grid preparation on the GPU: which cells / stations lie inside or
// within a buffer distance of the selection polygons (reference: misc.py:407-540,
// chk_pt_cntmnt_in_polys_mp -> OGR Contains on buffered polygons, a Python loop over
// every cell) and the sampling of drift rasters at cells and stations
// (interp/drift.py:165-226).
//
// Arithmetic is spelled out with round-to-nearest intrinsics (no FMA contraction) so
// that the NumPy statement of the same formulas in the oracle gives identical bits.
#include <cstdint>

#include "spx_b200.h"
#include "spx_common.cuh"

namespace spx {

constexpr int PIP_CHUNK = 256;   // edges staged per pass

// One thread per point.  Even-odd crossing test per ring (ring ids are non-decreasing
// along the edge list) and, with buf2 > 0, squared distance to every edge.  Edge chunks
// whose y-range cannot touch the block's points are skipped.
__global__ void __launch_bounds__(256) k_points_in_polygons(
    const double* __restrict__ px, const double* __restrict__ py, int64_t n_pts,
    const double* __restrict__ ex1, const double* __restrict__ ey1,
    const double* __restrict__ ex2, const double* __restrict__ ey2,
    const int32_t* __restrict__ ering, int64_t n_edges, const double* __restrict__ cymin,
    const double* __restrict__ cymax, double buf, uint8_t* __restrict__ inside) {
    __shared__ double sx1[PIP_CHUNK], sy1[PIP_CHUNK], sx2[PIP_CHUNK], sy2[PIP_CHUNK];
    __shared__ int sring[PIP_CHUNK];
    __shared__ double s_lo[8], s_hi[8];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + tid;
    const bool live = i < n_pts;
    const double x = live ? px[i] : 0.0, y = live ? py[i] : 0.0;
    // y-range of the block's points
    double lo = live ? y : CUDART_INF, hi = live ? y : -CUDART_INF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    lo -= buf; hi += buf;
    const double buf2 = __dmul_rn(buf, buf);
    int cur_ring = -1;
    bool parity = false, in = false;
    const int64_t n_chunks = (n_edges + PIP_CHUNK - 1) / PIP_CHUNK;
    for (int64_t c = 0; c < n_chunks; ++c) {
        if (cymin != nullptr && (cymax[c] < lo || cymin[c] > hi)) continue;   // block-uniform
        __syncthreads();
        const int64_t e0 = c * PIP_CHUNK;
        const int ne = (int)min((int64_t)PIP_CHUNK, n_edges - e0);
        if (tid < ne) {
            sx1[tid] = ex1[e0 + tid]; sy1[tid] = ey1[e0 + tid];
            sx2[tid] = ex2[e0 + tid]; sy2[tid] = ey2[e0 + tid];
            sring[tid] = ering[e0 + tid];
        }
        __syncthreads();
        if (!live) continue;
        for (int e = 0; e < ne; ++e) {
            const double ax = sx1[e], ay = sy1[e], bx = sx2[e], by = sy2[e];
            const int rg = sring[e];
            if (rg != cur_ring) { in |= parity; parity = false; cur_ring = rg; }
            const double dx = __dsub_rn(bx, ax), dy = __dsub_rn(by, ay);
            if ((ay > y) != (by > y)) {
                // x of the edge at height y:  dx * (y - ay) / dy + ax
                const double xi = __dadd_rn(__ddiv_rn(__dmul_rn(dx, __dsub_rn(y, ay)), dy), ax);
                if (x < xi) parity = !parity;
            }
            if (buf > 0.0) {
                const double wx = __dsub_rn(x, ax), wy = __dsub_rn(y, ay);
                const double l2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                double t = (l2 > 0.0)
                    ? __ddiv_rn(__dadd_rn(__dmul_rn(wx, dx), __dmul_rn(wy, dy)), l2) : 0.0;
                t = fmin(fmax(t, 0.0), 1.0);
                const double qx = __dsub_rn(wx, __dmul_rn(t, dx));
                const double qy = __dsub_rn(wy, __dmul_rn(t, dy));
                const double d2 = __dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy));
                if (d2 < buf2) in = true;
            }
        }
    }
    if (live) inside[i] = (uint8_t)(in || parity);
}

// out[i] = ras[rows[i], cols[i]], NaN where np.isclose(ndv, value) (rtol 1e-5, atol 1e-8,
// interp/drift.py:196, :217) or the index is outside the raster.
__global__ void __launch_bounds__(256) k_sample_raster(const double* __restrict__ ras,
                                                       int64_t n_rows, int64_t n_cols,
                                                       const int64_t* __restrict__ rows,
                                                       const int64_t* __restrict__ cols,
                                                       int64_t n, double ndv, int has_ndv,
                                                       double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = rows[i], c = cols[i];
    double v = CUDART_NAN;
    if (r >= 0 && r < n_rows && c >= 0 && c < n_cols) {
        v = ras[r * n_cols + c];
        // np.isclose(a = ndv, b = v): |a - b| <= atol + rtol * |b|; inf == inf is close
        if (has_ndv) {
            const bool close = (ndv == v) ||
                (isfinite(ndv) && isfinite(v) &&
                 fabs(__dsub_rn(ndv, v)) <= __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(v))));
            if (close) v = CUDART_NAN;
        }
    }
    out[i] = v;
}

}  // namespace spx

using namespace spx;

extern "C" {

int spx_points_in_polygons_dev(const double* px, const double* py, int64_t n_pts,
                               const double* ex1, const double* ey1, const double* ex2,
                               const double* ey2, const int32_t* ering, int64_t n_edges,
                               const double* chunk_ymin, const double* chunk_ymax,
                               double buffer_dist, uint8_t* inside, void* stream) {
    if (n_pts == 0) return SPX_OK;
    if (!px || !py || !inside || n_edges < 0 || (n_edges > 0 && (!ex1 || !ey1 || !ex2 || !ey2 ||
                                                                 !ering)) ||
        !(buffer_dist >= 0.0) || ((chunk_ymin == nullptr) != (chunk_ymax == nullptr))) {
        set_error("points_in_polygons: bad argument");
        return SPX_EINVAL;
    }
    k_points_in_polygons<<<(unsigned)((n_pts + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        px, py, n_pts, ex1, ey1, ex2, ey2, ering, n_edges, chunk_ymin, chunk_ymax, buffer_dist,
        inside);
    SPX_CHECK_LAUNCH("k_points_in_polygons");
    return SPX_OK;
}

int spx_points_in_polygons_chunk(void) { return PIP_CHUNK; }

int spx_sample_raster_dev(const double* ras, int64_t n_rows, int64_t n_cols, const int64_t* rows,
                          const int64_t* cols, int64_t n, double ndv, int32_t has_ndv,
                          double* out, void* stream) {
    if (n == 0) return SPX_OK;
    if (!ras || !rows || !cols || !out || n_rows < 1 || n_cols < 1) {
        set_error("sample_raster: bad argument");
        return SPX_EINVAL;
    }
    k_sample_raster<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        ras, n_rows, n_cols, rows, cols, n, ndv, has_ndv, out);
    SPX_CHECK_LAUNCH("k_sample_raster");
    return SPX_OK;
}

}  // extern "C"
