// Write-bandwidth microbenchmark for the output pattern of the local estimator:
// a [rows, cells] f32 field (rows = time steps, 4 MB pitch), every block writes a tile of
// TR rows x TC cells.  Variants: 4 / 8 / 16 bytes per lane and store, tile shapes, and a
// shared-memory staged tile written with cp.async.bulk (TMA bulk store, 1 row segment per
// copy).  Compared with cudaMemsetAsync of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VEC, int TR>
__global__ void __launch_bounds__(256) k_store(float* out, int64_t ld, int rows, int64_t cells, float v) {
    // block: 256 threads x VEC cells, TR rows
    const int64_t c0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * VEC;
    const int r0 = blockIdx.y * TR;
    if (c0 >= cells) return;
#pragma unroll 4
    for (int r = r0; r < min(rows, r0 + TR); ++r) {
        float* p = out + (int64_t)r * ld + c0;
        const float x = v + r;
        if (VEC == 1) *p = x;
        else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(x, x);
        else *reinterpret_cast<float4*>(p) = make_float4(x, x, x, x);
    }
}

// staged: block fills a [TR x 1024] f32 tile in shared memory 16 rows at a time and one
// thread writes each row segment (4 KB) with a bulk async store
template <int TR>
__global__ void __launch_bounds__(256) k_store_bulk(float* out, int64_t ld, int rows, int64_t cells, float v) {
    extern __shared__ __align__(128) float tile[];   // [2][16][1024]
    const int64_t c0 = (int64_t)blockIdx.x * 1024;
    const int r0 = blockIdx.y * TR;
    const int nc = (int)min((int64_t)1024, cells - c0);
    int buf = 0;
    for (int rb = r0; rb < min(rows, r0 + TR); rb += 16, buf ^= 1) {
        float* t = tile + buf * 16 * 1024;
        // wait until the bulk stores that read this buffer two rounds ago are done
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < 16 * 256; i += 256) {
            const int rr = i >> 8, cc = (i & 255) * 4;
            *reinterpret_cast<float4*>(t + rr * 1024 + cc) = make_float4(v + rb + rr, v, v, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 16 && rb + threadIdx.x < rows) {
            float* g = out + (int64_t)(rb + threadIdx.x) * ld + c0;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(t + threadIdx.x * 1024);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(g), "r"(s), "r"(nc * 4) : "memory");
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// same, but the block's tile is WC cells wide (one bulk store of WC * 4 bytes per row) and
// 256 threads write one float each: the shape the local estimator's block has
template <int TR, int WC, int RB>
__global__ void __launch_bounds__(256) k_store_bulk_w(float* out, int64_t ld, int rows, int64_t cells, float v) {
    extern __shared__ __align__(128) float tile[];   // [2][RB][WC]
    const int64_t c0 = (int64_t)blockIdx.x * WC;
    const int r0 = blockIdx.y * TR;
    const int nc = (int)min((int64_t)WC, cells - c0);
    int buf = 0;
    for (int rb = r0; rb < min(rows, r0 + TR); rb += RB, buf ^= 1) {
        float* t = tile + buf * RB * WC;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < RB * WC; i += 256) t[i] = v + rb + (i / WC);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < RB && rb + threadIdx.x < rows) {
            float* g = out + (int64_t)(rb + threadIdx.x) * ld + c0;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(t + threadIdx.x * WC);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(g), "r"(s), "r"(nc * 4) : "memory");
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// persistent, row-synchronous: CPS blocks per SM; block b owns tiles b, b + P, ... and walks
// the rows RG at a time (rows outer, tiles inner): the stores in flight across the GPU stay
// within a few rows
template <int RG>
__global__ void __launch_bounds__(256) k_store_persist(float* out, int64_t ld, int rows, int64_t cells, float v) {
    const int n_tiles = (int)((cells + 255) / 256);
    for (int g = 0; g < rows; g += RG) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int64_t c = (int64_t)t * 256 + threadIdx.x;
            if (c >= cells) continue;
#pragma unroll
            for (int u = 0; u < RG; ++u)
                if (g + u < rows) __stcs(out + (int64_t)(g + u) * ld + c, v + g + u);
        }
    }
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int rows = 1250;
    const int64_t cells = 1000000, ld = cells;
    const size_t bytes = (size_t)rows * ld * 4;
    float* out; cudaMalloc(&out, bytes);
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    rep("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(out, 0, bytes); }));
#define RUN(VEC, TR) { dim3 g((unsigned)((cells + 256 * VEC - 1) / (256 * VEC)), (rows + TR - 1) / TR); \
        rep("k_store VEC=" #VEC " TR=" #TR, time_ms([&] { k_store<VEC, TR><<<g, 256>>>(out, ld, rows, cells, 1.f); })); }
    RUN(1, 128) RUN(1, 64) RUN(1, 32) RUN(1, 1250)
    RUN(2, 128) RUN(2, 64)
    RUN(4, 128) RUN(4, 64) RUN(4, 32) RUN(4, 16)
#define RUNB(TR) { dim3 g((unsigned)((cells + 1023) / 1024), (rows + TR - 1) / TR); \
        cudaFuncSetAttribute(k_store_bulk<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 1024 * 4); \
        rep("k_store_bulk (TMA bulk store) TR=" #TR, time_ms([&] { k_store_bulk<TR><<<g, 256, 2 * 16 * 1024 * 4>>>(out, ld, rows, cells, 1.f); })); }
    RUNB(128) RUNB(64) RUNB(32)
#define RUNW(TR, WC, RB) { dim3 g((unsigned)((cells + WC - 1) / WC), (rows + TR - 1) / TR); \
        cudaFuncSetAttribute(k_store_bulk_w<TR, WC, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RB * WC * 4); \
        rep("k_store_bulk_w TR=" #TR " WC=" #WC " RB=" #RB, time_ms([&] { k_store_bulk_w<TR, WC, RB><<<g, 256, 2 * RB * WC * 4>>>(out, ld, rows, cells, 1.f); })); }
    RUNW(128, 256, 8) RUNW(128, 256, 16) RUNW(128, 256, 32) RUNW(128, 512, 8) RUNW(128, 512, 16)
    RUNW(64, 256, 16) RUNW(256, 256, 16) RUNW(1250, 256, 16) RUNW(128, 1024, 8)
    RUN(1, 16) RUN(1, 8)
#define RUNP(RG, CPS) rep("k_store_persist RG=" #RG " blocks/SM=" #CPS, time_ms([&] { k_store_persist<RG><<<148 * CPS, 256>>>(out, ld, rows, cells, 1.f); }));
    RUNP(8, 3) RUNP(8, 5) RUNP(8, 8) RUNP(4, 8) RUNP(16, 4) RUNP(2, 8)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
