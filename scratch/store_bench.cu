// Microbenchmark: how fast can one B200 stream observation-sized tiles from shared memory to HBM?
//   mode 0: cp.async.bulk.global.shared::cta (TMA 1-D bulk store), 2 buffers per warp
//   mode 1: LDS.128 + STG.128 by the warp's lanes
//   mode 2: STG.128 from registers (no smem read), upper bound of the LSU path
//   mode 3: TMA with NBUF buffers per warp (deeper queue)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int MODE, int NBUF>
__global__ void k(float* out, long n_tiles, int tile_floats, long stride_floats, unsigned* ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* tiles = reinterpret_cast<float*>(smem) + (size_t)warp * NBUF * tile_floats;
    for (int i = lane; i < NBUF * tile_floats; i += 32) tiles[i] = (float)(i & 1);
    __syncwarp();
    int buf = 0;
    const long gw = (long)blockIdx.x * nw + warp, tw = (long)gridDim.x * nw;
    for (long t = gw; t < n_tiles; t += tw) {
        float* tile = tiles + (size_t)buf * tile_floats;
        float* dst = out + t * stride_floats;
        if (MODE == 0 || MODE == 3) {
            if (lane == 0) bulk_wait_read<NBUF - 1>();
            __syncwarp();
            tile[lane] = (float)t;  // a token patch
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_store(dst, tile, tile_floats * 4); bulk_commit(); }
        } else if (MODE == 1) {
            tile[lane] = (float)t;
            __syncwarp();
            const float4* s4 = reinterpret_cast<const float4*>(tile);
            float4* d4 = reinterpret_cast<float4*>(dst);
            #pragma unroll 4
            for (int i = lane; i < tile_floats / 4; i += 32) d4[i] = s4[i];
            __syncwarp();
        } else {
            float4 v = make_float4((float)t, 0.f, 1.f, 0.f);
            float4* d4 = reinterpret_cast<float4*>(dst);
            #pragma unroll 4
            for (int i = lane; i < tile_floats / 4; i += 32) d4[i] = v;
        }
        buf = (buf + 1) % NBUF;
    }
    if ((MODE == 0 || MODE == 3) && lane == 0) bulk_wait_all();
}

template <int MODE, int NBUF>
float run(float* out, long n_tiles, int tile_floats, long stride_floats, int warps, int reps) {
    size_t smem = (size_t)warps * NBUF * tile_floats * 4;
    auto kern = k<MODE, NBUF>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1.f;
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, warps * 32, smem);
    if (bps < 1) return -1.f;
    int grid = 148 * bps;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, warps * 32, smem>>>(out, n_tiles, tile_floats, stride_floats, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<grid, warps * 32, smem>>>(out, n_tiles, tile_floats, stride_floats, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (cudaGetLastError() != cudaSuccess) return -2.f;
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double gbs = (double)n_tiles * tile_floats * 4 * reps / (ms / 1e3) / 1e9;
    printf("mode %d nbuf %d tile %6d B stride %6ld B warps/cta %2d cta/sm %d (warps/sm %2d): %8.1f GB/s  (%.1f us/launch)\n", MODE, NBUF,
           tile_floats * 4, stride_floats * 4, warps, bps, warps * bps, gbs, ms * 1e3 / reps);
    return (float)gbs;
}

int main() {
    const long total_bytes = 491L << 20;  // ~ one step of 65,536 lvl6 observations
    float* out;
    cudaMalloc(&out, total_bytes + (1 << 20));
    for (int tf : {1872, 1920, 468, 3744, 7488}) {       // 7488 B (lvl6), 7680 B (128-aligned), 1872 B, 14976 B, 29952 B
        long n_tiles = total_bytes / (tf * 4);
        for (int warps : {4, 8}) {
            run<0, 2>(out, n_tiles, tf, tf, warps, 20);
            run<3, 4>(out, n_tiles, tf, tf, warps, 20);
            run<1, 1>(out, n_tiles, tf, tf, warps, 20);
            run<2, 1>(out, n_tiles, tf, tf, warps, 20);
        }
    }
    // reference: cudaMemset
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(out, 0, total_bytes);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) cudaMemsetAsync(out, i, total_bytes);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset: %.1f GB/s\n", (double)total_bytes * 20 / (ms / 1e3) / 1e9);
    return 0;
}
