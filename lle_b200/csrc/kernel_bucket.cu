// One translation unit per kernel bucket (compiled with -DLLE_AMAX=.. -DLLE_NBMAX=..) so that the
// buckets build in parallel.  Exposes a launcher and an occupancy query with C++ linkage.
#include <cuda_runtime.h>

#include "vec_kernels.cuh"

#ifndef LLE_AMAX
#error "compile with -DLLE_AMAX=<n> -DLLE_NBMAX=<n>"
#endif

#define LLE_CAT3(a, b, c) a##b##_##c
#define LLE_NAME(prefix, a, b) LLE_CAT3(prefix, a, b)

namespace lle {

cudaError_t LLE_NAME(launch_bucket_, LLE_AMAX, LLE_NBMAX)(const KParams& p, int grid, size_t smem, cudaStream_t stream) {
    lle_fused_kernel<LLE_AMAX, LLE_NBMAX><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t LLE_NAME(occupancy_bucket_, LLE_AMAX, LLE_NBMAX)(size_t smem, int* blocks) {
    auto kern = lle_fused_kernel<LLE_AMAX, LLE_NBMAX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, kThreads, smem);
}

}  // namespace lle
