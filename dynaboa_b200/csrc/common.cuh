// Shared device helpers and launch plumbing for libdynaboa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DBOA_OK 0
#define DBOA_ERR_ARG (-1)
#define DBOA_ERR_SHAPE (-2)
#define DBOA_ERR_CUDA (-3)
#define DBOA_ERR_UNSUPPORTED (-4)

namespace dboa {

extern int g_last_cuda_error;      // set by check_launch (defined in hmr_plan.cu)
extern int g_launch_count;         // number of kernels this library launched (bench "gpu_launches")

inline int check_launch() {
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return DBOA_ERR_CUDA; }
    return DBOA_OK;
}

#define DBOA_TRY(expr) do { int _s = (expr); if (_s != DBOA_OK) return _s; } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree); result valid in every thread.  blockDim.x <= 1024.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect smem32 reuse across consecutive calls
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? smem32[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dboa
