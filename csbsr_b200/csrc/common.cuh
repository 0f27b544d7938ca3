// Shared helpers for the csbsr_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace csbsr {

void set_error(const char* fmt, ...);

#define CSBSR_CHECK_CUDA(expr)                                                        \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            csbsr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                             __FILE__, __LINE__);                                     \
            return -2;                                                                \
        }                                                                             \
    } while (0)

#define CSBSR_REQUIRE(cond, ...)                                                      \
    do {                                                                              \
        if (!(cond)) {                                                                \
            csbsr::set_error(__VA_ARGS__);                                            \
            return -1;                                                                \
        }                                                                             \
    } while (0)

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a FIXED order (shuffle tree inside each warp, then the warps in index order): bit-reproducible, unlike
// shared-memory atomics.  `scratch` holds one value per warp (<= 32); every thread returns the total.
template <typename T>
__device__ __forceinline__ T block_sum_det(T v, T* scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    T s = scratch[0];
    for (int i = 1; i < nw; ++i) s += scratch[i];
    return s;
}

}  // namespace csbsr
