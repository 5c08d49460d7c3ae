// common.cuh -- shared helpers for libvio_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vio_b200.h"

#define VIO_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            fprintf(stderr, "[vio_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                               \
            return VIO_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define VIO_MAXP 512          // compile-time cap on max_cnt (points per stream)
#define VIO_MAX_WIN 24        // compile-time cap on window_size

// BORDER_REFLECT_101 for |overshoot| < n
__device__ __forceinline__ int reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

// Explicitly rounded, never-contracted f32 ops: the parity contract with oracle/frontend_oracle.py
// names every f32 step, so the compiler must not fuse them.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
