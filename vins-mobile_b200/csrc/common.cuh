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

// BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba), repeated like cv::borderInterpolate when the overshoot exceeds the image
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while ((unsigned)i >= (unsigned)n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
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

// ---- optional per-kernel timing with CUDA events on the launching stream (bench.py roofline leg) -----------------
#include <vector>
#include <string>
#include <map>
struct KernelTimer {
    bool on = false;
    struct Rec { const char *name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    void begin(const char *name, cudaStream_t st) {
        if (!on) return;
        Rec r; r.name = name;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(cudaStream_t st) { if (on && !recs.empty()) cudaEventRecord(recs.back().b, st); }
    // drains: returns "name:count:total_ms;..." ; caller must have synchronised the stream
    std::string drain() {
        std::map<std::string, std::pair<int, double>> acc;
        for (auto &r : recs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            auto &e = acc[r.name]; e.first++; e.second += ms;
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        recs.clear();
        std::string out;
        char buf[160];
        for (auto &kv : acc) { snprintf(buf, sizeof(buf), "%s:%d:%.6f;", kv.first.c_str(), kv.second.first, kv.second.second); out += buf; }
        return out;
    }
};
#define VIO_LAUNCH(timer, stream, name, ...) do { (timer).begin(name, stream); __VA_ARGS__; (timer).end(stream); } while (0)
