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

// inside vio_*_create after the handle exists: a failing CUDA call frees the handle before returning
#define VIO_CUDA_TRY_OR(expr, cleanup)                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            fprintf(stderr, "[vio_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                               \
            cleanup;                                                                           \
            return VIO_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the kernel FUNCTION, not of a handle: raise it once to everything the
// device allows (opt-in maximum minus the kernel's static shared memory) so that handles with different requirements can coexist;
// each launch is then validated against *max_dynamic by its caller.
template <typename K>
inline cudaError_t vio_allow_max_dynamic_smem(K kernel, int device, size_t *max_dynamic) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    const size_t dyn = (size_t)optin > fa.sharedSizeBytes ? (size_t)optin - fa.sharedSizeBytes : 0;
    if (max_dynamic) *max_dynamic = dyn;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
}

// Debug build only (-DVIO_DEBUG_POISON, libvio_b200_dbg.so): fill the CTA's whole shared-memory window (static + dynamic) with a
// NaN / -1 pattern at kernel entry so that a read of uninitialised shared memory shows up deterministically.  kernel_bit selects
// the kernel in the VIO_POISON_MASK environment variable read at library load (tools/poison_check.py).
#ifdef VIO_DEBUG_POISON
static __device__ unsigned g_vio_poison_mask = 0;
#include <stdlib.h>
static inline void vio_poison_load_mask() { const char *e = getenv("VIO_POISON_MASK"); const unsigned m = e ? (unsigned)strtoul(e, nullptr, 0) : 0u; cudaMemcpyToSymbol(g_vio_poison_mask, &m, sizeof(m)); }
__device__ __forceinline__ void vio_poison_smem(unsigned kernel_bit) {
    if (!(g_vio_poison_mask & kernel_bit)) return;
    unsigned n;
    asm volatile("mov.u32 %0, %%total_smem_size;" : "=r"(n));
    const unsigned nt = blockDim.x * blockDim.y * blockDim.z, t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    for (unsigned i = t * 4; i + 4 <= n; i += nt * 4) asm volatile("st.shared.u32 [%0], %1;" :: "r"(i), "r"(0xFFFFFFFFu) : "memory");
    __syncthreads();
}
#define VIO_POISON(bit) vio_poison_smem(bit)
#else
#define VIO_POISON(bit) do { } while (0)
static inline void vio_poison_load_mask() {}
#endif

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
// exact warp sum of 32-bit partials whose total needs more than 32 bits: low and high halves go through the hardware integer reduction
// (redux.sync) separately -- two instructions instead of ten shuffles and five 64-bit adds
__device__ __forceinline__ long long warp_sum_split(int v) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);      // <= 32 * 65535
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);                          // |.| <= 32 * 32768
    return (long long)hi * 65536 + (long long)lo;
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

// ---- device allocation.  Debug build (-DVIO_DEBUG_POISON): every array gets a 256-byte guard band on both sides filled with 0xA5;
// vio_debug_check_guards() (tools/guard_check.py) reports which array was written out of bounds, where and with what.
#include <vector>
struct VioAlloc { void *raw; void *user; size_t bytes; int idx; };
#ifdef VIO_DEBUG_POISON
constexpr size_t VIO_GUARD = 256;
inline std::vector<VioAlloc> &vio_guard_registry() { static std::vector<VioAlloc> r; return r; }
#else
constexpr size_t VIO_GUARD = 0;
#endif
inline cudaError_t vio_dev_alloc(void **user, size_t bytes, std::vector<void *> &owner) {
    bytes = bytes ? bytes : 1;
    unsigned char *raw = nullptr;
    cudaError_t e = cudaMalloc((void **)&raw, bytes + 2 * VIO_GUARD);
    if (e != cudaSuccess) return e;
    owner.push_back(raw);
#ifdef VIO_DEBUG_POISON
    cudaMemset(raw, 0xA5, bytes + 2 * VIO_GUARD);
    vio_guard_registry().push_back({raw, raw + VIO_GUARD, bytes, (int)vio_guard_registry().size()});
#endif
    *user = raw + VIO_GUARD;
    return cudaMemset(*user, 0, bytes);
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
