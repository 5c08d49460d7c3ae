#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
#define WAIT(bar) asm volatile("{\n\t.reg .pred p;\n\tW%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W%=;\n\t}" ::"r"(smem_u32(bar)), "r"(0u) : "memory")
__global__ void k_bulk(const uint8_t *src, uint8_t *out) {
    __shared__ __align__(128) uint8_t tile[1024];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)), "l"(src), "r"(1024u), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(&bar);
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = tile[i];
}
template <int RANK>
__global__ void k_t(const CUtensorMap *map, int x, int y, int z, uint8_t *out, int bytes) {
    __shared__ __align__(128) uint8_t tile[4096];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tile)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(tile)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    WAIT(&bar);
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = tile[i];
}
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    const int cols = 480, rows = 640, B = 2;
    std::vector<uint8_t> h((size_t)cols * rows * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i / cols) * 3);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 4096);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    alignas(64) CUtensorMap m;
    CUtensorMap *dm; cudaMalloc(&dm, sizeof(m));
    cudaError_t e;
    if (which == 0) { k_bulk<<<1, 32>>>(d, o); e = cudaDeviceSynchronize(); printf("bulk 1D: %s\n", cudaGetErrorString(e)); return 0; }
    if (which == 1) {            // 2D u8, box 32 x 32
        const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows * B}; const cuuint64_t strides[1] = {(cuuint64_t)cols};
        const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = ((encode_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
        k_t<2><<<1, 32>>>(dm, 96, 77, 0, o, 1024); e = cudaDeviceSynchronize(); printf("2D u8 x=96 enc %d: %s\n", (int)r, cudaGetErrorString(e)); return 0;
    }
    if (which == 2) {            // 2D u8, unaligned x
        const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows * B}; const cuuint64_t strides[1] = {(cuuint64_t)cols};
        const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = ((encode_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
        k_t<2><<<1, 32>>>(dm, 101, 77, 0, o, 1024); e = cudaDeviceSynchronize(); printf("2D u8 x=101 enc %d: %s\n", (int)r, cudaGetErrorString(e)); return 0;
    }
    if (which == 3) {            // 2D f32 view of the same memory: 120 floats per row, box 8 floats x 32 rows
        const cuuint64_t dims[2] = {(cuuint64_t)cols / 4, (cuuint64_t)rows * B}; const cuuint64_t strides[1] = {(cuuint64_t)cols};
        const cuuint32_t box[2] = {8, 32}, es[2] = {1, 1};
        CUresult r = ((encode_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
        k_t<2><<<1, 32>>>(dm, 24, 77, 0, o, 1024); e = cudaDeviceSynchronize(); printf("2D f32 enc %d: %s\n", (int)r, cudaGetErrorString(e)); return 0;
    }
    return 0;
}
