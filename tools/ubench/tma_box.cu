// Minimal check of the TMA box load lk_kernel uses: rank-3 u8 tensor (x, y, stream), 32 x 32 x 1 box, mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_box tma_box.cu && ./tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const CUtensorMap *map, int x, int y, int z, uint8_t *out) {
    __shared__ __align__(128) uint8_t tile[1024];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024u) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = tile[i];
}
int main() {
    const int cols = 480, rows = 640, B = 2;
    std::vector<uint8_t> h((size_t)cols * rows * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i / cols) * 3);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 1024);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, p);
    alignas(64) CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)cols, (cuuint64_t)rows * cols};
    const cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
    CUresult r = ((encode_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CUtensorMap *dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    const int x = 101, y = 77, z = 1;
    k<<<1, 32>>>(dm, x, y, z, o);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint8_t> got(1024); cudaMemcpy(got.data(), o, 1024, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < 32; r2++) for (int c = 0; c < 32; c++) bad += got[r2 * 32 + c] != h[(size_t)z * rows * cols + (size_t)(y + r2) * cols + x + c];
    printf("mismatches: %d\n", bad);
    return 0;
}
