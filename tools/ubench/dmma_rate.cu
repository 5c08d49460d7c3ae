// Micro-benchmark: FP64 throughput of one B200 SM -- DFMA (CUDA cores) vs mma.sync m8n8k4 f64 (tensor pipe), 16 warps per CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_rate tools/ubench/dmma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k_dmma(double *out, int iters, long long *cyc) {
    double acc[ILP][2];
    for (int i = 0; i < ILP; i++) { acc[i][0] = threadIdx.x; acc[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma(acc[i][0], acc[i][1], a, b);
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma(double *out, int iters, long long *cyc) {
    double acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int threads : {32, 128, 512, 1024}) {
        k_dmma<8><<<148, threads>>>(out, iters, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c = (double)h[0];
        printf("DMMA m8n8k4 ILP8 %4d thr/SM: %.1f cycles/iter -> %.2f cycles per DMMA per warp-slot, %.1f FMA/clk/SM\n", threads, c / iters, c / iters / 8, 256.0 * 8 * (threads / 32) * iters / c);
        k_dmma<1><<<148, threads>>>(out, iters, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA dependent chain  %4d thr/SM: %.1f cycles latency\n", threads, (double)h[0] / iters);
        k_dfma<8><<<148, threads>>>(out, iters, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        c = (double)h[0];
        printf("DFMA ILP8             %4d thr/SM: %.1f FMA/clk/SM\n", threads, 8.0 * threads * iters / c);
        k_dfma<1><<<148, threads>>>(out, iters, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA dependent chain  %4d thr/SM: %.1f cycles latency\n", threads, (double)h[0] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
