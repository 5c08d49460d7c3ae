// Micro-benchmark: latency of the in-register 8x8 Cholesky + inverse of one warp (tc_diag_factor of be_tilechol.cuh), variants.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq3(double x) { double y = (double)rsqrtf((float)x); const double hx = 0.5 * x; y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); return y; }
__device__ __forceinline__ double rsq2(double x) { double y = (double)rsqrtf((float)x); const double hx = 0.5 * x; y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); return y; }
// variant A: as in be_tilechol.cuh (round 1 first version)
__device__ __forceinline__ bool fac_a(double &d0, double &d1, double &w0, double &w1, int npiv, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, t = lane & 3;
    double p0 = 0.0, p1 = 0.0; w0 = 0.0; w1 = 0.0; bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int jl = 4 * j + (j >> 1);
        const double piv = __shfl_sync(FULL, (j & 1) ? d1 : d0, jl);
        if (j < npiv) {
            ok &= (piv > 0.0) && (piv < 1e300);
            const double il = rsq3(piv);
            if (t == (j >> 1) && g >= j) { if (j & 1) d1 *= il; else d0 *= il; }
            const double lgj = __shfl_sync(FULL, (j & 1) ? d1 : d0, 4 * g + (j >> 1));
            const double lc0 = __shfl_sync(FULL, (j & 1) ? d1 : d0, 4 * (2 * t) + (j >> 1));
            const double lc1 = __shfl_sync(FULL, (j & 1) ? d1 : d0, 4 * (2 * t + 1) + (j >> 1));
            if (g > j) { if (2 * t > j && 2 * t <= g) d0 -= lgj * lc0; if (2 * t + 1 > j && 2 * t + 1 <= g) d1 -= lgj * lc1; }
            if (g == j) { w0 = (2 * t <= j) ? il * ((2 * t == j ? 1.0 : 0.0) - p0) : 0.0; w1 = (2 * t + 1 <= j) ? il * ((2 * t + 1 == j ? 1.0 : 0.0) - p1) : 0.0; }
            const double wj0 = __shfl_sync(FULL, w0, 4 * j + t), wj1 = __shfl_sync(FULL, w1, 4 * j + t);
            if (g > j) { p0 += lgj * wj0; p1 += lgj * wj1; }
        }
    }
    if (2 * t > g) d0 = 0.0; if (2 * t + 1 > g) d1 = 0.0; if (g >= npiv) { w0 = 0.0; w1 = 0.0; }
    return ok;
}
// variant B: branch-free (masks as multipliers), two Newton steps, the pivot of step j+1 is updated first
__device__ __forceinline__ bool fac_b(double &d0, double &d1, double &w0, double &w1, int npiv, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, t = lane & 3;
    double p0 = 0.0, p1 = 0.0; w0 = 0.0; w1 = 0.0; bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double piv = __shfl_sync(FULL, (j & 1) ? d1 : d0, 4 * j + (j >> 1));
        const bool act = j < npiv;
        ok &= !act || ((piv > 0.0) && (piv < 1e300));
        const double il = act ? rsq2(piv) : 0.0;
        const double cur = (j & 1) ? d1 : d0;
        const double scaled = (act && t == (j >> 1) && g >= j) ? cur * il : cur;
        if (j & 1) d1 = scaled; else d0 = scaled;
        const double lgj = __shfl_sync(FULL, scaled, 4 * g + (j >> 1));
        const double lc0 = __shfl_sync(FULL, scaled, 4 * (2 * t) + (j >> 1));
        const double lc1 = __shfl_sync(FULL, scaled, 4 * (2 * t + 1) + (j >> 1));
        const double m0 = (act && g > j && 2 * t > j && 2 * t <= g) ? lgj : 0.0, m1 = (act && g > j && 2 * t + 1 > j && 2 * t + 1 <= g) ? lgj : 0.0;
        d0 = fma(-m0, lc0, d0); d1 = fma(-m1, lc1, d1);
        const double nw0 = (2 * t <= j) ? il * ((2 * t == j ? 1.0 : 0.0) - p0) : 0.0, nw1 = (2 * t + 1 <= j) ? il * ((2 * t + 1 == j ? 1.0 : 0.0) - p1) : 0.0;
        w0 = (act && g == j) ? nw0 : w0; w1 = (act && g == j) ? nw1 : w1;
        const double wj0 = __shfl_sync(FULL, w0, 4 * j + t), wj1 = __shfl_sync(FULL, w1, 4 * j + t);
        const double mg = (act && g > j) ? lgj : 0.0;
        p0 = fma(mg, wj0, p0); p1 = fma(mg, wj1, p1);
    }
    if (2 * t > g) d0 = 0.0; if (2 * t + 1 > g) d1 = 0.0; if (g >= npiv) { w0 = 0.0; w1 = 0.0; }
    return ok;
}
// variant C: the pivot recurrence runs on 1/pivot (MUFU.RCP64H + two Newton steps), one shuffle stage per step; the column scaling by
// rsqrt(pivot) and the inverse are off the critical chain
__device__ __forceinline__ double rcp2(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r); return r; }
__device__ __forceinline__ double rsqa(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); const double hx = 0.5 * x; y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); return y; }
__device__ __forceinline__ bool fac_c(double &d0, double &d1, double &w0, double &w1, int npiv, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, t = lane & 3;
    double p0 = 0.0, p1 = 0.0; w0 = 0.0; w1 = 0.0; bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double cur = (j & 1) ? d1 : d0;
        const double piv = __shfl_sync(FULL, cur, 4 * j + (j >> 1));
        const double agj = __shfl_sync(FULL, cur, 4 * g + (j >> 1));
        const double ac0 = __shfl_sync(FULL, cur, 4 * (2 * t) + (j >> 1));
        const double ac1 = __shfl_sync(FULL, cur, 4 * (2 * t + 1) + (j >> 1));
        const bool act = j < npiv;
        ok &= !act || ((piv > 0.0) && (piv < 1e300));
        const double inv = act ? rcp2(piv) : 0.0;
        const double s = (g > j) ? agj * inv : 0.0;
        d0 = fma(-((2 * t > j && 2 * t <= g) ? s : 0.0), ac0, d0);
        d1 = fma(-((2 * t + 1 > j && 2 * t + 1 <= g) ? s : 0.0), ac1, d1);
        const double il = act ? rsqa(piv) : 0.0;
        const double lgj = agj * il;                                 // L[g][j] (rows g >= j)
        if (j & 1) d1 = (act && t == (j >> 1) && g >= j) ? lgj : d1; else d0 = (act && t == (j >> 1) && g >= j) ? lgj : d0;
        const double nw0 = (2 * t <= j) ? il * ((2 * t == j ? 1.0 : 0.0) - p0) : 0.0, nw1 = (2 * t + 1 <= j) ? il * ((2 * t + 1 == j ? 1.0 : 0.0) - p1) : 0.0;
        w0 = (act && g == j) ? nw0 : w0; w1 = (act && g == j) ? nw1 : w1;
        const double wj0 = __shfl_sync(FULL, w0, 4 * j + t), wj1 = __shfl_sync(FULL, w1, 4 * j + t);
        const double mg = (act && g > j) ? lgj : 0.0;
        p0 = fma(mg, wj0, p0); p1 = fma(mg, wj1, p1);
    }
    if (2 * t > g) d0 = 0.0; if (2 * t + 1 > g) d1 = 0.0; if (g >= npiv) { w0 = 0.0; w1 = 0.0; }
    return ok;
}
template <int V>
__global__ void k(double *out, long long *cyc, int reps, int nwarps_busy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double acc0 = 0, acc1 = 0; long long tot = 0;
    if (warp == 0) {
        for (int r = 0; r < reps; r++) {
            // SPD tile: 8 I + small symmetric part
            double d0 = (g == 2 * t ? 8.0 : 0.0) + 0.01 * (g + 2 * t + r % 3), d1 = (g == 2 * t + 1 ? 8.0 : 0.0) + 0.01 * (g + 2 * t + 1 + r % 3), w0, w1;
            const long long t0 = clock64();
            bool ok = V == 0 ? fac_a(d0, d1, w0, w1, 8, lane) : V == 1 ? fac_b(d0, d1, w0, w1, 8, lane) : fac_c(d0, d1, w0, w1, 8, lane);
            const long long t1 = clock64();
            tot += t1 - t0; acc0 += d0 + w0 + ok; acc1 += d1 + w1;
        }
        if (lane == 0) cyc[0] = tot / reps;
        out[lane] = acc0 + acc1;
    } else if (warp < 1 + nwarps_busy) {
        // background DMMA load on the other warps
        double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9, c0 = 0, c1 = 0, e0 = 0, e1 = 0;
        for (int r = 0; r < reps * 40; r++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
        }
        out[threadIdx.x] = c0 + c1 + e0 + e1;
    }
}
__global__ void check(double *o) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = (g == 2 * t ? 8.0 : 0.0) + 0.01 * (g + 2 * t) * (1 + (g ^ (2 * t))), a1 = (g == 2 * t + 1 ? 8.0 : 0.0) + 0.01 * (g + 2 * t + 1) * (1 + (g ^ (2 * t + 1)));
    // symmetric by construction: f(g,c) = 0.01 (g+c)(1+(g^c))
    double d0 = a0, d1 = a1, w0, w1, e0 = a0, e1 = a1, v0, v1;
    fac_a(d0, d1, w0, w1, 8, lane); fac_c(e0, e1, v0, v1, 8, lane);
    o[4 * lane] = fabs(d0 - e0); o[4 * lane + 1] = fabs(d1 - e1); o[4 * lane + 2] = fabs(w0 - v0); o[4 * lane + 3] = fabs(w1 - v1);
}
int main() {
    { double *o, h[128]; cudaMalloc(&o, 128 * 8); check<<<1, 32>>>(o); cudaMemcpy(h, o, 128 * 8, cudaMemcpyDeviceToHost); double m = 0; for (int i = 0; i < 128; i++) m = h[i] > m ? h[i] : m; printf("max |A - C| over L and W entries: %.3e\n", m); }
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    for (int busy : {0, 3, 12, 15}) {
        k<0><<<1, 512>>>(out, cyc, 200, busy); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("variant A, %2d busy DMMA warps: %lld cycles per 8x8 factor+inverse\n", busy, h);
        k<1><<<1, 512>>>(out, cyc, 200, busy); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("variant B, %2d busy DMMA warps: %lld cycles\n", busy, h);
        k<2><<<1, 512>>>(out, cyc, 200, busy); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("variant C, %2d busy DMMA warps: %lld cycles\n", busy, h);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
