// fe_kernels.cuh -- front-end CUDA kernels (sm_100a).  Every kernel is batched over streams (blockIdx.z or
// blockIdx.y = stream).  The arithmetic contract is oracle/frontend_oracle.py (r_* functions), which is in turn
// pinned to cv2 4.13; reference call sites are /root/reference/VINS_ios/feature_tracker.cpp:95,181,198,263.
#pragma once
#include "common.cuh"
#include <cuda.h>                      // CUtensorMap (the descriptors are encoded on the host through the driver entry point)

namespace fe {

constexpr int LK_WIN = 21;
constexpr int LK_LEVELS = 3;          // maxLevel (4 pyramid levels)
constexpr int LK_MAX_ITERS = 30;
constexpr int W_BITS = 14;
constexpr int CAND_CAP = 65536;       // raw local maxima per stream
constexpr int SORT_CAP = 8192;        // candidates above the quality threshold per stream

struct PyrLevels {
    const uint8_t *p[4];              // level base pointers for stream 0
    int rows[4], cols[4];
    size_t stride[4];                 // bytes between consecutive streams at this level
};

// TMA descriptors of the pyramid levels, rank 3 (x, y, stream), u8, box LK_BOX x 32 x 1 (48 x 32 x 1), no swizzle, out-of-bounds elements read as 0.
// The box must START on a 16-byte boundary in global memory (measured on B200: an unaligned x coordinate raises "illegal instruction",
// tools/ubench/tma_box2.cu), so a window is fetched as the aligned 48-column box that contains its 33 columns.  A level gets a
// descriptor only when its row stride is a multiple of 16 bytes (a tensor-map requirement: 480 and 240 at 640x480; every level at
// 1280x720); ok[l] = 0 keeps the level on the register-staged path.
struct LkMaps {
    const CUtensorMap *I, *J;         // [4] each, in global memory (64-byte aligned)
    int ok[4];
};

// ---- TMA / mbarrier primitives (PTX ISA: cp.async.bulk.tensor, mbarrier) ---------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, void *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// K1  pyrDown: 5-tap [1 4 6 4 1] x [1 4 6 4 1], REFLECT_101, (s + 128) >> 8, output ((H+1)/2,(W+1)/2)
//     (cv::pyrDown inside calcOpticalFlowPyrLK; oracle r_pyr_down).  One thread -> 4 horizontally adjacent
//     outputs (packed u8x4 store); the 5x11 input footprint is read as bytes through L1.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_down_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                       int ir, int ic, int orows, int ocols, size_t in_stride,
                                                       size_t out_stride) {
    const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ox0 >= ocols || oy >= orows) return;
    const uint8_t *src = in + (size_t)blockIdx.z * in_stride;
    uint8_t *dst = out + (size_t)blockIdx.z * out_stride;
    int acc[4] = {0, 0, 0, 0};
    const int kw[5] = {1, 4, 6, 4, 1};
    const bool interior = (2 * ox0 - 4 >= 0) && (2 * ox0 + 12 <= ic);
    const bool aligned = interior && ((ic & 7) == 0) && ((in_stride & 7) == 0);
#pragma unroll
    for (int dy = -2; dy <= 2; dy++) {
        const int y = reflect101(2 * oy + dy, ir);
        const uint8_t *row = src + (size_t)y * ic;
        int v[11];
        if (aligned) {
            // the 11-byte footprint [2ox0-2, 2ox0+8] sits inside 16 bytes starting at 2ox0-4: u32 + u64 + u32 (all naturally aligned).
            // Horizontal taps (1 4 6 4 | 1) times the vertical weight as two byte dot products (dp4a) per output: no byte extraction.
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(row + 2 * ox0 - 4);
            const uint2 w12 = *reinterpret_cast<const uint2 *>(row + 2 * ox0);
            const uint32_t w3 = *reinterpret_cast<const uint32_t *>(row + 2 * ox0 + 8);
            const unsigned kv = (unsigned)kw[dy + 2];
            const unsigned k4 = kv * 0x04060401u;                       // bytes (1,4,6,4) * kv  (<= 36)
            const unsigned kb0 = kv, kb2 = kv << 16;                    // the fifth tap sits in byte 0 / byte 2 of the next word
            const unsigned f0 = __funnelshift_r(w0, w12.x, 16), f1 = __funnelshift_r(w12.x, w12.y, 16);
            acc[0] = (int)__dp4a(f0, k4, __dp4a(w12.x, kb2, (unsigned)acc[0]));
            acc[1] = (int)__dp4a(w12.x, k4, __dp4a(w12.y, kb0, (unsigned)acc[1]));
            acc[2] = (int)__dp4a(f1, k4, __dp4a(w12.y, kb2, (unsigned)acc[2]));
            acc[3] = (int)__dp4a(w12.y, k4, __dp4a(w3, kb0, (unsigned)acc[3]));
            continue;
        } else if (2 * ox0 - 2 >= 0 && 2 * ox0 + 8 < ic) {
#pragma unroll
            for (int k = 0; k < 11; k++) v[k] = row[2 * ox0 - 2 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 11; k++) v[k] = row[reflect101(2 * ox0 - 2 + k, ic)];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int h = v[2 * j] + 4 * v[2 * j + 1] + 6 * v[2 * j + 2] + 4 * v[2 * j + 3] + v[2 * j + 4];
            acc[j] += kw[dy + 2] * h;
        }
    }
    uint8_t *o = dst + (size_t)oy * ocols + ox0;
    if (ox0 + 3 < ocols && ((ocols & 3) == 0)) {
        uchar4 r;
        r.x = (acc[0] + 128) >> 8; r.y = (acc[1] + 128) >> 8; r.z = (acc[2] + 128) >> 8; r.w = (acc[3] + 128) >> 8;
        *reinterpret_cast<uchar4 *>(o) = r;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ox0 + j < ocols) o[j] = (uint8_t)((acc[j] + 128) >> 8);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K0  CLAHE (cv::createCLAHE(), clipLimit 3, 8x8 tiles: ViewController.mm:438-441, the step upstream of readImage; oracle r_clahe).
//     clahe_lut_kernel: one CTA per tile -- 256-bin histogram (shared-memory atomics), clip at max(1, int(clip * area / 256)), excess
//     redistributed as OpenCV does (uniform batch + one extra count every `step` bins), LUT[i] = cvRound(cdf[i] * 255 / area) in f32.
//     clahe_apply_kernel: per pixel the bilinear blend of the four neighbouring tile LUTs, every f32 operation rounded separately in
//     OpenCV's order ((l11 xa1 + l12 xa) ya1 + (l21 xa1 + l22 xa) ya), cvRound, saturate.  In place (each pixel reads only itself).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t *__restrict__ img, size_t img_stride, int cols, int tw, int th, int tiles_x,
                                                        int clip, float lut_scale, uint8_t *__restrict__ lut) {
    __shared__ int hist[256];
    __shared__ int wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const uint8_t *src = img + (size_t)b * img_stride + (size_t)(ty * th) * cols + tx * tw;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < tw * th; i += 256) { const int y = i / tw, x = i - y * tw; atomicAdd(&hist[src[(size_t)y * cols + x]], 1); }
    __syncthreads();
    int h = hist[tid];
    if (clip > 0) {
        int ex = max(h - clip, 0);
        h = min(h, clip);
        ex = warp_sum_i(ex);
        if (lane == 0) wsum[warp] = ex;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) clipped += wsum[w];
        const int batch = clipped / 256, resid = clipped - batch * 256;
        h += batch;
        if (resid > 0) { const int step = max(256 / resid, 1); if (tid % step == 0 && tid / step < resid) h++; }
        __syncthreads();
    }
    // inclusive prefix sum over the 256 bins
    int v = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; w++) base += wsum[w];
    const int cdf = v + base;
    const int r = __float2int_rn(fmul((float)cdf, lut_scale));
    lut[((size_t)b * gridDim.x + tile) * 256 + tid] = (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) clahe_apply_kernel(uint8_t *__restrict__ img, size_t img_stride, int rows, int cols, int tw, int th,
                                                          int tiles_x, int tiles_y, const uint8_t *__restrict__ lut) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, b = blockIdx.z;
    if (x0 >= cols) return;
    uint8_t *row = img + (size_t)b * img_stride + (size_t)y * cols;
    const uint8_t *L = lut + (size_t)b * tiles_x * tiles_y * 256;
    const float inv_tw = __fdiv_rn(1.f, (float)tw), inv_th = __fdiv_rn(1.f, (float)th);
    const float tyf = fsub(fmul((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    const float ya = fsub(tyf, (float)ty1), ya1 = fsub(1.f, ya);
    const int ty2 = min(ty1 + 1, tiles_y - 1);
    ty1 = max(ty1, 0);
    const uint8_t *P1 = L + (size_t)ty1 * tiles_x * 256, *P2 = L + (size_t)ty2 * tiles_x * 256;
    uint8_t out[4];
    const bool vec = (x0 + 3 < cols) && ((cols & 3) == 0) && ((img_stride & 3) == 0);
    uint8_t in[4];
    if (vec) { const uchar4 q = *reinterpret_cast<const uchar4 *>(row + x0); in[0] = q.x; in[1] = q.y; in[2] = q.z; in[3] = q.w; }
    else for (int k = 0; k < 4; k++) in[k] = x0 + k < cols ? row[x0 + k] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k;
        const float txf = fsub(fmul((float)x, inv_tw), 0.5f);
        int tx1 = (int)floorf(txf);
        const float xa = fsub(txf, (float)tx1), xa1 = fsub(1.f, xa);
        const int tx2 = min(tx1 + 1, tiles_x - 1);
        tx1 = max(tx1, 0);
        const int sv = in[k];
        const float l11 = (float)P1[tx1 * 256 + sv], l12 = (float)P1[tx2 * 256 + sv], l21 = (float)P2[tx1 * 256 + sv], l22 = (float)P2[tx2 * 256 + sv];
        const float res = fadd(fmul(fadd(fmul(l11, xa1), fmul(l12, xa)), ya1), fmul(fadd(fmul(l21, xa1), fmul(l22, xa)), ya));
        out[k] = (uint8_t)min(max(__float2int_rn(res), 0), 255);
    }
    if (vec) *reinterpret_cast<uchar4 *>(row + x0) = make_uchar4(out[0], out[1], out[2], out[3]);
    else for (int k = 0; k < 4; k++) if (x0 + k < cols) row[x0 + k] = out[k];
}

// ---------------------------------------------------------------------------------------------------------
// K4  pyramidal LK, one warp per feature (cv::calcOpticalFlowPyrLK(Size(21,21), 3); oracle r_lk_track).
//     Template: 24x24 u8 patch of I (REFLECT_101) staged in shared memory, Scharr gradients formed on the fly
//     (zero outside the image), 14-bit fixed-point bilinear weights, int16-range template values held in
//     registers (14 window pixels per lane).  All window PRODUCTS are exact integers; they are ACCUMULATED in f32 in the
//     order of OpenCV's SSE code (oracle _lk_sum_cov / _lk_sum_mismatch): per sum four "SIMD lane" chains (columns k, k+4, k+8,
//     k+12 of rows 0..20) and one scalar-tail chain (columns 16..20), each a strictly sequential f32 accumulation, combined as
//     tail + ((c0 + c2) + (c1 + c3)).  That order is what makes positions bit-identical to cv2, so it is reproduced literally:
//     the lanes that own the window pixels write their addends to shared memory in chain order and one lane per chain adds
//     them up (15 chains for the covariance sums once per level, 10 chains for the mismatch vector once per iteration).
//     Pixel ownership: lanes 0..23 = (SIMD lane k = lane / 6, sixth j = lane % 6) own the pair addends 7j..7j+6 of chain k
//     (addend a = row a >> 1, columns k + 8 (a & 1) and that + 4); lanes 24..31 own 14 consecutive tail pixels each.
// ---------------------------------------------------------------------------------------------------------
constexpr int LK_WARPS = 4;
constexpr int LK_JP = 32;                                 // edge of the staged J neighbourhood: the 22x22 window may drift +-5 px before a re-stage
constexpr int LK_JM = 5;
#ifndef LK_IP_V
#define LK_IP_V 32
#endif
constexpr int LK_IP = LK_IP_V;                            // row stride of the staged 24 x 24 I patch
constexpr int LK_BOX = 48;                                // width of the TMA box: a 16-byte aligned start + up to 33 columns
constexpr int LK_PIX = 14;                                // window pixels per lane (32 * 14 = 448 >= 441)
constexpr int LK_QS = 448;                                // covariance phase: floats per sum (4 chains x 84 + tail 105, padded to 112)
constexpr int LK_ACC = 3 * LK_QS;                         // the mismatch phase uses the first two thirds

__device__ __forceinline__ void lk_weights(float fx, float fy, int &w00, int &w01, int &w10, int &w11) {
    const float s = (float)(1 << W_BITS);
    const float omx = fsub(1.f, fx), omy = fsub(1.f, fy);
    w00 = __float2int_rn(fmul(fmul(omx, omy), s));
    w01 = __float2int_rn(fmul(fmul(fx, omy), s));
    w10 = __float2int_rn(fmul(fmul(omx, fy), s));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// stage an (nr x nc) u8 window whose top-left is (y0, x0) into dst (row stride LK_IP), 4 loads in flight per lane
__device__ __forceinline__ void lk_stage(uint8_t *dst, const uint8_t *__restrict__ im, int rows, int cols, int y0, int x0, int nr, int nc,
                                         int lane) {
    const int total = nr * nc;
    const bool inner = x0 >= 0 && x0 + nc <= cols && y0 >= 0 && y0 + nr <= rows;
    const uint8_t *base = im + (ptrdiff_t)y0 * cols + x0;
    for (int i0 = lane; i0 < total; i0 += 128) {
        uint8_t v[4];
        int d[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = i0 + 32 * t;
            const int r = i / nc, c = i - r * nc;
            d[t] = r * LK_IP + c;
            if (i < total)
                v[t] = inner ? base[r * cols + c] : im[(size_t)reflect101(y0 + r, rows) * cols + reflect101(x0 + c, cols)];
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (i0 + 32 * t < total) dst[d[t]] = v[t];
    }
}

// Stage the LK_JP x LK_JP neighbourhood of J whose top-left is (y0, x0) twice: A[r][c] = J(y0 + r, x0 + c) and B = A shifted left by one
// byte (B[i] = A[i + 1]), so that the horizontally adjacent pair (A[i], A[i + 1]) is ONE aligned 16-bit load whatever the parity of i
// (even: A + i, odd: B + i - 1).  One lane per row; inside the image the row is fetched as nine aligned words and re-aligned with funnel
// shifts, otherwise byte by byte with BORDER_REFLECT_101.
// one row of the J neighbourhood: 33 bytes starting at g (any alignment, global or shared) -> row `lane` of copy A and of the shifted copy B
__device__ __forceinline__ void lk_repack_row(uint8_t *A, uint8_t *Bc, const uint8_t *g, int lane) {
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const unsigned *wp = reinterpret_cast<const unsigned *>(ga & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(ga & 3) * 8;
    unsigned w[10];
#pragma unroll
    for (int i = 0; i < 10; i++) w[i] = wp[i];
    unsigned a[8], bq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = __funnelshift_rc(w[i], w[i + 1], sh); }
#pragma unroll
    for (int i = 0; i < 8; i++) { const unsigned nx = i < 7 ? a[i + 1] : __funnelshift_rc(w[8], w[9], sh); bq[i] = __funnelshift_r(a[i], nx, 8); }
    uint4 *da = reinterpret_cast<uint4 *>(A + lane * LK_JP), *db = reinterpret_cast<uint4 *>(Bc + lane * LK_JP);
    da[0] = make_uint4(a[0], a[1], a[2], a[3]); da[1] = make_uint4(a[4], a[5], a[6], a[7]);
    db[0] = make_uint4(bq[0], bq[1], bq[2], bq[3]); db[1] = make_uint4(bq[4], bq[5], bq[6], bq[7]);
}
// border case (cold): byte by byte with BORDER_REFLECT_101
__device__ __forceinline__ void lk_stage_j_border(uint8_t *A, uint8_t *Bc, const uint8_t *__restrict__ im, int rows, int cols, int y0, int x0, int lane) {
    const int yy = reflect101(y0 + lane, rows);
    const uint8_t *row = im + (size_t)yy * cols;
    uint8_t prev = row[reflect101(x0, cols)];
    for (int c = 0; c < LK_JP; c++) {
        const uint8_t nxt = row[reflect101(x0 + c + 1, cols)];
        A[lane * LK_JP + c] = prev; Bc[lane * LK_JP + c] = nxt;
        prev = nxt;
    }
}
__device__ __forceinline__ void lk_stage_j(uint8_t *A, uint8_t *Bc, const uint8_t *__restrict__ im, int rows, int cols, int y0, int x0, int lane) {
    const bool inner = x0 >= 4 && x0 + LK_JP + 4 <= cols && y0 >= 0 && y0 + LK_JP <= rows;
    if (inner) lk_repack_row(A, Bc, im + (size_t)(y0 + lane) * cols + x0, lane);
    else lk_stage_j_border(A, Bc, im, rows, cols, y0, x0, lane);
}

// two-way dot product of SIGNED 16-bit weights (low/high half of a) with the two low UNSIGNED bytes of b, plus c.  The fourth bilinear
// weight 2^14 - w00 - w01 - w10 is -1 when the three rounded weights add up to 2^14 + 1 (OpenCV keeps it as a signed short too).
__device__ __forceinline__ int lk_dp2a(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// one chain: strictly sequential f32 accumulation of 4 * n4 addends starting from +0 (adding the +0 pads changes nothing)
__device__ __forceinline__ float lk_chain(const float *p, int n4) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    float acc = 0.f;
#pragma unroll 3
    for (int i = 0; i < n4; i++) {
        const float4 v = q[i];
        acc = fadd(acc, v.x); acc = fadd(acc, v.y); acc = fadd(acc, v.z); acc = fadd(acc, v.w);
    }
    return acc;
}
// tail + ((c0 + c2) + (c1 + c3)): v_reduce_sum of the SSE accumulator, then added to the scalar accumulator
__device__ __forceinline__ float lk_combine(float r, int l0) {
    const float c0 = __shfl_sync(0xffffffffu, r, l0), c1 = __shfl_sync(0xffffffffu, r, l0 + 1), c2 = __shfl_sync(0xffffffffu, r, l0 + 2),
                c3 = __shfl_sync(0xffffffffu, r, l0 + 3);
    return fadd(fadd(c0, c2), fadd(c1, c3));
}

#ifndef LK_TMA_I
#define LK_TMA_I 1            // template patches through TMA
#endif
#ifndef LK_TMA_J
#define LK_TMA_J 0            // 1: J neighbourhoods through TMA as well (measured slower on B200: DESIGN.md section 4)
#endif
#ifndef LK_MINB
#define LK_MINB 4            // 128 registers: measured best on B200 (163 unconstrained, 96 and 80 spill in the iteration loop)
#endif
__global__ void __launch_bounds__(LK_WARPS * 32, LK_MINB) lk_kernel(PyrLevels I, PyrLevels J, LkMaps maps,
                                                           const float2 *__restrict__ prev_pts,
                                                           float2 *__restrict__ next_pts, uint8_t *__restrict__ status,
                                                           const int *__restrict__ n_pts, int maxp) {
    VIO_POISON(256u);
    __shared__ __align__(16) uint8_t sI[LK_WARPS][24 * LK_IP];          // I patch: rows ipy-1 .. ipy+22, cols ipx-1 .. ipx+22 (stride 32)
    __shared__ __align__(128) uint8_t sB[LK_WARPS][(1 + LK_TMA_J) * LK_BOX * 32];   // TMA landing boxes: [0] I patch, [1] J neighbourhood; re-aligned into sI / sJ
    __shared__ __align__(128) uint8_t sJ[LK_WARPS][2][LK_JP * LK_JP];   // copy A and the one-byte-shifted copy B (lk_stage_j)
    __shared__ __align__(16) float sAcc[LK_WARPS][LK_ACC];              // addends in chain order
    __shared__ __align__(8) unsigned long long sBar[LK_WARPS][2];       // two mbarriers per warp: [0] I-patch loads, [1] J-box loads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int f = blockIdx.x * LK_WARPS + warp;
    if (f >= n_pts[b]) return;
    if (lane == 0) { mbar_init(&sBar[warp][0], 1); mbar_init(&sBar[warp][1], 1); fence_proxy_async_smem(); }
    __syncwarp();
    // TMA loads are issued ahead of their use so that their latency hides behind the template build: the I patch of level l - 1 while
    // level l iterates, the first J box of a level while its template is built.  ipend / jpend = a load is in flight on the barrier.
    unsigned phI = 0, phJ = 0;
    bool ipend = false, jpend = false;
    int jp_x = 0, jp_y = 0;
    void *barI = &sBar[warp][0], *barJ = &sBar[warp][1];
    uint8_t *sIb = sB[warp], *sJb = sB[warp] + LK_TMA_J * LK_BOX * 32;
    const float2 pt = prev_pts[(size_t)b * maxp + f];
    uint8_t *sS = sI[warp];
    uint8_t *pJA = sJ[warp][0], *pJB = sJ[warp][1];
    float *acc = sAcc[warp];
    const float half = 10.f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    float nx = 0.f, ny = 0.f;          // nextPts[ptidx] (level coordinates, window centre)
    bool ok = true;                    // status[ptidx]

    // window pixels owned by this lane: woff = y * 32 + x (window coordinates), two per register; nvalid of them are real
    const bool is_tail = lane >= 24;
    const int ck = lane / 6, cj = lane - 6 * ck, ct = lane - 24;
    unsigned woffp[LK_PIX / 2];
#pragma unroll
    for (int s2 = 0; s2 < LK_PIX / 2; s2++) {
        int wo[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (!is_tail) {
                const int a = 7 * cj + s2;
                wo[h] = (a >> 1) * LK_JP + ck + 8 * (a & 1) + 4 * h;
            } else {
                const int a = min(14 * ct + 2 * s2 + h, 104);
                const int y = (a * 13108) >> 16;                         // a / 5 for a < 105
                wo[h] = y * LK_JP + 16 + a - 5 * y;
            }
        }
        woffp[s2] = (unsigned)wo[0] | ((unsigned)wo[1] << 16);
    }
    const int nvalid = is_tail ? min(105 - 14 * ct, LK_PIX) : LK_PIX;
    // where this lane's addends go (same layout for the covariance sums q = 0..2 and the mismatch sums q = 0, 1: + q * LK_QS): a SIMD
    // lane chain holds 6 x 14 floats, the tail chain 8 x 14.  In the mismatch phase a SIMD-lane owner writes (pair sum, +0) per two
    // pixels -- adding +0 changes nothing -- so that both kinds of lanes issue the same 8-byte stores.
    const int off_a = is_tail ? 336 + 14 * ct : 84 * ck + 14 * cj;
    // which chain this lane adds up: covariance phase lane = 5 q + c (c = 4: tail), mismatch phase the same with q = 0, 1
    const int qa = lane / 5, ca = lane - 5 * qa;
    const float *chain_a = acc + qa * LK_QS + (ca < 4 ? 84 * ca : 336);
    const int n4_a = lane < 15 ? (ca < 4 ? 21 : 27) : 0;
    const int n4_b = lane < 10 ? n4_a : 0;
    // cv::buildOpticalFlowPyramid stops at the last level whose successor would not be larger than the window in both directions
    int top = 0;
    while (top < LK_LEVELS && min(I.rows[top + 1], I.cols[top + 1]) > LK_WIN) top++;
#pragma unroll 1
    for (int level = top; level >= 0; level--) {
        const int rows = I.rows[level], cols = I.cols[level];
        const uint8_t *__restrict__ imI = I.p[level] + (size_t)b * I.stride[level];
        const uint8_t *__restrict__ imJ = J.p[level] + (size_t)b * J.stride[level];
        const float scale = 1.f / (float)(1 << level);
        float px = fmul(pt.x, scale), py = fmul(pt.y, scale);
        if (level == top) { nx = px; ny = py; }
        else { nx = fmul(nx, 2.f); ny = fmul(ny, 2.f); }
        px = fsub(px, half); py = fsub(py, half);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -LK_WIN || ipx >= cols || ipy < -LK_WIN || ipy >= rows) {
            if (level == 0) ok = false;
            continue;
        }
        __syncwarp();
        if (jpend) { mbar_wait(barJ, phJ); phJ ^= 1u; jpend = false; }          // an unused prefetch of the previous level
        const uint8_t *pI = sS;                                          // patch origin = image (ipy - 1, ipx - 1)
        const bool tma_i = LK_TMA_I && maps.ok[level] && ipx >= 1 && ipx + 23 <= cols && ipy >= 1 && ipy + 23 <= rows;
        if (tma_i) {
            // the 24 x 24 template neighbourhood lies inside the image: it arrives as ONE TMA box that starts at the 16-byte boundary below
            // its first column (what hangs over the right / bottom edge reads as 0 and is never used); normally the box was requested
            // while the previous level iterated
            const int xa = (ipx - 1) & ~15;
            if (!ipend) {
                if (lane == 0) {
                    fence_proxy_async_smem();                           // earlier generic-proxy accesses of this warp to the buffer are done (__syncwarp above)
                    mbar_expect_tx(barI, LK_BOX * 32);
                    tma_load_3d(sIb, maps.I + level, xa, ipy - 1, b, barI);
                }
            }
#if LK_TMA_J
            // the first J neighbourhood of this level is known already (nextPt of the level above): fetch it while the template is built
            {
                const int jx0 = (int)floorf(fsub(nx, half)) - LK_JM, jy0 = (int)floorf(fsub(ny, half)) - LK_JM;
                if (jx0 >= 0 && jx0 + LK_JP + 1 <= cols && jy0 >= 0 && jy0 + LK_JP <= rows) {
                    if (lane == 0) {
                        fence_proxy_async_smem();
                        mbar_expect_tx(barJ, LK_BOX * 32);
                        tma_load_3d(sJb, maps.J + level, jx0 & ~15, jy0, b, barJ);
                    }
                    jpend = true; jp_x = jx0; jp_y = jy0;
                }
            }
#endif
            mbar_wait(barI, phI);
            phI ^= 1u; ipend = false;
            // re-align the box into the patch layout the template build uses (fixed origin, stride 32): 24 rows x 3 eight-byte pieces
            {
                const int d = ipx - 1 - xa;
                for (int t = lane; t < 72; t += 32) {
                    const int r = (t * 21846) >> 16, c8 = (t - 3 * r) * 8;                 // t / 3 for t < 72
                    const uint8_t *g = sIb + r * LK_BOX + d + c8;
                    const unsigned *wp = reinterpret_cast<const unsigned *>(g - ((d + c8) & 3));   // rows start 16-byte aligned
                    const unsigned sh = (unsigned)((d + c8) & 3) * 8;
                    const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2];
                    *reinterpret_cast<uint2 *>(sS + r * LK_IP + c8) = make_uint2(__funnelshift_rc(w0, w1, sh), __funnelshift_rc(w1, w2, sh));
                }
            }
        } else {
            if (ipend) { mbar_wait(barI, phI); phI ^= 1u; ipend = false; }        // cannot happen (a prefetch is only issued for a TMA-able level); keeps the phase honest
            lk_stage(sS, imI, rows, cols, ipy - 1, ipx - 1, 24, 24, lane);      // rows ipy-1 .. ipy+22, cols ipx-1 .. ipx+22
        }
        __syncwarp();
        int w00, w01, w10, w11;
        lk_weights(fsub(px, (float)ipx), fsub(py, (float)ipy), w00, w01, w10, w11);
        // pass 1: Scharr derivatives at the 22x22 source pixels of the bilinear window (window pixel (yy, xx) = patch (yy+1, xx+1)),
        // zero where the source lies outside the image -- each is used by up to four window pixels.  The J staging area is free here.
        short2 *sD = reinterpret_cast<short2 *>(pJA);
        for (int i = lane; i < 22 * 22; i += 32) {
            const int yy = (i * 2979) >> 16, xx = i - 22 * yy;           // i / 22 for i < 484
            const uint8_t *q = pI + (yy + 1) * LK_IP + (xx + 1);
            const int gyi = ipy + yy, gxi = ipx + xx;
            short2 d = make_short2(0, 0);
            if (gyi >= 0 && gyi < rows && gxi >= 0 && gxi < cols) {
                const int tl = q[-LK_IP - 1], tc = q[-LK_IP], tr = q[-LK_IP + 1], ml = q[-1], mr = q[1], bl = q[LK_IP - 1], bc = q[LK_IP], br = q[LK_IP + 1];
                d.x = (short)(3 * (tr - tl) + 10 * (mr - ml) + 3 * (br - bl));
                d.y = (short)(3 * (bl - tl) + 10 * (bc - tc) + 3 * (br - tr));
            }
            sD[i] = d;
        }
        __syncwarp();
        // pass 2: bilinear template (Iw, gx, gy) of the lane's own window pixels, kept in registers (Iw | gx << 16; the gy of two
        // pixels share a register); covariance products to the chains
        int tIG[LK_PIX];
        unsigned tGy[LK_PIX / 2];
#pragma unroll
        for (int i = 0; i < LK_PIX; i += 2) {
            float pa[2], pb[2], pc[2];
            int gyp[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int wo = h ? (int)(woffp[i >> 1] >> 16) : (int)(woffp[i >> 1] & 0xffffu);
                const int y = wo >> 5, x = wo & 31;
                const uint8_t *q = pI + (y + 1) * LK_IP + (x + 1);
                const short2 *dq = sD + y * 22 + x;
                const short2 d00 = dq[0], d01 = dq[1], d10 = dq[22], d11 = dq[23];
                const int iv = w00 * q[0] + w01 * q[1] + w10 * q[LK_IP] + w11 * q[LK_IP + 1];
                const int dxv = w00 * d00.x + w01 * d01.x + w10 * d10.x + w11 * d11.x;
                const int dyv = w00 * d00.y + w01 * d01.y + w10 * d10.y + w11 * d11.y;
                const bool valid = i + h < nvalid;
                const int ivs = valid ? (iv + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5) : 0;      // 0 .. 8160
                const int gxs = valid ? (dxv + (1 << (W_BITS - 1))) >> W_BITS : 0;
                const int gys = valid ? (dyv + (1 << (W_BITS - 1))) >> W_BITS : 0;
                tIG[i + h] = ivs | (gxs << 16);
                gyp[h] = gys;
                pa[h] = (float)(gxs * gxs); pb[h] = (float)(gxs * gys); pc[h] = (float)(gys * gys);     // RN of the exact product, as the f32 multiply
            }
            tGy[i >> 1] = ((unsigned)gyp[0] & 0xffffu) | ((unsigned)gyp[1] << 16);
            *reinterpret_cast<float2 *>(acc + off_a + i) = make_float2(pa[0], pa[1]);
            *reinterpret_cast<float2 *>(acc + LK_QS + off_a + i) = make_float2(pb[0], pb[1]);
            *reinterpret_cast<float2 *>(acc + 2 * LK_QS + off_a + i) = make_float2(pc[0], pc[1]);
        }
        __syncwarp();
        if (level > 0) {
            // the I patch is dead from here on: fetch the next level's patch (its position depends on the point alone)
            const int nrows = I.rows[level - 1], ncols = I.cols[level - 1];
            const float sc2 = 1.f / (float)(1 << (level - 1));
            const int qx = (int)floorf(fsub(fmul(pt.x, sc2), half)), qy = (int)floorf(fsub(fmul(pt.y, sc2), half));
            if (LK_TMA_I && maps.ok[level - 1] && qx >= 1 && qx + 23 <= ncols && qy >= 1 && qy + 23 <= nrows) {
                if (lane == 0) {
                    fence_proxy_async_smem();
                    mbar_expect_tx(barI, LK_BOX * 32);
                    tma_load_3d(sIb, maps.I + (level - 1), (qx - 1) & ~15, qy - 1, b, barI);
                }
                ipend = true;
            }
        }
        const float ra = lk_chain(chain_a, n4_a);
        __syncwarp();
        const float A11 = fmul(fadd(__shfl_sync(0xffffffffu, ra, 4), lk_combine(ra, 0)), FLT_SCALE);
        const float A12 = fmul(fadd(__shfl_sync(0xffffffffu, ra, 9), lk_combine(ra, 5)), FLT_SCALE);
        const float A22 = fmul(fadd(__shfl_sync(0xffffffffu, ra, 14), lk_combine(ra, 10)), FLT_SCALE);
        const float D = fsub(fmul(A11, A22), fmul(A12, A12));
        const float d12 = fsub(A11, A22);
        const float mineig = __fdiv_rn(fsub(fadd(A22, A11), __fsqrt_rn(fadd(fmul(d12, d12), fmul(fmul(4.f, A12), A12)))),
                                       (float)(2 * LK_WIN * LK_WIN));
        if (mineig < 1e-4f || D < 1.1920929e-07f) {
            if (level == 0) ok = false;
            continue;
        }
        const float Dinv = __fdiv_rn(1.f, D);
        float cx = fsub(nx, half), cy = fsub(ny, half);       // nextPt (window top-left, float)
        float pdx = 0.f, pdy = 0.f;
        int sx0 = 0, sy0 = 0;
        bool staged = false;
#pragma unroll 1
        for (int j = 0; j < LK_MAX_ITERS; j++) {
            const int jx = (int)floorf(cx), jy = (int)floorf(cy);
            if (jx < -LK_WIN || jx >= cols || jy < -LK_WIN || jy >= rows) {
                if (level == 0) ok = false;
                break;
            }
            // the window needs rows jy .. jy+21 and columns jx .. jx+21 of J: re-stage only when it leaves the staged neighbourhood
            if (!staged || jx < sx0 || jx + 22 > sx0 + LK_JP || jy < sy0 || jy + 22 > sy0 + LK_JP) {
                sx0 = jx - LK_JM; sy0 = jy - LK_JM; staged = true;
                __syncwarp();
                if (jpend) { mbar_wait(barJ, phJ); phJ ^= 1u; }                  // the prefetched box has landed (used below if it is the right one)
                const bool hit = jpend && jp_x == sx0 && jp_y == sy0;
                jpend = false;
                if (LK_TMA_J && maps.ok[level] && sx0 >= 0 && sx0 + LK_JP + 1 <= cols && sy0 >= 0 && sy0 + LK_JP <= rows) {
                    // interior: the aligned 48 x 32 box that contains the 33 x 32 neighbourhood arrives by TMA; each lane then re-aligns one
                    // row into copy A and the one-byte-shifted copy B
                    const int xa = sx0 & ~15;
                    if (!hit) {
                        if (lane == 0) {
                            fence_proxy_async_smem();
                            mbar_expect_tx(barJ, LK_BOX * 32);
                            tma_load_3d(sJb, maps.J + level, xa, sy0, b, barJ);
                        }
                        mbar_wait(barJ, phJ);
                        phJ ^= 1u;
                    }
                    lk_repack_row(pJA, pJB, sJb + lane * LK_BOX + (sx0 - xa), lane);
                } else {
                    lk_stage_j(pJA, pJB, imJ, rows, cols, sy0, sx0, lane);
                }
                __syncwarp();
            }
            int v00, v01, v10, v11;
            lk_weights(fsub(cx, (float)jx), fsub(cy, (float)jy), v00, v01, v10, v11);
            const unsigned wt = ((unsigned)v00 & 0xffffu) | ((unsigned)v01 << 16), wb = ((unsigned)v10 & 0xffffu) | ((unsigned)v11 << 16);
            const int o0 = (jy - sy0) * LK_JP + (jx - sx0);
            // addends in chain order: a SIMD lane adds the int32 dot product of two pixels (v_dotprod), the tail one pixel at a time
#pragma unroll
            for (int i = 0; i < LK_PIX; i += 2) {
                int p1[2], p2[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int wo = h ? (int)(woffp[i >> 1] >> 16) : (int)(woffp[i >> 1] & 0xffffu);
                    const int off = o0 + wo;                              // (jy - sy0 + y) * LK_JP + (jx - sx0 + x)
                    const uint8_t *src = pJA + off + (off & 1) * (LK_JP * LK_JP - 1);     // odd: copy B at off - 1 -> aligned 16-bit load of (A[off], A[off + 1])
                    const unsigned top = *reinterpret_cast<const unsigned short *>(src);
                    const unsigned bot = *reinterpret_cast<const unsigned short *>(src + LK_JP);
                    const int jv = lk_dp2a(wt, top, lk_dp2a(wb, bot, 1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                    const int diff = jv - (tIG[i + h] & 0xffff);
                    const int gy = h ? (int)tGy[i >> 1] >> 16 : (int)(short)(tGy[i >> 1] & 0xffffu);
                    p1[h] = diff * (tIG[i + h] >> 16);
                    p2[h] = diff * gy;
                }
                const float2 u = is_tail ? make_float2((float)p1[0], (float)p1[1]) : make_float2((float)(p1[0] + p1[1]), 0.f);
                const float2 v = is_tail ? make_float2((float)p2[0], (float)p2[1]) : make_float2((float)(p2[0] + p2[1]), 0.f);
                *reinterpret_cast<float2 *>(acc + off_a + i) = u;
                *reinterpret_cast<float2 *>(acc + LK_QS + off_a + i) = v;
            }
            __syncwarp();
            const float rb = lk_chain(chain_a, n4_b);
            __syncwarp();
            const float B1 = fmul(fadd(__shfl_sync(0xffffffffu, rb, 4), lk_combine(rb, 0)), FLT_SCALE);
            const float B2 = fmul(fadd(__shfl_sync(0xffffffffu, rb, 9), lk_combine(rb, 5)), FLT_SCALE);
            const float dx = fmul(fsub(fmul(A12, B2), fmul(A22, B1)), Dinv);
            const float dy = fmul(fsub(fmul(A12, B1), fmul(A11, B2)), Dinv);
            cx = fadd(cx, dx); cy = fadd(cy, dy);
            nx = fadd(cx, half); ny = fadd(cy, half);
            const double dd = (double)dx * (double)dx + (double)dy * (double)dy;
            if (dd <= 0.01 * 0.01) break;
            if (j > 0 && fabsf(fadd(dx, pdx)) < 0.01f && fabsf(fadd(dy, pdy)) < 0.01f) {
                nx = fsub(nx, fmul(dx, 0.5f));
                ny = fsub(ny, fmul(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (level == 0 && ok) {
            const int fx = (int)floorf(fsub(nx, half)), fy = (int)floorf(fsub(ny, half));
            if (fx < -LK_WIN || fx >= cols || fy < -LK_WIN || fy >= rows) ok = false;
        }
    }
    // no TMA write may still be in flight into this CTA's shared memory when the warp exits
    if (jpend) mbar_wait(barJ, phJ);
    if (ipend) mbar_wait(barI, phI);
    if (lane == 0) {
        next_pts[(size_t)b * maxp + f] = make_float2(nx, ny);
        status[(size_t)b * maxp + f] = ok ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2  Shi-Tomasi response + 3x3 NMS + mask test, candidate compaction (cv::goodFeaturesToTrack front half;
//     oracle r_min_eig_map / r_good_features).  The eig map is never written to HBM.  One WARP streams down a vertical strip:
//     lane = image column (32 input columns -> 26 output columns, 3 columns of halo on either side), one image row per step,
//     everything the separable pipeline needs from the rows above kept in registers:
//       pixel row -> Sobel (f32, the exact fma pattern cv2 uses) -> covariance products -> 3x3 box sums in f64 in cv2's order
//       (row sum (left + centre) + right first, then (above + row) + below) -> min-eigenvalue -> 3x3 NMS,
//     horizontal neighbours through warp shuffles.  Borders: pixels are REFLECT_101 for the Sobel stage; the box filter reflects
//     the COVARIANCE image (column -1 := column 1, row -1 := row 1 -- not the same thing: dx changes sign under pixel reflection),
//     which is a matter of which lane / which register a neighbour is taken from.
//     Emits (value, address) of every masked interior local maximum plus the masked global maximum.
//     The min-distance mask of setMask() is evaluated analytically: a pixel is masked iff it lies within
//     min_dist of a kept point's ROUNDED centre (cv::circle raster == Euclidean disc).
//     NMS is threshold independent: (e > thr) && (e == dilate(threshold(e)))  <=>  (e > thr) && e >= 8 nbrs.
// ---------------------------------------------------------------------------------------------------------
constexpr int EG_W = 26;               // output columns per warp (32 lanes - 2 x 3 halo)
constexpr int EG_ROWS = 64;            // output rows per warp (4 rows of halo are recomputed per chunk)
constexpr int EG_WARPS = 4;
constexpr int EG_KCAP = 64;            // kept points whose disc can reach one strip chunk

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ float eg_min_eig(double sxx, double sxy, double syy) {
    const float fa = (float)sxx, fb = (float)sxy, fc = (float)syy;
    const float ha = fmul(fa, 0.5f), hc = fmul(fc, 0.5f);
    const float t = fsub(ha, hc);
    return fsub(fadd(ha, hc), __fsqrt_rn(fadd(fmul(t, t), fmul(fb, fb))));
}

__global__ void __launch_bounds__(EG_WARPS * 32) eig_candidates_kernel(const uint8_t *__restrict__ img, size_t img_stride, int rows,
                                                                       int cols, const int2 *__restrict__ kept, const int *__restrict__ n_kept,
                                                                       int maxp, int min_dist, unsigned *__restrict__ max_bits,
                                                                       unsigned long long *__restrict__ cand, int *__restrict__ cand_cnt) {
    VIO_POISON(512u);
    __shared__ int2 sk[EG_WARPS][EG_KCAP];
    __shared__ short sw[128];                                        // chord half widths of the min-distance disc (min_dist <= 127)
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    for (int d = threadIdx.x; d <= min(min_dist, 127); d += blockDim.x) {
        const int v = min_dist * min_dist - d * d;
        int w = (int)sqrtf((float)v);
        while (w * w > v) w--;
        while ((w + 1) * (w + 1) <= v) w++;
        sw[d] = (short)w;
    }
    __syncthreads();
    const int x0 = (blockIdx.x * EG_WARPS + warp) * EG_W;
    if (x0 >= cols) return;                                          // whole warp
    const int y0 = blockIdx.y * EG_ROWS, y1 = min(y0 + EG_ROWS, rows);
    const uint8_t *im = img + (size_t)b * img_stride;
    // kept points whose disc can touch this strip chunk (warp-level compaction)
    int nk = 0;
    {
        const int n = n_kept[b];
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            bool near = false;
            int2 c = make_int2(0, 0);
            if (i < n) {
                c = kept[(size_t)b * maxp + i];
                near = c.x >= x0 - min_dist && c.x < x0 + EG_W + min_dist && c.y >= y0 - min_dist && c.y < y1 + min_dist;
            }
            const unsigned m = __ballot_sync(FULL, near);
            const int pos = nk + __popc(m & ((1u << lane) - 1u));
            if (near && pos < EG_KCAP) sk[warp][pos] = c;
            nk += __popc(m);
        }
        nk = min(nk, EG_KCAP);
        __syncwarp();
    }
    const int md2 = min_dist * min_dist;
    const int xin = x0 - 3 + lane;                                   // image column of this lane
    const int xr = reflect101(xin, cols);                            // where its pixels come from (Sobel: BORDER_REFLECT_101)
    const bool own_x = lane >= 3 && lane < 3 + EG_W && xin < cols;   // columns whose results this warp publishes
    const int pl = max(lane - 1, 0), pr = min(lane + 1, 31);         // pixel neighbours
    const int cl = (xin == 0) ? pr : pl, cr = (xin == cols - 1) ? pl : pr;      // covariance neighbours (box filter border)
    const float a = (float)(1.0 / 3060.0);
    const float a2 = fmul(2.f, a);
    const int r_start = max(0, y0 - 2), r_end = min(rows - 1, y1 + 1);
    auto ldpix = [&](int r) { return (float)im[(size_t)reflect101(r, rows) * cols + xr]; };
    // per pixel row k: rd[k] = p[k][x+1] - p[k][x-1],  sv[k] = fma(p[k][x+1], a, fma(p[k][x], 2a, a p[k][x-1]))
    auto row_terms = [&](float p, float &rd, float &sv) {
        const float L = __shfl_sync(FULL, p, pl), R = __shfl_sync(FULL, p, pr);
        rd = fsub(R, L);
        sv = ffma(R, a, ffma(p, a2, fmul(a, L)));
    };
    float rd_a, sv_a, rd_b, sv_b, rd_c, sv_c;
    row_terms(ldpix(r_start - 1), rd_a, sv_a);
    row_terms(ldpix(r_start), rd_b, sv_b);
    float q0 = ldpix(r_start + 1), q1 = ldpix(r_start + 2), q2 = ldpix(r_start + 3);      // pixel rows in flight (three steps ahead)
    double xx0 = 0, xy0 = 0, yy0 = 0, xx1 = 0, xy1 = 0, yy1 = 0;     // row sums of rows r-2, r-1
    float e0 = -1.f, e1 = -1.f, e2 = -1.f;                           // eig rows ye-2, ye-1, ye
    float local_max = 0.f;
    unsigned mrow1 = 0, mrow2 = 0;                                   // lanes masked by a kept point's disc in eig rows ye-1, ye
    unsigned long long *cout = cand + (size_t)b * CAND_CAP;
    // setMask() for one image row as a lane bit mask: lane k owns kept point k, turns its disc's chord on that row (half width
    // sw[|dy|] = floor(sqrt(min_dist^2 - dy^2)), tabulated once per CTA) into a run of lane bits; the runs are OR-ed across the warp
    auto row_mask = [&](int gy) {
        unsigned m = 0;
        for (int k0 = 0; k0 < nk; k0 += 32) {
            unsigned bits = 0;
            if (k0 + lane < nk) {
                const int2 c = sk[warp][k0 + lane];
                const int dy = abs(gy - c.y);
                if (dy <= min_dist) {
                    const int w = sw[dy];
                    const int lo = max(c.x - w - (x0 - 3), 0), hi = min(c.x + w - (x0 - 3), 31);
                    if (lo <= hi) bits = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
                }
            }
            m |= __reduce_or_sync(FULL, bits);
        }
        return m;
    };
    // a new eig row ye: (1) running masked maximum over the pixels this warp owns, (2) NMS of row ye - 1
    auto push_eig = [&](float e, int ye) {
        e0 = e1; e1 = e2; e2 = e;
        mrow1 = mrow2; mrow2 = nk ? row_mask(ye) : 0u;
        if (own_x && ye >= y0 && ye < y1 && !((mrow2 >> lane) & 1u)) local_max = fmaxf(local_max, e);
        const int y = ye - 1;
        const float vm = fmaxf(fmaxf(e0, e1), e2);
        const float vl = __shfl_sync(FULL, vm, pl), vr = __shfl_sync(FULL, vm, pr);
        if (own_x && y >= y0 && y < y1 && y >= 1 && y < rows - 1 && xin >= 1 && xin < cols - 1 && e1 > 0.f && !((mrow1 >> lane) & 1u)) {
            const float m = fmaxf(fmaxf(vl, vr), fmaxf(e0, e2));
            if (e1 >= m) {
                const int sl = atomicAdd(&cand_cnt[b], 1);
                if (sl < CAND_CAP) cout[sl] = ((unsigned long long)__float_as_uint(e1) << 32) | (unsigned)(y * cols + xin);
            }
        }
    };
#pragma unroll 1
    for (int r = r_start; r <= r_end; r++) {
        row_terms(q0, rd_c, sv_c);
        q0 = q1; q1 = q2;
        if (r + 4 <= r_end + 1) {                                    // pixel row r + 1 is needed at step r: loads run three steps ahead
            const int rr = r + 4 < rows ? r + 4 : 2 * rows - 2 - (r + 4);
            q2 = (float)im[(size_t)rr * cols + xr];
        }
        const float dx = ffma(fadd(rd_a, rd_c), a, fmul(rd_b, a2));
        const float dy = fsub(sv_c, sv_a);
        rd_a = rd_b; sv_a = sv_b; rd_b = rd_c; sv_b = sv_c;
        const double pxx = (double)fmul(dx, dx), pxy = (double)fmul(dx, dy), pyy = (double)fmul(dy, dy);
        const double xx2 = __dadd_rn(__dadd_rn(shfl_d(pxx, cl), pxx), shfl_d(pxx, cr));
        const double xy2 = __dadd_rn(__dadd_rn(shfl_d(pxy, cl), pxy), shfl_d(pxy, cr));
        const double yy2 = __dadd_rn(__dadd_rn(shfl_d(pyy, cl), pyy), shfl_d(pyy, cr));
        if (r == 1)                                                  // eig row 0: box rows (1, 0, 1)
            push_eig(eg_min_eig(__dadd_rn(__dadd_rn(xx2, xx1), xx2), __dadd_rn(__dadd_rn(xy2, xy1), xy2), __dadd_rn(__dadd_rn(yy2, yy1), yy2)), 0);
        if (r - r_start >= 2)                                        // eig row r - 1: box rows (r - 2, r - 1, r)
            push_eig(eg_min_eig(__dadd_rn(__dadd_rn(xx0, xx1), xx2), __dadd_rn(__dadd_rn(xy0, xy1), xy2), __dadd_rn(__dadd_rn(yy0, yy1), yy2)), r - 1);
        if (r == rows - 1 && r >= 1)                                 // eig row rows - 1: box rows (rows - 2, rows - 1, rows - 2)
            push_eig(eg_min_eig(__dadd_rn(__dadd_rn(xx1, xx2), xx1), __dadd_rn(__dadd_rn(xy1, xy2), xy1), __dadd_rn(__dadd_rn(yy1, yy2), yy1)), rows - 1);
        xx0 = xx1; xy0 = xy1; yy0 = yy1; xx1 = xx2; xy1 = xy2; yy1 = yy2;
    }
    // masked global maximum (positive floats order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(FULL, local_max, o));
    if (lane == 0 && local_max > 0.f) atomicMax(&max_bits[b], __float_as_uint(local_max));
}

}  // namespace fe
