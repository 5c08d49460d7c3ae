// be_tilechol.cuh -- K10/K11 on the FP64 tensor pipe (mma.sync.m8n8k4.f64): Schur elimination of the landmark blocks and the dense
// Cholesky solve of the reduced pose / speed-bias system of one window, one CTA (16 warps) per stream.
//   build    S = Sc H Sc + mu D^2, rows/cols permuted (all pose dofs, then all speed-bias dofs, the right-hand side as border row),
//            packed lower triangle in shared memory
//   schur    S -= sum_l v_l v_l^T  (schur_eliminator_impl.h:170-298) -- a SYRK over the pose block: every warp owns up to four 8x8
//            tiles, one DMMA per tile and four landmarks, operands staged in shared memory in chunks of TC_CHUNK landmarks
//   factor   LEFT-looking blocked Cholesky (Eigen LLT of schur_complement_solver.cc:123-224), 8 columns per panel: a warp loads
//            its 8x8 tile of the panel into DMMA accumulators, subtracts the contributions of all finished panels in one long
//            register-resident K loop (no read-modify-write of shared memory, one barrier), the owner of the diagonal tile factors
//            it in registers by shuffles and forms L_kk^-1 alongside, the others apply L_kk^-T with two more DMMAs.
//            The border row makes z = L^-1 rhs a by-product.
//   back     L^T y = z block by block with the stored L_kk^-1 (an 8x8 mat-vec instead of a sequential substitution).
// Requires 512 threads and at most 16 panel tiles per warp pair (window <= 10).
#pragma once
#include "be_kernels.cuh"

namespace be {

constexpr int TC_CHUNK = 32;            // landmarks staged per SYRK chunk
constexpr int TC_LD = 88;               // doubles per staged landmark: 72 pose dofs (66 used) + 8 border (index 7 = rhs term) + pad; 88 = 8 mod 16
                                         // puts the four k-lanes of a fragment load on two disjoint bank halves

struct TileGeom { int npose, ppad, nsb, npad, nt, rhs_row; };
__host__ __device__ inline TileGeom tile_geom(int NF) {
    TileGeom g;
    g.npose = 6 * NF; g.ppad = (g.npose + 7) & ~7; g.nsb = 9 * NF;
    g.npad = (g.ppad + g.nsb + 1 + 7) & ~7; g.nt = g.npad / 8; g.rhs_row = g.npad - 1;
    return g;
}
__host__ __device__ inline bool tile_path_fits(int NF, int threads) {
    const TileGeom g = tile_geom(NF);
    return threads == 512 && g.nt <= 31 && g.ppad + 16 <= TC_LD && (g.ppad / 8) * (g.ppad / 8 + 1) / 2 + g.ppad / 8 <= 64;
}
// shared memory of the tile path, in doubles: packed L (+ border row) | L_kk^-1 blocks | landmark staging
__host__ __device__ inline size_t tile_smem_doubles(int NF) {
    const TileGeom g = tile_geom(NF);
    return (size_t)g.npad * (g.npad + 1) / 2 + (size_t)g.nt * 64 + (size_t)TC_CHUNK * TC_LD + 8 + (g.npad + 3) / 4;
}
// permuted row -> canonical dof of the [pose_i(6) sb_i(9)] layout; -1 = padding (unit diagonal), -2 = right-hand-side border row
__device__ __forceinline__ int tc_canon(const TileGeom &G, int r) {
    if (r < G.npose) { const int f = r / 6; return 15 * f + (r - 6 * f); }
    if (r < G.ppad) return -1;
    const int q = r - G.ppad;
    if (q < G.nsb) { const int f = q / 9; return 15 * f + 6 + (q - 9 * f); }
    return r == G.rhs_row ? -2 : -1;
}
__device__ __forceinline__ int tc_pidx(int i, int j) { return i * (i + 1) / 2 + j; }

__device__ __forceinline__ void tc_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 1 / x and 1 / sqrt(x) for a positive, finite, normal-range x: hardware approximation (MUFU.RCP64H / MUFU.RSQ64H, ~2^-23) + two
// Newton steps; no fix-up branches on the pivot chain
__device__ __forceinline__ double tc_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double tc_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}

// In-register Cholesky of one 8x8 diagonal tile held by a warp in accumulator layout (lane (g = lane >> 2, t = lane & 3) holds
// D[g][2t], D[g][2t+1]), together with W = L^-1 (same layout).  npiv = 8, or 7 for the last tile whose row 7 is the border row (it is
// scaled and updated like any other row below the pivots, which IS the forward substitution; row 7 of W is zero).
// The pivot recurrence runs on 1 / pivot with ONE shuffle stage per column (the unscaled column is broadcast and the update uses
// a_gj a_cj / pivot); the scaling by 1 / sqrt(pivot) and the inverse are off the critical chain.  Branch-free.  Measured 2.0 k cycles
// (tools/ubench/diag_factor.cu; 3.4 k for the textbook order).  Returns false (warp-uniform) on a non-positive or non-finite pivot.
__device__ __forceinline__ bool tc_diag_factor(double &d0, double &d1, double &w0, double &w1, int npiv, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const int g = lane >> 2, t = lane & 3;
    double p0 = 0.0, p1 = 0.0;                                   // partial sums  P[g][c] = sum_{m<j} L[g][m] W[m][c]
    w0 = 0.0; w1 = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double cur = (j & 1) ? d1 : d0;                    // the register that holds column j
        const double piv = __shfl_sync(FULL, cur, 4 * j + (j >> 1));
        const double agj = __shfl_sync(FULL, cur, 4 * g + (j >> 1));              // D[g][j]
        const double ac0 = __shfl_sync(FULL, cur, 4 * (2 * t) + (j >> 1));        // D[2t][j]
        const double ac1 = __shfl_sync(FULL, cur, 4 * (2 * t + 1) + (j >> 1));    // D[2t+1][j]
        const bool act = j < npiv;
        ok &= !act || ((piv > 0.0) && (piv < 1e300));
        const double inv = act ? tc_rcp(piv) : 0.0;
        const double sg = (g > j) ? agj * inv : 0.0;
        d0 = fma(-((2 * t > j && 2 * t <= g) ? sg : 0.0), ac0, d0);
        d1 = fma(-((2 * t + 1 > j && 2 * t + 1 <= g) ? sg : 0.0), ac1, d1);
        const double il = act ? tc_rsqrt(piv) : 0.0;
        const double lgj = agj * il;                             // L[g][j] for g >= j
        if (j & 1) d1 = (act && t == (j >> 1) && g >= j) ? lgj : d1;
        else d0 = (act && t == (j >> 1) && g >= j) ? lgj : d0;
        // row j of W: W[j][c] = il * (delta_jc - P[j][c]), c <= j; then P[g][c] += L[g][j] W[j][c] for g > j
        const double nw0 = (2 * t <= j) ? il * ((2 * t == j ? 1.0 : 0.0) - p0) : 0.0;
        const double nw1 = (2 * t + 1 <= j) ? il * ((2 * t + 1 == j ? 1.0 : 0.0) - p1) : 0.0;
        w0 = (act && g == j) ? nw0 : w0; w1 = (act && g == j) ? nw1 : w1;
        const double wj0 = __shfl_sync(FULL, w0, 4 * j + t), wj1 = __shfl_sync(FULL, w1, 4 * j + t);
        const double mg = (act && g > j) ? lgj : 0.0;
        p0 = fma(mg, wj0, p0); p1 = fma(mg, wj1, p1);
    }
    // zero the strict upper triangle of L (it is never read, but the tile is stored as a whole)
    if (2 * t > g) d0 = 0.0;
    if (2 * t + 1 > g) d1 = 0.0;
    if (g >= npiv) { w0 = 0.0; w1 = 0.0; }
    return ok;
}

// Solves  (Sc H Sc + mu D^2 - landmark Schur terms) y = Sc g - ...  for the pose / speed-bias block.  Inputs as build_reduced_smem /
// chol_solve_packed (be_solve.cuh); output ws_y[NP] in canonical order.  sm: tile_smem_doubles() doubles of shared memory.
// Returns false in all threads when a pivot fails (caller raises mu, as DoglegStrategy does when the linear solver fails).
__device__ __noinline__ bool tile_reduced_solve(const BeState &s, const double *H, const double *gvec, const double *w, const double *hll,
                                                const double *gl, const double *sc_p, const double *sc_l, const double *d_p, const double *d_l,
                                                double *u_l, int nl, double mu, double *ws_y, double *sm, int *sh_flag, long long *pp) {
    BE_PROF2_INIT;
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NP = s.NPS, NPW = s.NPWS;
    const TileGeom G = tile_geom(s.NFS);
    const int nt = G.nt, npad = G.npad;
    const int npt = G.ppad / 8, nsy = npt * (npt + 1) / 2 + npt;       // pose tile rows; SYRK tiles = pose x pose (lower) + border x pose
    double *Lp = sm;
    double *Winv = Lp + (size_t)npad * (npad + 1) / 2;
    double *stage = Winv + (size_t)nt * 64;
    short *canon = reinterpret_cast<short *>(stage + (size_t)TC_CHUNK * TC_LD + 8);    // [npad] permuted row -> canonical dof

    for (int r = tid; r < npad; r += T) canon[r] = (short)tc_canon(G, r);
    if (tid == 0) *sh_flag = 1;
    for (int l = tid; l < nl; l += T) {                             // per-landmark weight sqrt(s_l^2 / (h_l s_l^2 + mu d_l^2))
        const double sl = sc_l[l];
        u_l[l] = sqrt(sl * sl / (hll[l] * sl * sl + mu * d_l[l] * d_l[l]));
    }
    __syncthreads();
    // ---- build: S = Sc H Sc + mu D^2 in permuted order, border row = Sc g (one warp per row: no index arithmetic per element) -------
    for (int r = warp; r < npad; r += 16) {
        const int ci = canon[r];
        double *row = Lp + tc_pidx(r, 0);
        const double sci = ci >= 0 ? sc_p[ci] : 1.0;
        for (int c = lane; c <= r; c += 32) {
            const int cj = canon[c];
            double v = 0.0;
            if (ci >= 0 && cj >= 0) {
                const int hi = max(ci, cj), lo = min(ci, cj);
                v = H[(size_t)hi * NP + lo] * sci * sc_p[cj];
                if (r == c) v += mu * d_p[ci] * d_p[ci];
            } else if (ci == -2 && cj >= 0) v = gvec[cj] * sc_p[cj];
            else if (ci == -1 && r == c) v = 1.0;
            row[c] = v;
        }
    }
    // ---- schur: S -= sum_l ve_l ve_l^T with ve_l = [ w_l u_l Sc (pose dofs) | 0.. | g_l u_l at the border index ] --------------------
    {
        int ao[4], bo[4], tI[4], tJ[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int li = 16 * q + warp;
            ao[q] = bo[q] = G.ppad + 8;                             // zero columns of the staging rows: contributes nothing
            tI[q] = tJ[q] = -1;
            if (li < nsy - npt) {
                int J = 0;
                while ((J + 1) * npt - (J + 1) * J / 2 <= li) J++;
                const int I = J + li - (J * npt - J * (J - 1) / 2);
                ao[q] = 8 * I; bo[q] = 8 * J; tI[q] = I; tJ[q] = J;
            } else if (li < nsy) { const int J = li - (nsy - npt); ao[q] = G.ppad; bo[q] = 8 * J; tI[q] = nt - 1; tJ[q] = J; }
        }
        double sy[4][2];
#pragma unroll
        for (int q = 0; q < 4; q++) sy[q][0] = sy[q][1] = 0.0;
        for (int l0 = 0; l0 < nl; l0 += TC_CHUNK) {
            const int cn = min(TC_CHUNK, nl - l0);
            __syncthreads();
            {
                constexpr int NE = (TC_CHUNK * TC_LD + 511) / 512;   // elements per thread: all loads are issued before the first store
                double v[NE];
#pragma unroll
                for (int i = 0; i < NE; i++) {
                    const int e = tid + i * T;
                    const int cl = e / TC_LD, a = e - cl * TC_LD, l = l0 + cl;
                    v[i] = 0.0;
                    if (e < TC_CHUNK * TC_LD && cl < cn) {
                        if (a < NPW) { const int f = a / 6; v[i] = w[(size_t)l * NPW + a] * u_l[l] * sc_p[15 * f + (a - 6 * f)]; }
                        else if (a == G.ppad + 7) v[i] = gl[l] * u_l[l];
                    }
                }
#pragma unroll
                for (int i = 0; i < NE; i++) { const int e = tid + i * T; if (e < TC_CHUNK * TC_LD) stage[e] = v[i]; }
            }
            __syncthreads();
            const double *base = stage + t * TC_LD + g;
#pragma unroll
            for (int k4 = 0; k4 < TC_CHUNK; k4 += 4)
#pragma unroll
                for (int q = 0; q < 4; q++) tc_dmma(sy[q][0], sy[q][1], -base[k4 * TC_LD + ao[q]], base[k4 * TC_LD + bo[q]]);
        }
        __syncthreads();                                            // build complete (and the last chunk consumed)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (tI[q] < 0) continue;
            if (tI[q] == nt - 1) {                                  // border row: only its last row (the right-hand side) is real
                if (g == 7) { double *d = Lp + tc_pidx(G.rhs_row, 8 * tJ[q] + 2 * t); d[0] += sy[q][0]; d[1] += sy[q][1]; }
            } else {
                const int r = 8 * tI[q] + g, c = 8 * tJ[q] + 2 * t;
                if (c <= r) Lp[tc_pidx(r, c)] += sy[q][0];
                if (c + 1 <= r) Lp[tc_pidx(r, c + 1)] += sy[q][1];
            }
        }
    }
    __syncthreads();
    BE_PROF2(pp, 2);
    // ---- factor: left-looking, panel k = columns 8k .. 8k+7.  Warp 0 owns the diagonal tile: its pivot chain (8 x rsqrt + scale +
    //      update) is a string of dependent FP64 operations, and the FP64 pipe of its SM sub-partition is shared with warps 4, 8 and
    //      12 -- so those three sit the K loops out (measured: the chain runs ~3x slower against a saturated DMMA pipe).  The other
    //      twelve warps take the tiles (I, k), I > k: worker wi owns I = k + 1 + wi and I + 12. ------------------------------------
    const int wi = (warp & 3) ? (warp >> 2) * 3 + (warp & 3) - 1 : -1;
    BE_PROF_ONLY(long long c_k = 0, c_f = 0, c_a = 0, c_b = 0, c_t;)
    for (int k = 0; k < nt; k++) {
        const int npiv = (k == nt - 1) ? 7 : 8;
        const int I0 = k + 1 + wi, I1 = I0 + 12;
        const bool h0 = wi >= 0 && I0 < nt, h1 = wi >= 0 && I1 < nt;
        const int rb0 = tc_pidx(8 * min(I0, nt - 1) + g, 0), rb1 = tc_pidx(8 * min(I1, nt - 1) + g, 0);
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
        BE_PROF_ONLY(c_t = clock64();)
        if (warp == 0) {
            const int rb = tc_pidx(8 * k + g, 0);
            double d0 = (2 * t <= g) ? Lp[rb + 8 * k + 2 * t] : 0.0, d1 = (2 * t + 1 <= g) ? Lp[rb + 8 * k + 2 * t + 1] : 0.0;
            double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, h0_ = 0.0, h1_ = 0.0;   // four accumulators: a quarter of the dependent DMMA chain
            const double *pa = Lp + rb + t;
            // operands of the next group are loaded before the MMAs of the current one are issued (software pipelining by hand)
            int j = 0;
            const int nk = 2 * k;
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
            if (nk >= 4) { n0 = pa[0]; n1 = pa[4]; n2 = pa[8]; n3 = pa[12]; }
            for (; j + 3 < nk; j += 4) {
                const double a0 = n0, a1 = n1, a2 = n2, a3 = n3;
                if (j + 7 < nk) { n0 = pa[4 * j + 16]; n1 = pa[4 * j + 20]; n2 = pa[4 * j + 24]; n3 = pa[4 * j + 28]; }
                tc_dmma(d0, d1, -a0, a0);
                tc_dmma(e0, e1, -a1, a1);
                tc_dmma(f0, f1, -a2, a2);
                tc_dmma(h0_, h1_, -a3, a3);
            }
            if (j < nk) {
                const double a0 = pa[4 * j], a1 = pa[4 * j + 4];
                tc_dmma(d0, d1, -a0, a0);
                tc_dmma(e0, e1, -a1, a1);
            }
            d0 += (e0 + f0) + h0_; d1 += (e1 + f1) + h1_;
            BE_PROF_ONLY({ const long long c = clock64(); c_k += c - c_t; c_t = c; })
            double w0, w1;
            const bool ok = tc_diag_factor(d0, d1, w0, w1, npiv, lane);
            if (!ok && lane == 0) *sh_flag = 0;
            if (2 * t <= g) Lp[rb + 8 * k + 2 * t] = d0;
            if (2 * t + 1 <= g) Lp[rb + 8 * k + 2 * t + 1] = d1;
            double *Wd = Winv + k * 64;
            Wd[g * 8 + 2 * t] = w0; Wd[g * 8 + 2 * t + 1] = w1;
            BE_PROF_ONLY({ const long long c = clock64(); c_f += c - c_t; c_t = c; })
        } else if (h0) {
            // x -= L[I, 0:8k] L[k, 0:8k]^T, kept in registers until L_kk^-1 is known
            x0 = Lp[rb0 + 8 * k + 2 * t]; x1 = Lp[rb0 + 8 * k + 2 * t + 1];
            const double *pb = Lp + tc_pidx(8 * k + g, 0) + t, *p0 = Lp + rb0 + t, *p1 = Lp + rb1 + t;
            if (h1) {
                y0 = Lp[rb1 + 8 * k + 2 * t]; y1 = Lp[rb1 + 8 * k + 2 * t + 1];
                double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
                const int nk = 2 * k;
                double nb0 = 0.0, na0 = 0.0, nc0 = 0.0, nb1 = 0.0, na1 = 0.0, nc1 = 0.0;
                if (nk >= 2) { nb0 = pb[0]; na0 = p0[0]; nc0 = p1[0]; nb1 = pb[4]; na1 = p0[4]; nc1 = p1[4]; }
                for (int j = 0; j + 1 < nk; j += 2) {
                    const double b0 = nb0, a0 = na0, c0 = nc0, b1 = nb1, a1 = na1, c1 = nc1;
                    if (j + 3 < nk) { nb0 = pb[4 * j + 8]; na0 = p0[4 * j + 8]; nc0 = p1[4 * j + 8]; nb1 = pb[4 * j + 12]; na1 = p0[4 * j + 12]; nc1 = p1[4 * j + 12]; }
                    tc_dmma(x0, x1, -a0, b0);
                    tc_dmma(y0, y1, -c0, b0);
                    tc_dmma(e0, e1, -a1, b1);
                    tc_dmma(f0, f1, -c1, b1);
                }
                x0 += e0; x1 += e1; y0 += f0; y1 += f1;
            } else {
                double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, q0 = 0.0, q1 = 0.0;
                int j = 0;
                const int nk = 2 * k;
                double nb[4] = {0.0, 0.0, 0.0, 0.0}, na[4] = {0.0, 0.0, 0.0, 0.0};
                if (nk >= 4) {
#pragma unroll
                    for (int q = 0; q < 4; q++) { nb[q] = pb[4 * q]; na[q] = p0[4 * q]; }
                }
                for (; j + 3 < nk; j += 4) {
                    double b[4], a[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) { b[q] = nb[q]; a[q] = na[q]; }
                    if (j + 7 < nk) {
#pragma unroll
                        for (int q = 0; q < 4; q++) { nb[q] = pb[4 * (j + 4 + q)]; na[q] = p0[4 * (j + 4 + q)]; }
                    }
                    tc_dmma(x0, x1, -a[0], b[0]);
                    tc_dmma(e0, e1, -a[1], b[1]);
                    tc_dmma(f0, f1, -a[2], b[2]);
                    tc_dmma(q0, q1, -a[3], b[3]);
                }
                if (j < nk) {
                    const double b0 = pb[4 * j], a0 = p0[4 * j], b1 = pb[4 * j + 4], a1 = p0[4 * j + 4];
                    tc_dmma(x0, x1, -a0, b0);
                    tc_dmma(e0, e1, -a1, b1);
                }
                x0 += (e0 + f0) + q0; x1 += (e1 + f1) + q1;
            }
        }
        __syncthreads();                                            // L_kk and W_k are in shared memory
        BE_PROF_ONLY({ const long long c = clock64(); c_a += c - c_t; c_t = c; })
        if (!*sh_flag) return false;
        if (h0) {
            const double *Wk = Winv + k * 64;
            const double b0 = Wk[g * 8 + t], b1 = Wk[g * 8 + 4 + t];        // B[t'][n = g] = W[g][4c + t']
            // A operand chunk c: A[g][4c + t'] sits in lane (g, 2c + (t' >> 1)), register t' & 1
            const int s0 = 4 * g + (t >> 1), s1 = s0 + 2;
            const double u00 = __shfl_sync(FULL, x0, s0), u01 = __shfl_sync(FULL, x1, s0), u10 = __shfl_sync(FULL, x0, s1), u11 = __shfl_sync(FULL, x1, s1);
            const double v00 = __shfl_sync(FULL, y0, s0), v01 = __shfl_sync(FULL, y1, s0), v10 = __shfl_sync(FULL, y0, s1), v11 = __shfl_sync(FULL, y1, s1);
            double rx0 = 0.0, rx1 = 0.0, ry0 = 0.0, ry1 = 0.0;
            tc_dmma(rx0, rx1, (t & 1) ? u01 : u00, b0);
            tc_dmma(ry0, ry1, (t & 1) ? v01 : v00, b0);
            tc_dmma(rx0, rx1, (t & 1) ? u11 : u10, b1);
            tc_dmma(ry0, ry1, (t & 1) ? v11 : v10, b1);
            Lp[rb0 + 8 * k + 2 * t] = rx0; Lp[rb0 + 8 * k + 2 * t + 1] = rx1;
            if (h1) { Lp[rb1 + 8 * k + 2 * t] = ry0; Lp[rb1 + 8 * k + 2 * t + 1] = ry1; }
        }
        __syncthreads();                                            // panel k complete
        BE_PROF_ONLY({ const long long c = clock64(); c_b += c - c_t; c_t = c; })
    }
    BE_PROF_ONLY(if (tid == 0) { pp[21] += c_k; pp[22] += c_f; pp[24] += c_a; pp[25] += c_b; })
    BE_PROF2(pp, 20);
    // ---- back substitution  L^T y = z,  z = border row ---------------------------------------------------------------------------
    double *z = Lp + tc_pidx(G.rhs_row, 0);
    const int n = G.rhs_row;                                        // unknowns 0 .. n-1
    for (int k = nt - 1; k >= 0; k--) {
        const int c0 = 8 * k;
        if (tid < 8) {                                              // y_k = W_k^T z_k
            const double *Wk = Winv + k * 64;
            double y = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) { const double zj = (c0 + j < n) ? z[c0 + j] : 0.0; y += Wk[j * 8 + tid] * zj; }
            stage[tid] = y;
        }
        __syncthreads();
        if (tid < 8 && c0 + tid < n) z[c0 + tid] = stage[tid];
        for (int i = tid; i < c0; i += T) {                         // fold into the rows above
            double v = z[i];
#pragma unroll
            for (int j = 0; j < 8; j++) if (c0 + j < n) v -= Lp[tc_pidx(c0 + j, i)] * stage[j];
            z[i] = v;
        }
        __syncthreads();
    }
    bool ok = true;
    for (int r = tid; r < n; r += T) {
        const int c = canon[r];
        if (c >= 0) { const double v = z[r]; ws_y[c] = v; ok &= isfinite(v); }
    }
    BE_PROF2(pp, 23);
    return __syncthreads_and(ok) != 0;
}

}  // namespace be
