// be_solve.cuh -- K9-K12: the window solve, one persistent CTA per stream (VINS::solve_ceres, VINS.cpp:480-682).
//
// What ceres::Solve does for this problem (DENSE_SCHUR + DOGLEG, vendored ceres-solver 1.12.0) restated on the normal
// equations, so that no Jacobian is ever materialised:
//   TrustRegionMinimizer::Minimize            internal/ceres/trust_region_minimizer.cc:66-119
//   EvaluateGradientAndJacobian (Jacobi scaling frozen at iteration 0)          :231-279
//   DoglegStrategy::ComputeStep / Cauchy point / Gauss-Newton step / traditional dogleg   dogleg_strategy.cc:77-255,515-572
//   SchurEliminator::Eliminate + dense LLT + BackSubstitute   schur_eliminator_impl.h:170-298, schur_complement_solver.cc:123-224
//   TrustRegionStepEvaluator::StepQuality     trust_region_step_evaluator.cc:51-59
//   termination tests                         trust_region_minimizer.cc:616-705
// Notation: H = J^T J (unscaled, block structure [pose/speed-bias NP x NP | landmark 1x1 blocks | coupling w_l (6 per observing
// frame)]), g = J^T r, s = Jacobi scaling 1/(1+sqrt(diag H at iteration 0)).  Every quantity Ceres forms from the scaled
// Jacobian J*S is obtained from H and g:  (JS)^T(JS) = S H S,  ||J S v||^2 = (S v)^T H (S v).
// The eliminated set is exactly the inverse-depth blocks; Ceres' own ordering additionally eliminates 6 of the pose blocks
//  -- the Schur step is exact for any independent set, so the Gauss-Newton step is the same up to round-off.
#pragma once
#include "be_kernels.cuh"

namespace be {

constexpr int SOLVE_T = 512;

struct SolveWs {
    double *H, *S, *g, *hll, *gl, *w;
    double *sc_p, *sc_l, *d_p, *d_l, *gr_p, *gr_l, *gn_p, *gn_l, *st_p, *st_l, *u_p, *u_l, *rhs, *y;
    double *dx, *lt;
};

// The matrices (H, S, w, per-factor landmark terms) live in the stream's global scratch; the vectors every phase of the dogleg loop
// touches live in shared memory (sv), the landmark-sized ones only when the problem has at most SOLVE_LV landmarks.
constexpr int SOLVE_LV = 384;
__host__ __device__ inline size_t solve_vec_doubles(int NP, int NPX) {
    auto r = [](size_t n) { return (n + 3) & ~(size_t)3; };
    return 9 * r(NP) + r(NPX) + 8 * r(SOLVE_LV);
}
__device__ inline SolveWs carve(const BeState &s, int b, double *sv, int nl) {
    SolveWs w;
    double *p = s.scratch + (size_t)b * s.scratch_stride;
    auto take = [&](size_t n) { double *r = p; p += (n + 3) & ~(size_t)3; return r; };
    auto stake = [&](size_t n) { double *r = sv; sv += (n + 3) & ~(size_t)3; return r; };
    w.H = take((size_t)s.NPS * s.NPS); w.S = take((size_t)s.NPS * s.NPS); w.w = take((size_t)s.LCAP * s.NPWS); w.lt = take((size_t)s.PCAP * 8);
    w.g = stake(s.NPS); w.sc_p = stake(s.NPS); w.d_p = stake(s.NPS); w.gr_p = stake(s.NPS); w.gn_p = stake(s.NPS); w.st_p = stake(s.NPS);
    w.u_p = stake(s.NPS); w.rhs = stake(s.NPS); w.y = stake(s.NPS); w.dx = stake(s.NPX);
    if (nl <= SOLVE_LV) {
        w.hll = stake(SOLVE_LV); w.gl = stake(SOLVE_LV); w.sc_l = stake(SOLVE_LV); w.d_l = stake(SOLVE_LV);
        w.gr_l = stake(SOLVE_LV); w.gn_l = stake(SOLVE_LV); w.st_l = stake(SOLVE_LV); w.u_l = stake(SOLVE_LV);
    } else {
        w.hll = take(s.LCAP); w.gl = take(s.LCAP); w.sc_l = take(s.LCAP); w.d_l = take(s.LCAP);
        w.gr_l = take(s.LCAP); w.gn_l = take(s.LCAP); w.st_l = take(s.LCAP); w.u_l = take(s.LCAP);
    }
    return w;
}
__host__ __device__ inline size_t solve_scratch_doubles(int NP, int NPX, int NPW, int LCAP, int PCAP) {
    auto r = [](size_t n) { return (n + 3) & ~(size_t)3; };
    return 2 * r((size_t)NP * NP) + 7 * r(NP) + 8 * r(LCAP) + r((size_t)LCAP * NPW) + r(NPX) + r((size_t)PCAP * 8) + r(NP) * 2 + 64;
}

// prior dx over the canonical layout (MarginalizationFactor::Evaluate, marginalization_factor.cpp:340-366)
__device__ inline void prior_dx(const BeState &s, int b, const double *par, double *dx) {
    const int tid = threadIdx.x;
    const double *x0 = s.x0 + (size_t)b * (s.NF * 16 + 7);
    const int *pres = s.present + (size_t)b * (2 * s.NF + 1);
    for (int i = tid; i < s.NF; i += blockDim.x) {
        const double *x = par + 16 * i, *z = x0 + 16 * i;
        double *d = dx + 15 * i;
        if (pres[2 * i]) {
            d[0] = x[0] - z[0]; d[1] = x[1] - z[1]; d[2] = x[2] - z[2];
            const Q4 qd = qmul(qinv(ldq(z + 3)), ldq(x + 3));
            const double sg = (qd.w >= 0) ? 2.0 : -2.0;
            d[3] = sg * qd.x; d[4] = sg * qd.y; d[5] = sg * qd.z;
        } else for (int k = 0; k < 6; k++) d[k] = 0;
        if (pres[2 * i + 1]) for (int k = 0; k < 9; k++) d[6 + k] = x[7 + k] - z[7 + k];
        else for (int k = 0; k < 9; k++) d[6 + k] = 0;
    }
    for (int k = tid; k < 6; k += blockDim.x) dx[s.NP + k] = 0.0;      // para_Ex_Pose is constant: x == x0
}

// D (8x8, f64) += A (8x4, row) * B (4x8, col): lane (g = lane >> 2, k = lane & 3) supplies A[g][k] and B[k][g] and holds
// D[g][2k], D[g][2k+1]
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// adds a warp's accumulated pair blocks to H / g.  hacc[ps] = {cross c0,c1, jj c0,c1, ii c0,c1} fragments of pair (round*4+ps)*nwarp+warp.
// The (j,i) block has a single owner (plain read-modify-write); the diagonal blocks and the gradient are shared between pairs.
__device__ __forceinline__ void proj_flush(const SolveWs &ws, int NP, int NF, int npair, int nwarp, int warp, int lane, int round, const double (*hacc)[6]) {
    const int g = lane >> 2, c = 2 * (lane & 3);
#pragma unroll
    for (int ps = 0; ps < 4; ps++) {
        const int q = (round * 4 + ps) * nwarp + warp;
        if (q >= npair || g >= 6) continue;
        int i = 0, rem = q;
        while (rem >= NF - 1 - i) { rem -= NF - 1 - i; i++; }
        const int j = i + 1 + rem;
        double *Hj = ws.H + (size_t)(15 * j + g) * NP, *Hi = ws.H + (size_t)(15 * i + g) * NP;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int cc = c + e;
            if (cc < 6) {
                Hj[15 * i + cc] += hacc[ps][e];
                if (cc <= g) { atomic_add(&Hj[15 * j + cc], hacc[ps][2 + e]); atomic_add(&Hi[15 * i + cc], hacc[ps][4 + e]); }
            } else if (cc == 6) {
                atomic_add(&ws.g[15 * j + g], hacc[ps][e]);
                atomic_add(&ws.g[15 * i + g], hacc[ps][4 + e]);
            }
        }
    }
}

// cost (and, when lin != 0, H / g / landmark terms) at parameter vector `par`.  Returns the total cost in every thread.
__device__ __noinline__ double evaluate(const BeState &s, int b, const SolveWs &ws, const double *par, int lin, double *sh_red, double *smem_scratch) {
    long long *pp = s.prof + (size_t)b * 32; BE_PROF2_INIT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int *iv = S_iv(s, b);
    const int NP = s.NPS, NPW = s.NPWS, NPwin = s.NP, nl = iv[IV_N_LM], nfac = iv[IV_N_FAC_ALL], nfac_reg = iv[IV_N_FAC];
    double cost = 0.0;
    const int has_prior = iv[IV_PRIOR_VALID];
    const int *pres = s.present + (size_t)b * (2 * s.NF + 1);
    if (lin) {
        // H (lower triangle; the upper one is never read) starts as the prior's J0^T J0 over the pose / speed-bias part of the
        // canonical layout, or zero.  One warp per row: coalesced, no index arithmetic.  Rows of blocks the prior does not contain
        // are zero in Hp (marg_kernel writes them so) and are not read.
        const double *Hp = s.Hp + (size_t)b * s.NPX * s.NPX;
        for (int i = warp; i < NP; i += nwarp) {
            double *row = ws.H + (size_t)i * NP;
            const double *src = Hp + (size_t)i * s.NPX;
            const int fi = i / 15;
            const bool rp = has_prior && fi < s.NF && pres[2 * fi + ((i - 15 * fi) >= 6)];       // the loop-closure pose has no prior
            for (int j = lane; j <= i; j += 32) row[j] = rp ? src[j] : 0.0;
        }
        for (int i = tid; i < NP; i += blockDim.x) ws.g[i] = 0.0;
        // w_l rows: the entries of observing frames are rewritten by every linearisation and the others stay zero for the whole solve
        if (lin == 2) for (int i = tid; i < nl * NPW; i += blockDim.x) ws.w[i] = 0.0;
    }
    // ---- prior ------------------------------------------------------------------------------------------
    if (has_prior) {
        const int NPX = s.NPX, T = blockDim.x;
        const double *Hp = s.Hp + (size_t)b * NPX * NPX, *bp = s.bp + (size_t)b * NPX;
        prior_dx(s, b, par, ws.dx);
        // t = Hp dx over the dofs the prior contains (typically all poses + the first speed-bias block: 75 of 171): compact list cl[],
        // rows split over P thread groups (Hp symmetric: thread ii of a group walks the column of its dof, neighbouring threads read
        // neighbouring addresses); the loads of a thread are issued eight at a time so that the L2 latency is paid once per eight rows
        double *part_sum = smem_scratch;                                   // [P][nc]
        double *dxc = smem_scratch + 4 * NPX;                              // [nc] dx of the listed dofs
        int *cl = reinterpret_cast<int *>(dxc + NPX);                      // [nc] canonical dof
        int *ncp = cl + NPX;
        if (tid == 0) *ncp = 0;
        __syncthreads();
        for (int d = tid; d < NPX; d += T) {
            const int f = d / 15, blk = d < NPwin ? 2 * f + ((d - 15 * f) >= 6) : 2 * s.NF;
            if (pres[blk]) {
                int pos = 0;
                for (int q = 0; q < blk; q++) pos += pres[q] ? ((q & 1) ? 9 : 6) : 0;
                pos += d < NPwin ? ((blk & 1) ? d - 15 * f - 6 : d - 15 * f) : d - NPwin;
                cl[pos] = d; dxc[pos] = ws.dx[d];
            }
        }
        if (tid == 0) { int n = 0; for (int q = 0; q <= 2 * s.NF; q++) n += pres[q] ? (q == 2 * s.NF ? 6 : ((q & 1) ? 9 : 6)) : 0; *ncp = n; }
        __syncthreads();
        const int nc = *ncp;
        const int P = nc > 0 ? min(4, max(1, T / nc)) : 1, part = nc > 0 ? tid / nc : T;
        const int rows = (nc + P - 1) / P;
        if (part < P) {
            const int ii = tid - part * nc;
            const int j0 = part * rows, j1 = min(nc, (part + 1) * rows);
            const double *col = Hp + cl[ii];
            double t = 0;
            int j = j0;
            for (; j + 8 <= j1; j += 8) {
                double h[8];
#pragma unroll
                for (int q = 0; q < 8; q++) h[q] = __ldg(col + (size_t)cl[j + q] * NPX);
#pragma unroll
                for (int q = 0; q < 8; q++) t += h[q] * dxc[j + q];
            }
            for (; j < j1; j++) t += __ldg(col + (size_t)cl[j] * NPX) * dxc[j];
            part_sum[part * nc + ii] = t;
        }
        __syncthreads();
        for (int i2 = tid; i2 < nc; i2 += T) {
            double t = 0;
            for (int q = 0; q < P; q++) t += part_sum[q * nc + i2];
            const int d = cl[i2];
            cost += 0.5 * dxc[i2] * t + bp[d] * dxc[i2];
            if (lin && d < NPwin) ws.g[d] += t + bp[d];
        }
        if (tid == 0) cost += 0.5 * S_dv(s, b)[DV_PRIOR_C0];
    }
    __syncthreads();
    if (lin) BE_PROF2(pp, 16); else BE_PROF2(pp, 30);
    // ---- IMU factors: one warp per factor ---------------------------------------------------------------
    for (int f = warp; f < s.W; f += nwarp) {
        const double *pr = S_pre(s, b, f + 1);
        // raw J 450 | weighted J 450 | raw res 15 | weighted res 15, in the shared scratch (free outside ComputeGaussNewtonStep)
        double *J = smem_scratch + (size_t)f * 930;
        double *Jw = J + 450, *rr = J + 900, *rw = J + 915;
        imu_residual_warp(pr, s.gravity, par + 16 * f, par + 16 * f + 7, par + 16 * (f + 1), par + 16 * (f + 1) + 7, rr, lin ? J : nullptr, lane);
        __syncwarp();
        const double *U = pr + PR_SQI;                   // upper triangular
        double r_w = 0.0;                                // lanes 0..14 hold the weighted residual
        if (lane < 15) { for (int k = lane; k < 15; k++) r_w += U[lane * 15 + k] * rr[k]; cost += 0.5 * r_w * r_w; rw[lane] = r_w; }
        __syncwarp();
        if (lin) {
            for (int e = lane; e < 450; e += 32) {
                const int r = e / 30, c = e - r * 30;
                double t = 0;
                for (int k = r; k < 15; k++) t += U[r * 15 + k] * J[k * 30 + c];
                Jw[e] = t;
            }
            __syncwarp();
            const int off = 15 * f;
            for (int e = lane; e < 900; e += 32) {
                const int r = e / 30, c = e - r * 30;
                if (c > r) continue;                      // lower triangle only; the upper one is mirrored after the accumulation
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + r] * Jw[k * 30 + c];
                atomic_add(&ws.H[(size_t)(off + r) * NP + off + c], t);
            }
            for (int c = lane; c < 30; c += 32) {
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + c] * rw[k];
                atomic_add(&ws.g[off + c], t);
            }
        }
    }
    __syncthreads();
    if (lin) BE_PROF2(pp, 17); else BE_PROF2(pp, 31);
    // ---- projection factors ---------------------------------------------------------------------------------
    // The factors are visited in (anchor frame i, observing frame j) order (prepare_kernel's fac_sorted / pair_off).  Everything
    // that depends on the frame pair alone -- ric^T Rj^T Ri ric, the translation part, the rotation products of the Jacobians --
    // is formed once per pair in shared memory, so a factor costs ~150 FMAs.  Linearisation runs in chunks of blockDim factors:
    //   J-phase: one thread per factor writes its corrected Ji, Jj (2x6 each) and residual to shared memory (SoA) and adds the
    //            landmark terms (w_l, h_ll, g_l) -- 8 atomics per factor;
    //   H-phase: one thread per (pair, block part) sums Jj^T Ji, Jj^T Jj, Ji^T Ji, J^T r over the pair's factors in registers and
    //            issues one atomic per output element and chunk (about 5k instead of 140k atomics per linearisation).
    {
        const double *dv = S_dv(s, b);
        const M3 ric = ldm(dv + DV_RIC), ricT = tr(ric);
        const V3 tic = ld3(dv + DV_TIC);
        const int T = blockDim.x, NF = s.NFS, npair = NF * (NF - 1) / 2;
        const int *lml = s.lm_loop + (size_t)b * s.LCAP;
        double *fr = smem_scratch, *prs = fr + 21 * NF, *Js = prs + 33 * npair + ((33 * npair + 21 * NF) & 1);
        for (int f = tid; f < NF; f += T) {
            const M3 R = q2R(ldq(par + 16 * f + 3));
            stm(fr + 21 * f, R); st3(fr + 21 * f + 9, ld3(par + 16 * f)); stm(fr + 21 * f + 12, ricT * tr(R));
        }
        __syncthreads();
        for (int q = tid; q < npair; q += T) {
            int i = 0, rem = q;
            while (rem >= NF - 1 - i) { rem -= NF - 1 - i; i++; }
            const int j = i + 1 + rem;
            const M3 Ri = ldm(fr + 21 * i), RjT = tr(ldm(fr + 21 * j));
            const M3 RR = RjT * Ri;
            const V3 d = RjT * (ld3(fr + 21 * i + 9) - ld3(fr + 21 * j + 9));
            const M3 AR = ricT * RR;
            double *o = prs + 33 * q;
            stm(o, AR * ric); st3(o + 9, ricT * (RR * tic + d - tic)); stm(o + 12, AR); stm(o + 21, RR); st3(o + 30, d);
        }
        __syncthreads();
        BE_PROF_ONLY(long long *pq2 = pp; long long _qt1 = clock64();
                     if (lin && tid == 0) { pq2[27] += _qt1 - _qt0; })
        const int *fs = s.fac_sorted + (size_t)b * s.PCAP, *po = s.pair_off + (size_t)b * (NF * NF + 1);
        const double *fobs = s.fac_obs + (size_t)b * s.PCAP * 4;
        const double *lam_p = par + 16 * NF;
        const int *lf0 = s.lm_fac0 + (size_t)b * (s.LCAP + 1);
        double hacc[4][6];
#pragma unroll
        for (int ps = 0; ps < 4; ps++)
#pragma unroll
            for (int e = 0; e < 6; e++) hacc[ps][e] = 0.0;
        const bool hold = npair <= 4 * nwarp;
        for (int c0 = 0; c0 < nfac; c0 += T) {
            const int pos = c0 + tid;
            if (pos < nfac) {
                const int rec = fs[pos], l = rec & 0xffff, i = (rec >> 16) & 0xff, j = (rec >> 24) & 0xff;
                const double4 ob = *reinterpret_cast<const double4 *>(fobs + 4 * (size_t)pos);
                const double lam = lam_p[l], inv = 1.0 / lam;
                const double *pq = prs + 33 * (i * NF - i * (i + 1) / 2 + (j - i - 1));
                const V3 pci = v3(ob.x * inv, ob.y * inv, inv);
                const V3 pcj = ldm(pq) * pci + ld3(pq + 9);
                const double dep = pcj.z;
                const double rx = s.sqrt_info * (pcj.x / dep - ob.z), ry = s.sqrt_info * (pcj.y / dep - ob.w);
                const double sq = rx * rx + ry * ry, sum = 1.0 + sq;
                cost += 0.5 * log(sum);                                  // CauchyLoss(1.0): rho = log(1 + s)
                if (lin) {
                    const double sr = sqrt(fmax(2.2250738585072014e-308, 1.0 / sum));   // corrector.cc:113-118 (rho'' < 0: alpha = 0)
                    const double r0 = s.sqrt_info / dep * sr, r2x = -s.sqrt_info * pcj.x / (dep * dep) * sr, r2y = -s.sqrt_info * pcj.y / (dep * dep) * sr;
                    // corrected 2x3 "reduce" = sr * [r0 0 r2x; 0 r0 r2y]; red(X) = reduce * X for a 3x3 X
                    const V3 p_i = ric * pci + tic;
                    const V3 p_j = ldm(pq + 21) * p_i + ld3(pq + 30);
                    const double *A = fr + 21 * j + 12, *AR = pq + 12;
                    double Ji[12], Jj[12];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double a0 = r0 * A[c] + r2x * A[6 + c], a1 = r0 * A[3 + c] + r2y * A[6 + c];
                        Ji[c] = a0; Ji[6 + c] = a1; Jj[c] = -a0; Jj[6 + c] = -a1;
                    }
                    {   // Ji rotation part: -reduce * (AR * skew(p_i));  Jj rotation part: reduce * (ric^T * skew(p_j))
                        double X[9], Y[9];
#pragma unroll
                        for (int r = 0; r < 3; r++) {
                            X[3 * r + 0] = -(AR[3 * r + 1] * p_i.z - AR[3 * r + 2] * p_i.y);
                            X[3 * r + 1] = -(AR[3 * r + 2] * p_i.x - AR[3 * r + 0] * p_i.z);
                            X[3 * r + 2] = -(AR[3 * r + 0] * p_i.y - AR[3 * r + 1] * p_i.x);
                            Y[3 * r + 0] = ricT.m[3 * r + 1] * p_j.z - ricT.m[3 * r + 2] * p_j.y;
                            Y[3 * r + 1] = ricT.m[3 * r + 2] * p_j.x - ricT.m[3 * r + 0] * p_j.z;
                            Y[3 * r + 2] = ricT.m[3 * r + 0] * p_j.y - ricT.m[3 * r + 1] * p_j.x;
                        }
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            Ji[3 + c] = r0 * X[c] + r2x * X[6 + c]; Ji[9 + c] = r0 * X[3 + c] + r2y * X[6 + c];
                            Jj[3 + c] = r0 * Y[c] + r2x * Y[6 + c]; Jj[9 + c] = r0 * Y[3 + c] + r2y * Y[6 + c];
                        }
                    }
                    const V3 tl = ldm(pq) * v3(ob.x, ob.y, 1.0);
                    const double kl = -1.0 / (lam * lam);
                    const double Jl0 = kl * (r0 * tl.x + r2x * tl.z), Jl1 = kl * (r0 * tl.y + r2y * tl.z);
                    const double c0r = sr * rx, c1r = sr * ry;
#pragma unroll
                    for (int c = 0; c < 12; c++) { Js[c * T + tid] = Ji[c]; Js[(12 + c) * T + tid] = Jj[c]; }
                    Js[24 * T + tid] = c0r; Js[25 * T + tid] = c1r;
                    double *wl = ws.w + (size_t)l * NPW;
                    double lt8[8];
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        wl[6 * j + a] = Jj[a] * Jl0 + Jj[6 + a] * Jl1;
                        lt8[a] = Ji[a] * Jl0 + Ji[6 + a] * Jl1;
                    }
                    lt8[6] = Jl0 * Jl0 + Jl1 * Jl1; lt8[7] = Jl0 * c0r + Jl1 * c1r;
                    // terms that are sums over the landmark's factors (w_l of the anchor frame, h_ll, g_l): stored per factor in
                    // landmark order and summed by one thread per landmark after the last chunk -- no atomics
                    // a loop-closure factor (observing "frame" = the loop pose, solve frame s.NF) keeps its landmark terms behind the regular ones
                    const int lts = (s.loop_on && j == s.NF) ? nfac_reg + lml[l] : lf0[l] + (j - i - 1);
                    double4 *dst = reinterpret_cast<double4 *>(ws.lt + 8 * (size_t)lts);
                    dst[0] = make_double4(lt8[0], lt8[1], lt8[2], lt8[3]); dst[1] = make_double4(lt8[4], lt8[5], lt8[6], lt8[7]);
                }
            }
            if (!lin) continue;
            __syncthreads();
            BE_PROF_ONLY(if (tid == 0) { const long long t = clock64(); pq2[28] += t - _qt1; _qt1 = t; })
            // H-phase on the FP64 tensor pipe: per frame pair, [Jj^T Ji | Jj^T r], Jj^T Jj and [Ji^T Ji | Ji^T r] are 8x8 (6x7 used)
            // products with the pair's 2 m residual rows as the inner dimension -- one m8n8k4 DMMA per block and two factors.
            // A fragment: lane (g = lane >> 2, k = lane & 3) holds element [dof g][row k]; the B fragment [row k][col g] is the
            // same element, with column 6 = the residual.  Warp w owns pairs w, w + nwarp, ...; with <= 4 pairs per warp the
            // accumulators stay in registers across the chunks and every output element is written once per linearisation.
            for (int round = 0; round * 4 * nwarp < npair; round++) {
#pragma unroll
                for (int ps = 0; ps < 4; ps++) {
                    const int q = (round * 4 + ps) * nwarp + warp;
                    if (q >= npair) continue;
                    int i = 0, rem = q;
                    while (rem >= NF - 1 - i) { rem -= NF - 1 - i; i++; }
                    const int j = i + 1 + rem, key = i * NF + j;
                    const int lo = max(po[key], c0) - c0, hi = min(po[key + 1], c0 + T) - c0;
                    const int g = lane >> 2, k = lane & 3, rr = k & 1;
                    const double *pJi = Js + ((g < 6 ? 6 * rr + g : 24 + rr)) * T, *pJj = Js + (12 + 6 * rr + min(g, 5)) * T;
                    for (int t0 = lo; t0 < hi; t0 += 2) {
                        const int t = t0 + (k >> 1);
                        const bool in = t < hi;
                        const double bi = (in && g < 7) ? pJi[t] : 0.0;          // B: [Ji | r | 0]
                        const double aj = (in && g < 6) ? pJj[t] : 0.0;          // A and B: Jj
                        const double ai = (g < 6) ? bi : 0.0;                    // A: Ji
                        dmma_m8n8k4(hacc[ps][0], hacc[ps][1], aj, bi);
                        dmma_m8n8k4(hacc[ps][2], hacc[ps][3], aj, aj);
                        dmma_m8n8k4(hacc[ps][4], hacc[ps][5], ai, bi);
                    }
                }
                if (!hold) { proj_flush(ws, NP, NF, npair, nwarp, warp, lane, round, hacc); 
#pragma unroll
                    for (int ps = 0; ps < 4; ps++)
#pragma unroll
                        for (int e = 0; e < 6; e++) hacc[ps][e] = 0.0;
                }
            }
            __syncthreads();
            BE_PROF_ONLY(if (tid == 0) { const long long t = clock64(); pq2[29] += t - _qt1; _qt1 = t; })
        }
        if (lin && hold) proj_flush(ws, NP, NF, npair, nwarp, warp, lane, 0, hacc);
        if (lin) {
            const int *anchor = s.lm_anchor + (size_t)b * s.LCAP;
            for (int l = tid; l < nl; l += T) {
                double a8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) a8[k] = 0.0;
                const int f1 = lf0[l + 1];
                for (int f = lf0[l]; f < f1; f++) {
                    const double4 u = *reinterpret_cast<const double4 *>(ws.lt + 8 * (size_t)f), v = *reinterpret_cast<const double4 *>(ws.lt + 8 * (size_t)f + 4);
                    a8[0] += u.x; a8[1] += u.y; a8[2] += u.z; a8[3] += u.w; a8[4] += v.x; a8[5] += v.y; a8[6] += v.z; a8[7] += v.w;
                }
                if (s.loop_on && lml[l] >= 0 && nfac > nfac_reg) {
                    const size_t f = (size_t)nfac_reg + lml[l];
                    const double4 u = *reinterpret_cast<const double4 *>(ws.lt + 8 * f), v = *reinterpret_cast<const double4 *>(ws.lt + 8 * f + 4);
                    a8[0] += u.x; a8[1] += u.y; a8[2] += u.z; a8[3] += u.w; a8[4] += v.x; a8[5] += v.y; a8[6] += v.z; a8[7] += v.w;
                }
                double *wl = ws.w + (size_t)l * NPW + 6 * anchor[l];
#pragma unroll
                for (int k = 0; k < 6; k++) wl[k] = a8[k];
                ws.hll[l] = a8[6]; ws.gl[l] = a8[7];
            }
        }
    }
    __syncthreads();
    if (lin) BE_PROF2(pp, 18);
    const double total = block_sum_d(cost, sh_red);
    __syncthreads();
    if (lin) BE_PROF2(pp, 19);
    return total;
}

// sum_{f,a} w_l[6f+a] * v[15f+a] with the 6*NF entries of the landmark's coupling row spread over the lanes of a warp (coalesced);
// the result is valid in every lane
__device__ __forceinline__ double w_dot_warp(const double *wl, const double *v, int NPW, int lane) {
    double t = 0;
    for (int e = lane; e < NPW; e += 32) { const int f = e / 6; t += wl[e] * v[e + 9 * f]; }
    return warp_sum_d(t);
}

// u^T H u over the full (pose/speed-bias + landmark) system.  Only the LOWER triangle of H is valid (the accumulation writes one
// triangle); elements are visited in memory order, so the reads are coalesced.
__device__ inline double quad_form(const BeState &s, const SolveWs &ws, int nl, const double *up, const double *ul, double *sh_red) {
    const int tid = threadIdx.x, T = blockDim.x, NP = s.NPS, NPW = s.NPWS, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    double acc = 0;
    for (int i = warp; i < NP; i += nwarp) {                         // one warp per row of the lower triangle
        const double *row = ws.H + (size_t)i * NP;
        double t = 0;
#pragma unroll 2
        for (int j = lane; j < i; j += 32) t += row[j] * up[j];
        acc += 2.0 * up[i] * t;
        if (lane == 0) acc += row[i] * up[i] * up[i];
    }
    for (int l = warp; l < nl; l += nwarp) {                         // one warp per landmark row of the coupling block
        const double t = w_dot_warp(ws.w + (size_t)l * NPW, up, NPW, lane);
        if (lane == 0) acc += 2.0 * ul[l] * t + ws.hll[l] * ul[l] * ul[l];
    }
    return block_sum_d(acc, sh_red);
}

__device__ inline double block_max_d(double v, double *smem /*>=32*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = smem[0];
    for (int i = 1; i < nw; i++) r = fmax(r, smem[i]);
    return r;
}

__device__ inline double dot2(const double *ap, const double *al, const double *bp, const double *bl, int NP, int nl, double *sh_red) {
    double acc = 0;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) acc += ap[i] * bp[i];
    for (int l = threadIdx.x; l < nl; l += blockDim.x) acc += al[l] * bl[l];
    return block_sum_d(acc, sh_red);
}

// In-place lower Cholesky of the n x n matrix A (ld = n, only the lower triangle is referenced) + solve A y = rhs.
// Returns false (in all threads) on a non-positive pivot / non-finite value (Eigen LLT info() != Success).
__device__ inline bool chol_solve(double *A, int n, const double *rhs, double *y, int *sh_flag) {
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) *sh_flag = 1;
    __syncthreads();
    for (int k = 0; k < n; k++) {
        if (tid == 0) { const double d = A[(size_t)k * n + k]; if (!(d > 0) || !isfinite(d)) *sh_flag = 0; else A[(size_t)k * n + k] = sqrt(d); }
        __syncthreads();
        if (!*sh_flag) return false;
        const double piv = A[(size_t)k * n + k];
        for (int i = k + 1 + tid; i < n; i += T) A[(size_t)i * n + k] /= piv;
        __syncthreads();
        const int r = n - k - 1;
        for (int e = tid; e < r * r; e += T) {
            const int ii = e / r, jj = e - ii * r;
            if (jj <= ii) { const int i = k + 1 + ii, j = k + 1 + jj; A[(size_t)i * n + j] -= A[(size_t)i * n + k] * A[(size_t)j * n + k]; }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += T) y[i] = rhs[i];
    __syncthreads();
    for (int k = 0; k < n; k++) {                          // L z = rhs
        if (tid == 0) y[k] /= A[(size_t)k * n + k];
        __syncthreads();
        const double yk = y[k];
        for (int i = k + 1 + tid; i < n; i += T) y[i] -= A[(size_t)i * n + k] * yk;
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; k--) {                     // L^T y = z
        if (tid == 0) y[k] /= A[(size_t)k * n + k];
        __syncthreads();
        const double yk = y[k];
        for (int i = tid; i < k; i += T) y[i] -= A[(size_t)k * n + i] * yk;
        __syncthreads();
    }
    bool ok = true;
    for (int i = 0; i < n; i++) ok &= isfinite(y[i]);
    return ok;
}


// ---- shared-memory path (reduced system fits one SM: NP*(NP+1)/2 doubles, 110 KB at W=10) -------------------------------------
__device__ __forceinline__ int pidx(int i, int j) { return i * (i + 1) / 2 + j; }       // packed lower, j <= i
__device__ __forceinline__ int tri_row(int e) {             // largest i with i (i + 1) / 2 <= e
    int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    while (i * (i + 1) / 2 > e) i--;
    return i;
}

// 1 / sqrt(x) for a positive, finite, normal-range x: single-precision seed + two Newton steps + one to absorb the seed's rounding (no MUFU.RSQ64H fix-up branches on
// the critical path of the factorisation); relative error ~1e-16
__device__ __forceinline__ double fast_rsqrt(double x) {
    if (!(x > 1e-30 && x < 1e30)) return rsqrt(x);
    double y = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}

// In-register Cholesky of one diagonal block (at most 8 x 8, identity-padded) by a single thread: 8 pivots, 28 scalings, 84 FMAs.
// Writes L over the block, 1 / L_kk to dinv; returns false on a non-positive / non-finite pivot.
__device__ __forceinline__ bool chol_diag8(double *A, int c0, int nb, double *dinv) {
    double a[36];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) a[r * (r + 1) / 2 + c] = (r < nb) ? A[pidx(c0 + r, c0 + c)] : (r == c ? 1.0 : 0.0);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double piv = a[k * (k + 1) / 2 + k];
        ok &= (piv > 0) && isfinite(piv);
        const double il = fast_rsqrt(piv);
        a[k * (k + 1) / 2 + k] = piv * il;
        if (k < nb) dinv[c0 + k] = il;
#pragma unroll
        for (int r = k + 1; r < 8; r++) a[r * (r + 1) / 2 + k] *= il;
#pragma unroll
        for (int r = k + 1; r < 8; r++)
#pragma unroll
            for (int c = k + 1; c <= r; c++) a[r * (r + 1) / 2 + c] -= a[r * (r + 1) / 2 + k] * a[c * (c + 1) / 2 + k];
    }
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) if (r < nb) A[pidx(c0 + r, c0 + c)] = a[r * (r + 1) / 2 + c];
    return ok;
}

// L^T y = z in place (z in shared memory), L packed lower in shared memory, dinv = 1 / diag(L).  Blocked 16 wide: warp 0 back-solves
// a diagonal block with its columns held in registers, then every thread folds the block's solution into the rows above.
__device__ inline void chol_backward_packed(const double *A, int n, double *z, const double *dinv) {
    constexpr int BB = 16;
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    for (int c0 = ((n - 1) / BB) * BB; c0 >= 0; c0 -= BB) {
        const int nb = min(BB, n - c0);
        if (tid < 32) {
            double col[BB];                                          // col[t] = L[c0 + t][c0 + lane], t > lane
#pragma unroll
            for (int t = 0; t < BB; t++) col[t] = (t > lane && t < nb) ? A[pidx(c0 + t, c0 + lane)] : 0.0;
            double yv = (lane < nb) ? z[c0 + lane] : 0.0;
            const double di = (lane < nb) ? dinv[c0 + lane] : 0.0;
#pragma unroll
            for (int j = BB - 1; j >= 0; j--) {
                if (lane == j) yv *= di;
                const double yj = __shfl_sync(FULL, yv, j);
                if (lane < j) yv -= col[j] * yj;
            }
            if (lane < nb) z[c0 + lane] = yv;
        }
        __syncthreads();
        for (int i = tid; i < c0; i += T) {
            double v = z[i];
#pragma unroll
            for (int t = 0; t < BB; t++) if (t < nb) v -= A[pidx(c0 + t, i)] * z[c0 + t];
            z[i] = v;
        }
        __syncthreads();
    }
}

// L y = z in place, same conventions (used to re-solve with an existing factor)
__device__ inline void chol_forward_packed(const double *A, int n, double *z, const double *dinv) {
    constexpr int BB = 16;
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    for (int c0 = 0; c0 < n; c0 += BB) {
        const int nb = min(BB, n - c0);
        if (tid < 32) {
            double row[BB];                                          // row[t] = L[c0 + lane][c0 + t], t < lane
#pragma unroll
            for (int t = 0; t < BB; t++) row[t] = (t < lane && lane < nb) ? A[pidx(c0 + lane, c0 + t)] : 0.0;
            double yv = (lane < nb) ? z[c0 + lane] : 0.0;
            const double di = (lane < nb) ? dinv[c0 + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < BB; j++) {
                if (lane == j) yv *= di;
                const double yj = __shfl_sync(FULL, yv, j);
                if (lane > j) yv -= row[j] * yj;
            }
            if (lane < nb) z[c0 + lane] = yv;
        }
        __syncthreads();
        for (int i = c0 + nb + tid; i < n; i += T) {
            double v = z[i];
            const double *ri = A + pidx(i, c0);
#pragma unroll
            for (int t = 0; t < BB; t++) if (t < nb) v -= ri[t] * z[c0 + t];
            z[i] = v;
        }
        __syncthreads();
    }
}

// Blocked RIGHT-looking Cholesky on a packed lower matrix in shared memory, BORDERED by the right-hand side (row n of the packed
// array holds rhs, so the forward substitution z = L^-1 rhs is a by-product of the factorisation).  Per 8-column panel:
//   (b) one thread per row below the diagonal block solves its 8 entries against the block (block reads are warp broadcasts);
//   (c) the trailing matrix gets the rank-8 update in 4x4 register tiles (0.5 shared loads per FMA) from warps 1..; meanwhile
//   (a) warp 0 LOOKS AHEAD: its lanes update the next diagonal block, then lane 0 factors it in registers (chol_diag8) -- the
//       sequential pivot chain (8 x rsqrt + scaling + update) overlaps the trailing update instead of adding to it.
// Two barriers per panel.  Backward substitution L^T y = z is blocked 16 wide: warp 0 back-solves a diagonal block with its columns
// held in registers, then every thread folds the block's solution into the rows above.
// A must hold (n+1)(n+2)/2 doubles, dinv n doubles (shared).  Returns false (in all threads) on a non-positive pivot / non-finite
// value (Eigen LLT info() != Success).
__device__ __noinline__ bool chol_solve_packed(double *A, int n, const double *rhs, double *y, int *sh_flag, double *dinv, long long *pp) {
    BE_PROF2_INIT;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    constexpr int NB = 8;
    if (tid == 0) *sh_flag = 1;
    for (int j = tid; j < n; j += T) A[pidx(n, j)] = rhs[j];
    __syncthreads();
    if (tid == 0 && !chol_diag8(A, 0, min(NB, n), dinv)) *sh_flag = 0;
    __syncthreads();
    BE_PROF2(pp, 21);
    for (int c0 = 0; c0 < n; c0 += NB) {
        const int nb = min(NB, n - c0);
        if (!*sh_flag) return false;
        for (int i = c0 + nb + tid; i <= n; i += T) {                // (b) rows below the block (and the rhs row)
            double *ri = A + pidx(i, c0);
            double v[NB];
#pragma unroll
            for (int jj = 0; jj < NB; jj++) {
                if (jj < nb) {
                    double w = ri[jj];
                    const double *rj = A + pidx(c0 + jj, c0);
#pragma unroll
                    for (int t = 0; t < jj; t++) w -= v[t] * rj[t];
                    v[jj] = w * dinv[c0 + jj];
                    ri[jj] = v[jj];
                }
            }
        }
        __syncthreads();
        BE_PROF2(pp, 22);
        const int r0 = c0 + nb;                                      // first trailing row / column
        const int nb2 = min(NB, n - r0);                             // size of the next diagonal block (<= 0: none)
        if (tid < 32) {                                              // (a) look-ahead on the next diagonal block
            if (nb2 > 0) {
                for (int e = lane; e < 36; e += 32) {
                    const int r = tri_row(e), c = e - r * (r + 1) / 2;
                    if (r < nb2) {
                        const double *pr_ = A + pidx(r0 + r, c0), *pc_ = A + pidx(r0 + c, c0);
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < NB; t++) if (t < nb) acc += pr_[t] * pc_[t];
                        A[pidx(r0 + r, r0 + c)] -= acc;
                    }
                }
                __syncwarp();
                if (lane == 0 && !chol_diag8(A, r0, nb2, dinv)) *sh_flag = 0;
            }
        } else {                                                     // (c) trailing update, 4x4 register tiles
            const int R = n + 1 - r0;                                // rows r0 .. n (the rhs row included)
            const int RT = (R + 3) >> 2;
            const int ntile = RT * (RT + 1) / 2;
            const int skip = nb2 > 0 ? r0 + nb2 : r0;                // rows below `skip` belong to the look-ahead block
            for (int e = tid - 32; e < ntile; e += T - 32) {
                const int ti = tri_row(e), tj = e - ti * (ti + 1) / 2;
                const int I = r0 + 4 * ti, J = r0 + 4 * tj;
                if (I + 3 < skip) continue;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
                const double *pi_[4], *pj_[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { pi_[a] = A + pidx(min(I + a, n), c0); pj_[a] = A + pidx(min(J + a, n), c0); }
#pragma unroll
                for (int t = 0; t < NB; t++) {
                    if (t < nb) {
                        double li[4], lj[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) { li[a] = pi_[a][t]; lj[a] = pj_[a][t]; }
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int c = 0; c < 4; c++) acc[a][c] += li[a] * lj[c];
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int i = I + a, j = J + c;
                        if (i <= n && i >= skip && j <= i && j < n) A[pidx(i, j)] -= acc[a][c];
                    }
            }
        }
        __syncthreads();
        BE_PROF2(pp, 20);
    }
    double *z = A + pidx(n, 0);
    chol_backward_packed(A, n, z, dinv);
    bool ok = true;
    for (int i = tid; i < n; i += T) { const double v = z[i]; y[i] = v; ok &= isfinite(v); }
    BE_PROF2(pp, 23);
    return __syncthreads_and(ok) != 0;
}

// S (packed, shared) = S H S + mu D^2 - sum_l ws_l ws_l^T / h_l ;  rhs = S g - sum_l ws_l gs_l / h_l   (landmark blocks eliminated)
// The landmark sum is a SYRK over the (NPW+1)-vectors v_l = [w_l ; g_l] * sqrt(s_l^2 / h_l) (the border row yields the rhs
// correction).  Chunks of SCHUR_CHUNK landmarks are staged in shared memory (wt = [SCHUR_CHUNK][SCHUR_LD]); a thread owns one 4x4
// tile of the lower triangle (two 16-byte shared loads per operand per landmark, 0.25 loads per FMA) and the landmarks of a chunk
// are split over G = blockDim / ntile thread groups whose partial tiles are applied one group after the other (fixed order).
constexpr int SCHUR_CHUNK = 64;
__host__ __device__ inline int schur_ld(int NPW) { return (NPW + 1 + 3) & ~3; }
__host__ __device__ inline int schur_ntile(int NPW) { const int rt = schur_ld(NPW) / 4; return rt * (rt + 1) / 2; }
__device__ __noinline__ void build_reduced_smem(const BeState &s, const SolveWs &ws, int nl, double mu, double *S, double *wt) {
    const int tid = threadIdx.x, T = blockDim.x, NP = s.NPS, NPW = s.NPWS;
    for (int e = tid; e < NP * (NP + 1) / 2; e += T) {
        const int i = tri_row(e), j = e - pidx(i, 0);
        double v = ws.H[(size_t)i * NP + j] * ws.sc_p[i] * ws.sc_p[j];
        if (i == j) v += mu * ws.d_p[i] * ws.d_p[i];
        S[e] = v;
    }
    for (int i = tid; i < NP; i += T) ws.rhs[i] = ws.g[i] * ws.sc_p[i];
    for (int l = tid; l < nl; l += T) {                             // per-landmark weight sqrt(s_l^2 / h_l) (u_l is free scratch here)
        const double sl = ws.sc_l[l];
        ws.u_l[l] = sqrt(sl * sl / (ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l]));
    }
    const int ld = schur_ld(NPW), ntile = schur_ntile(NPW), G = T / ntile;          // host guarantees G >= 1
    const int grp = tid / ntile, tile = tid - grp * ntile;
    const bool active = grp < G;
    int I = 0, J = 0;
    if (active) { const int ti = tri_row(tile); I = 4 * ti; J = 4 * (tile - ti * (ti + 1) / 2); }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
    for (int l0 = 0; l0 < nl; l0 += SCHUR_CHUNK) {
        const int cn = min(SCHUR_CHUNK, nl - l0);
        __syncthreads();
        for (int e = tid; e < cn * ld; e += T) {
            const int cl = e / ld, a = e - cl * ld, l = l0 + cl;
            wt[e] = (a < NPW) ? ws.w[(size_t)l * NPW + a] * ws.u_l[l] : (a == NPW ? ws.gl[l] * ws.u_l[l] : 0.0);
        }
        __syncthreads();
        if (active)
            for (int cl = grp; cl < cn; cl += G) {
                const double *row = wt + cl * ld;
                const double2 a01 = *reinterpret_cast<const double2 *>(row + I), a23 = *reinterpret_cast<const double2 *>(row + I + 2);
                const double2 b01 = *reinterpret_cast<const double2 *>(row + J), b23 = *reinterpret_cast<const double2 *>(row + J + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] += av[a] * bv[c];
            }
    }
    for (int g = 0; g < G; g++) {
        __syncthreads();
        if (active && grp == g) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int ra = I + a, rc = J + c;
                    if (rc > ra || ra > NPW || rc >= NPW) continue;
                    const int ic = 15 * (rc / 6) + rc % 6;
                    if (ra == NPW) ws.rhs[ic] -= acc[a][c] * ws.sc_p[ic];
                    else { const int ia = 15 * (ra / 6) + ra % 6; S[pidx(ia, ic)] -= acc[a][c] * ws.sc_p[ia] * ws.sc_p[ic]; }
                }
        }
    }
    __syncthreads();
}

// ---- global-memory path (reduced system too large for one SM: W = 20 -> 315 x 315) -----------------------------------------------
// Same mathematics as build_reduced_smem / chol_solve_packed with the matrix in the stream's global scratch (L2-resident: 0.8 MB) and
// only the hot operands in shared memory: the landmark chunk of the Schur SYRK, and the current panel of the blocked factorisation.
__device__ __noinline__ void build_reduced_global(const BeState &s, const SolveWs &ws, int nl, double mu, double *wt /* smem [SCHUR_CHUNK][ld] */) {
    const int tid = threadIdx.x, T = blockDim.x, NP = s.NPS, NPW = s.NPWS;
    for (int e = tid; e < NP * NP; e += T) {
        const int i = e / NP, j = e - i * NP;
        if (j <= i) {
            double v = ws.H[e] * ws.sc_p[i] * ws.sc_p[j];
            if (i == j) v += mu * ws.d_p[i] * ws.d_p[i];
            ws.S[e] = v;
        }
    }
    for (int i = tid; i < NP; i += T) ws.rhs[i] = ws.g[i] * ws.sc_p[i];
    for (int l = tid; l < nl; l += T) {                             // per-landmark weight sqrt(s_l^2 / h_l) (u_l is free scratch here)
        const double sl = ws.sc_l[l];
        ws.u_l[l] = sqrt(sl * sl / (ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l]));
    }
    const int ld = schur_ld(NPW), rt = ld / 4, ntile = rt * (rt + 1) / 2;
    // a thread owns up to TPT 4x4 tiles of the lower triangle of the (NPW+1) x (NPW+1) SYRK (border row = right-hand side)
    constexpr int TPT = 2;
    int I[TPT], J[TPT];
    double acc[TPT][4][4];
#pragma unroll
    for (int q = 0; q < TPT; q++) {
        const int tile = tid + q * T;
        if (tile < ntile) { const int ti = tri_row(tile); I[q] = 4 * ti; J[q] = 4 * (tile - ti * (ti + 1) / 2); } else { I[q] = -1; J[q] = 0; }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[q][a][c] = 0.0;
    }
    for (int l0 = 0; l0 < nl; l0 += SCHUR_CHUNK) {
        const int cn = min(SCHUR_CHUNK, nl - l0);
        __syncthreads();
        for (int e = tid; e < cn * ld; e += T) {
            const int cl = e / ld, a = e - cl * ld, l = l0 + cl;
            wt[e] = (a < NPW) ? ws.w[(size_t)l * NPW + a] * ws.u_l[l] : (a == NPW ? ws.gl[l] * ws.u_l[l] : 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TPT; q++) {
            if (I[q] < 0) continue;
            for (int cl = 0; cl < cn; cl++) {
                const double *row = wt + cl * ld;
                const double2 a01 = *reinterpret_cast<const double2 *>(row + I[q]), a23 = *reinterpret_cast<const double2 *>(row + I[q] + 2);
                const double2 b01 = *reinterpret_cast<const double2 *>(row + J[q]), b23 = *reinterpret_cast<const double2 *>(row + J[q] + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[q][a][c] += av[a] * bv[c];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TPT; q++) {
        if (I[q] < 0) continue;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int ra = I[q] + a, rc = J[q] + c;
                if (rc > ra || ra > NPW || rc >= NPW) continue;
                const int ic = 15 * (rc / 6) + rc % 6;
                if (ra == NPW) ws.rhs[ic] -= acc[q][a][c] * ws.sc_p[ic];          // one owner per element: no atomics
                else { const int ia = 15 * (ra / 6) + ra % 6; ws.S[(size_t)ia * NP + ic] -= acc[q][a][c] * ws.sc_p[ia] * ws.sc_p[ic]; }
            }
    }
    __syncthreads();
}

// Blocked right-looking Cholesky of the lower triangle of A (global, ld = n) with 16-column panels + the two triangular solves.
//   per panel: (a) warp 0 factors the 16 x 16 diagonal block in shared memory; (b) one thread per row below solves its 16 entries
//   against the block and leaves the row in the shared panel; (c) the trailing matrix gets the rank-16 update in 4x4 register tiles
//   whose operands come from the shared panel.  smem: panel [n][16] + diag [16][17] + dinv [n] doubles.
constexpr int GB_NB = 16;
__host__ __device__ inline size_t chol_global_smem_doubles(int n) { return (size_t)n * GB_NB + GB_NB * (GB_NB + 1) + n + 8; }
__device__ __noinline__ bool chol_solve_blocked_global(double *A, int n, const double *rhs, double *y, int *sh_flag, double *sm) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    constexpr int NB = GB_NB, DL = GB_NB + 1;
    double *panel = sm, *diag = panel + (size_t)n * NB, *dinv = diag + NB * DL;
    if (tid == 0) *sh_flag = 1;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += NB) {
        const int nb = min(NB, n - c0);
        // (a) diagonal block -> shared, factored by warp 0 (column by column, lanes = rows)
        for (int e = tid; e < NB * NB; e += T) { const int r = e / NB, c = e - r * NB; diag[r * DL + c] = (r < nb && c <= r) ? A[(size_t)(c0 + r) * n + c0 + c] : (r == c ? 1.0 : 0.0); }
        __syncthreads();
        if (tid < 32) {
            for (int k = 0; k < nb; k++) {
                const double piv = diag[k * DL + k];
                if (lane == 0) { if (!(piv > 0) || !isfinite(piv)) *sh_flag = 0; }
                const double il = fast_rsqrt(piv);
                __syncwarp();
                if (lane == k) { diag[k * DL + k] = piv * il; dinv[c0 + k] = il; }
                if (lane > k && lane < nb) diag[lane * DL + k] *= il;
                __syncwarp();
                for (int e = lane; e < (nb - k - 1) * (nb - k - 1); e += 32) {
                    const int r = k + 1 + e / (nb - k - 1), c = k + 1 + e % (nb - k - 1);
                    if (c <= r) diag[r * DL + c] -= diag[r * DL + k] * diag[c * DL + k];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (!*sh_flag) return false;
        for (int e = tid; e < nb * nb; e += T) { const int r = e / nb, c = e - r * nb; if (c <= r) A[(size_t)(c0 + r) * n + c0 + c] = diag[r * DL + c]; }
        // (b) rows below the block
        const int r0 = c0 + nb;
        for (int i = r0 + tid; i < n; i += T) {
            double *ri = A + (size_t)i * n + c0;
            double v[NB];
#pragma unroll
            for (int jj = 0; jj < NB; jj++) {
                if (jj < nb) {
                    double w = ri[jj];
#pragma unroll
                    for (int t = 0; t < jj; t++) w -= v[t] * diag[jj * DL + t];
                    v[jj] = w * dinv[c0 + jj];
                    ri[jj] = v[jj];
                    panel[(size_t)(i - r0) * NB + jj] = v[jj];
                } else panel[(size_t)(i - r0) * NB + jj] = 0.0;
            }
        }
        __syncthreads();
        // (c) trailing update A[i][j] -= sum_t P[i][t] P[j][t], i >= j >= r0
        const int R = n - r0, RT = (R + 3) >> 2, ntile = RT * (RT + 1) / 2;
        for (int e = tid; e < ntile; e += T) {
            const int ti = tri_row(e), tj = e - ti * (ti + 1) / 2;
            const int I = 4 * ti, J = 4 * tj;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
            const double *pi_[4], *pj_[4];
#pragma unroll
            for (int a = 0; a < 4; a++) { pi_[a] = panel + (size_t)min(I + a, R - 1) * NB; pj_[a] = panel + (size_t)min(J + a, R - 1) * NB; }
#pragma unroll
            for (int t = 0; t < NB; t += 2) {
                double2 li[4], lj[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { li[a] = *reinterpret_cast<const double2 *>(pi_[a] + t); lj[a] = *reinterpret_cast<const double2 *>(pj_[a] + t); }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] += li[a].x * lj[c].x + li[a].y * lj[c].y;
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int i = I + a, j = J + c;
                    if (i < R && j <= i) A[(size_t)(r0 + i) * n + r0 + j] -= acc[a][c];
                }
        }
        __syncthreads();
    }
    // L z = rhs, L^T y = z (column sweeps; dinv = 1 / diag(L))
    for (int i = tid; i < n; i += T) y[i] = rhs[i];
    __syncthreads();
    for (int k = 0; k < n; k++) {
        if (tid == 0) y[k] *= dinv[k];
        __syncthreads();
        const double yk = y[k];
        for (int i = k + 1 + tid; i < n; i += T) y[i] -= A[(size_t)i * n + k] * yk;
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; k--) {
        if (tid == 0) y[k] *= dinv[k];
        __syncthreads();
        const double yk = y[k];
        for (int i = tid; i < k; i += T) y[i] -= A[(size_t)k * n + i] * yk;
        __syncthreads();
    }
    bool ok = true;
    for (int i = tid; i < n; i += T) ok &= isfinite(y[i]);
    return __syncthreads_and(ok) != 0;
}

}  // namespace be
#include "be_tilechol.cuh"
namespace be {

// shared scratch evaluate() needs: IMU J buffers, then (aliased) frame / pair tables + the J store of one chunk of SOLVE_T factors
__host__ __device__ inline size_t eval_smem_bytes(int W) {
    const size_t NF = W + 1, npair = NF * (NF - 1) / 2;
    const size_t a = (size_t)930 * W, c = 21 * NF + 33 * npair + 1 + (size_t)26 * SOLVE_T;
    return (a > c ? a : c) * sizeof(double);
}
__host__ __device__ inline size_t solve_smem_bytes(int NP, int NPW) {
    return ((size_t)(NP + 1) * (NP + 2) / 2 + 10 + (size_t)SCHUR_CHUNK * schur_ld(NPW)) * sizeof(double);
}

__global__ void __launch_bounds__(SOLVE_T) solve_kernel(BeState s, int use_smem, int vec_off) {
    VIO_POISON(128u);
    extern __shared__ __align__(16) double sm_dyn[];
    const int T = blockDim.x;                              // 256 or 512 (VIO_BE_THREADS)
    __shared__ double sh_red[32];
    __shared__ int sh_flag;
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    double *dvs = S_dv(s, b);
    const int NP = s.NPS, NPW = s.NPWS, NF = s.NFS, nl = iv[IV_N_LM];
    const SolveWs ws = carve(s, b, sm_dyn + vec_off, nl);
    double *par = s.par + (size_t)b * s.par_stride;
    double *cand = s.cand + (size_t)b * s.par_stride;

    BE_PROF_INIT;
    double x_cost = evaluate(s, b, ws, par, 2, sh_red, sm_dyn);
    BE_PROF(0);
    if (tid == 0) dvs[DV_COST0] = x_cost;
    // Jacobi scaling, frozen at iteration 0 (trust_region_minimizer.cc:239-254)
    for (int i = tid; i < NP; i += T) ws.sc_p[i] = 1.0 / (1.0 + sqrt(ws.H[(size_t)i * NP + i]));
    for (int l = tid; l < nl; l += T) ws.sc_l[l] = 1.0 / (1.0 + sqrt(ws.hll[l]));
    __syncthreads();

    double radius = 1e4, mu = 1e-8, alpha = 0.0, dogleg_norm = 0.0, x_norm = -1.0, mu_gn = 0.0, g2_lin = 0.0, gHg_lin = 0.0;
    bool reuse = false;
    // The candidate is evaluated WITH its linearisation (speculatively): an accepted step -- the usual case -- then needs no second
    // pass over the factors (Ceres evaluates the candidate cost-only and re-evaluates with Jacobians after accepting; same numbers).
    // After a rejected candidate H / g / w belong to the candidate; lin_valid says whether they belong to the current point.
    bool lin_valid = true;
    int iter = 0, invalid_run = 0;
    bool step_ok = true;                                   // iteration 0 counts as successful
    while (true) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue()
        if (iter >= s.max_iters) break;
        if (step_ok) {
            double gm = 0;                                 // gradient_max_norm ~ max |g|
            for (int i = tid; i < NP; i += T) gm = fmax(gm, fabs(ws.g[i]));
            for (int l = tid; l < nl; l += T) gm = fmax(gm, fabs(ws.gl[l]));
            const double m = block_max_d(gm, sh_red);
            if (m <= 1e-10) break;
        }
        if (radius <= 1e-32) break;
        iter++;
        // ---- DoglegStrategy::ComputeStep ------------------------------------------------------------------
        bool linear_ok = true;
        if (!reuse) {
            reuse = true;
            if (!lin_valid) { (void)evaluate(s, b, ws, par, 1, sh_red, sm_dyn); lin_valid = true; }     // rejected candidate, then an invalid step
            for (int i = tid; i < NP; i += T) {
                const double sc = ws.sc_p[i];
                const double d = sqrt(fmin(fmax(ws.H[(size_t)i * NP + i] * sc * sc, 1e-6), 1e32));
                ws.d_p[i] = d;
                ws.gr_p[i] = ws.g[i] * sc / d;                       // gradient_ = (JS)^T r ./ diagonal
                ws.u_p[i] = sc * (ws.g[i] * sc / (d * d));           // S * (gradient_ ./ diagonal)
            }
            for (int l = tid; l < nl; l += T) {
                const double sc = ws.sc_l[l];
                const double d = sqrt(fmin(fmax(ws.hll[l] * sc * sc, 1e-6), 1e32));
                ws.d_l[l] = d;
                ws.gr_l[l] = ws.gl[l] * sc / d;
                ws.u_l[l] = sc * (ws.gl[l] * sc / (d * d));
            }
            __syncthreads();
            const double g2 = dot2(ws.gr_p, ws.gr_l, ws.gr_p, ws.gr_l, NP, nl, sh_red);
            const double jg2 = quad_form(s, ws, nl, ws.u_p, ws.u_l, sh_red);
            alpha = g2 / jg2;                                         // ComputeCauchyPoint
            g2_lin = g2; gHg_lin = jg2;
            BE_PROF(1);
            // ---- ComputeGaussNewtonStep: (S H S + mu D^2) y = S g by Schur elimination of the landmark blocks ----
            linear_ok = false;
            while (mu < 1.0) {
              bool ok;
              if (use_smem == 2) {
                ok = tile_reduced_solve(s, ws.H, ws.g, ws.w, ws.hll, ws.gl, ws.sc_p, ws.sc_l, ws.d_p, ws.d_l, ws.u_l, nl, mu, ws.y, sm_dyn, &sh_flag,
                                        s.prof + (size_t)b * 32);
                BE_PROF(3);
              } else if (use_smem) {
                double *Ssm = sm_dyn, *wt = sm_dyn + ((((size_t)(NP + 1) * (NP + 2) / 2 + 8) + 1) & ~(size_t)1);   // 16-byte aligned
                build_reduced_smem(s, ws, nl, mu, Ssm, wt);
                BE_PROF(2);
                ok = chol_solve_packed(Ssm, NP, ws.rhs, ws.y, &sh_flag, wt, s.prof + (size_t)b * 32);
                BE_PROF(3);
              } else {
                // reduced system in global memory (W = 20): tiled Schur SYRK + blocked Cholesky with shared panels
                double *wt = sm_dyn;
                build_reduced_global(s, ws, nl, mu, wt);
                ok = chol_solve_blocked_global(ws.S, NP, ws.rhs, ws.y, &sh_flag, sm_dyn);
              }
                __syncthreads();
                if (ok) {
                    // back-substitution y_l = (gs_l - ws_l . y_p) / h_l ;  gauss_newton_step_ = -diagonal .* y
                    for (int i = tid; i < NP; i += T) { ws.gn_p[i] = -ws.d_p[i] * ws.y[i]; ws.rhs[i] = ws.sc_p[i] * ws.y[i]; }   // rhs: free scratch
                    __syncthreads();
                    for (int l = tid >> 5; l < nl; l += T >> 5) {
                        const double t = w_dot_warp(ws.w + (size_t)l * NPW, ws.rhs, NPW, tid & 31);
                        if ((tid & 31) == 0) {
                            const double sl = ws.sc_l[l];
                            const double h = ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l];
                            const double yl = (ws.gl[l] * sl - sl * t) / h;
                            ws.gn_l[l] = -ws.d_l[l] * yl;
                        }
                    }
                    __syncthreads();
                    mu_gn = mu;
                    linear_ok = true;
                    break;
                }
                mu *= 10.0;
                if (tid == 0) iv[IV_CHOL_RETRY] += 1;
            }
        }
        bool valid = false;
        double model_change = 0.0;
        if (linear_ok) {
            // ---- ComputeTraditionalDoglegStep (dogleg_strategy.cc:199-255) ------------------------------------
            // In the dogleg coordinates (scaled by S and D): gradient g^ = gr, model Hessian M = D^-1 S H S D^-1, and the Gauss-Newton
            // step satisfies (M + mu I) gn = -g^ exactly as solved above.
            const double g_norm = sqrt(g2_lin);
            const double gn2 = dot2(ws.gn_p, ws.gn_l, ws.gn_p, ws.gn_l, NP, nl, sh_red);
            const double gngr = dot2(ws.gr_p, ws.gr_l, ws.gn_p, ws.gn_l, NP, nl, sh_red);
            const double gn_norm = sqrt(gn2);
            double ca, cb;                                            // step = ca * gradient_ + cb * gauss_newton_step_
            if (gn_norm <= radius) { ca = 0; cb = 1; dogleg_norm = gn_norm; }
            else if (g_norm * alpha >= radius) { ca = -(radius / g_norm); cb = 0; dogleg_norm = radius; }
            else {
                const double b_dot_a = -alpha * gngr;
                const double a2 = (alpha * g_norm) * (alpha * g_norm);
                const double bma2 = a2 - 2 * b_dot_a + gn_norm * gn_norm;
                const double c = b_dot_a - a2;
                const double d = sqrt(c * c + bma2 * (radius * radius - a2));
                const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
                ca = -alpha * (1.0 - beta); cb = beta;
                dogleg_norm = -1.0;
            }
            for (int i = tid; i < NP; i += T) {
                const double st = ca * ws.gr_p[i] + cb * ws.gn_p[i];
                ws.st_p[i] = st;
                ws.u_p[i] = st / ws.d_p[i] * ws.sc_p[i];              // trust_region_step_ = dogleg ./ diagonal ; delta = .* jacobian_scaling_
            }
            for (int l = tid; l < nl; l += T) {
                const double st = ca * ws.gr_l[l] + cb * ws.gn_l[l];
                ws.st_l[l] = st;
                ws.u_l[l] = st / ws.d_l[l] * ws.sc_l[l];
            }
            __syncthreads();
            if (dogleg_norm < 0) dogleg_norm = sqrt(dot2(ws.st_p, ws.st_l, ws.st_p, ws.st_l, NP, nl, sh_red));
            // model_cost_change = -(J d)^T (r + J d / 2) = -(d^.g^ + d^^T M d^ / 2), with M gn = -g^ - mu gn and g^T M g^ from the
            // Cauchy-point computation: no pass over H
            const double dg = ca * g2_lin + cb * gngr;
            const double dMd = ca * ca * gHg_lin + 2.0 * ca * cb * (-g2_lin - mu_gn * gngr) + cb * cb * (-gngr - mu_gn * gn2);
            model_change = -(dg + 0.5 * dMd);
            valid = model_change > 0.0;
            BE_PROF(4);
        }
        if (!valid) {                                                 // HandleInvalidStep()
            if (++invalid_run >= 5) break;
            mu *= 10.0; reuse = false; step_ok = false;
            continue;
        }
        invalid_run = 0;
        // ---- candidate = Plus(x, delta), cost-only evaluation ---------------------------------------------
        for (int i = tid; i < NF; i += T) {
            pose_plus(par + 16 * i, ws.u_p + 15 * i, cand + 16 * i);
            for (int k = 0; k < 9; k++) cand[16 * i + 7 + k] = par[16 * i + 7 + k] + ws.u_p[15 * i + 6 + k];
        }
        for (int l = tid; l < nl; l += T) cand[16 * NF + l] = par[16 * NF + l] + ws.u_l[l];
        __syncthreads();
        const double cand_cost = evaluate(s, b, ws, cand, 1, sh_red, sm_dyn);
        lin_valid = false;
        BE_PROF(5);
        // ParameterToleranceReached / FunctionToleranceReached (trust_region_minimizer.cc:662-705)
        double sn = 0;
        for (int i = tid; i < 16 * NF + nl; i += T) { const double d = par[i] - cand[i]; sn += d * d; }
        const double step_norm = sqrt(block_sum_d(sn, sh_red));
        if (step_norm <= 1e-8 * (x_norm + 1e-8)) break;
        if (fabs(x_cost - cand_cost) <= 1e-6 * x_cost) break;
        const double quality = (x_cost - cand_cost) / model_change;   // StepQuality with max_consecutive_nonmonotonic_steps = 0
        if (quality > 1e-3) {                                         // HandleSuccessfulStep()
            for (int i = tid; i < 16 * NF + nl; i += T) par[i] = cand[i];
            __syncthreads();
            double xn = 0;
            for (int i = tid; i < 16 * NF + nl; i += T) xn += par[i] * par[i];
            x_norm = sqrt(block_sum_d(xn, sh_red));
            BE_PROF(6);
            x_cost = cand_cost;                                       // H, g, w already hold the linearisation at the accepted point
            lin_valid = true;
            step_ok = true;
            if (quality < 0.25) radius *= 0.5;                        // DoglegStrategy::StepAccepted
            if (quality > 0.75) radius = fmax(radius, 3.0 * dogleg_norm);
            mu = fmax(1e-8, 2.0 * mu / 10.0);
            reuse = false;
        } else {                                                      // HandleUnsuccessfulStep(): StepRejected
            step_ok = false;
            radius *= 0.5;
            reuse = true;
        }
    }
    if (tid == 0) { dvs[DV_COST1] = x_cost; iv[IV_ITERS] = iter; }
}

}  // namespace be
