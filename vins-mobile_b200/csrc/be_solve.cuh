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
    double *dx, *imuJ;
};

__device__ inline SolveWs carve(const BeState &s, int b) {
    SolveWs w;
    double *p = s.scratch + (size_t)b * s.scratch_stride;
    auto take = [&](size_t n) { double *r = p; p += (n + 3) & ~(size_t)3; return r; };
    w.H = take((size_t)s.NP * s.NP); w.S = take((size_t)s.NP * s.NP); w.g = take(s.NP);
    w.hll = take(s.LCAP); w.gl = take(s.LCAP); w.w = take((size_t)s.LCAP * s.NPW);
    w.sc_p = take(s.NP); w.sc_l = take(s.LCAP); w.d_p = take(s.NP); w.d_l = take(s.LCAP);
    w.gr_p = take(s.NP); w.gr_l = take(s.LCAP); w.gn_p = take(s.NP); w.gn_l = take(s.LCAP);
    w.st_p = take(s.NP); w.st_l = take(s.LCAP); w.u_p = take(s.NP); w.u_l = take(s.LCAP);
    w.rhs = take(s.NP); w.y = take(s.NP);
    w.dx = take(s.NPX); w.imuJ = take((size_t)s.W * 930);
    return w;
}
__host__ __device__ inline size_t solve_scratch_doubles(int NP, int NPX, int NPW, int LCAP, int W) {
    auto r = [](size_t n) { return (n + 3) & ~(size_t)3; };
    return 2 * r((size_t)NP * NP) + 7 * r(NP) + 8 * r(LCAP) + r((size_t)LCAP * NPW) + r(NPX) + r((size_t)W * 930) + r(NP) * 2 + 64;
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

// cost (and, when lin != 0, H / g / landmark terms) at parameter vector `par`.  Returns the total cost in every thread.
__device__ __noinline__ double evaluate(const BeState &s, int b, const SolveWs &ws, const double *par, int lin, double *sh_red) {
    long long *pp = s.prof + (size_t)b * 32; BE_PROF2_INIT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int *iv = S_iv(s, b);
    const int NP = s.NP, NPW = s.NPW, nl = iv[IV_N_LM], nfac = iv[IV_N_FAC];
    double cost = 0.0;
    const int has_prior = iv[IV_PRIOR_VALID];
    if (lin) {
        // H starts as the prior's J0^T J0 (pose / speed-bias rows and columns of the canonical layout), or zero
        if (has_prior) {
            const double *Hp = s.Hp + (size_t)b * s.NPX * s.NPX;
            for (int e = tid; e < NP * NP; e += blockDim.x) { const int i = e / NP, j = e - i * NP; ws.H[e] = Hp[(size_t)i * s.NPX + j]; }
        } else
            for (int i = tid; i < NP * NP; i += blockDim.x) ws.H[i] = 0.0;
        for (int i = tid; i < NP; i += blockDim.x) ws.g[i] = 0.0;
        for (int i = tid; i < nl; i += blockDim.x) { ws.hll[i] = 0.0; ws.gl[i] = 0.0; }
        for (int i = tid; i < nl * NPW; i += blockDim.x) ws.w[i] = 0.0;
    }
    // ---- prior ------------------------------------------------------------------------------------------
    if (has_prior) {
        const int NPX = s.NPX;
        const double *Hp = s.Hp + (size_t)b * NPX * NPX, *bp = s.bp + (size_t)b * NPX;
        prior_dx(s, b, par, ws.dx);
        __syncthreads();
        for (int i = tid; i < NPX; i += blockDim.x) {
            double t = 0;
            for (int j = 0; j < NPX; j++) t += Hp[(size_t)j * NPX + i] * ws.dx[j];      // Hp symmetric: column read = coalesced
            cost += 0.5 * ws.dx[i] * t + bp[i] * ws.dx[i];
            if (lin && i < NP) ws.g[i] += t + bp[i];
        }
        if (tid == 0) cost += 0.5 * S_dv(s, b)[DV_PRIOR_C0];
    }
    __syncthreads();
    if (lin) BE_PROF2(pp, 16);
    // ---- IMU factors: one warp per factor ---------------------------------------------------------------
    for (int f = warp; f < s.W; f += nwarp) {
        const double *pr = S_pre(s, b, f + 1);
        double *J = ws.imuJ + (size_t)f * 930;          // raw J 450 | weighted J 450 | raw res 15 | weighted res 15
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
    if (lin) BE_PROF2(pp, 17);
    // ---- projection factors: one thread per factor ------------------------------------------------------
    {
        const double *dv = S_dv(s, b);
        ProjConst K; K.ric = ldm(dv + DV_RIC); K.tic = ld3(dv + DV_TIC); K.sqrt_info = s.sqrt_info;
        const int *fl = s.fac_lm + (size_t)b * s.PCAP, *fj = s.fac_j + (size_t)b * s.PCAP;
        const int *slot = s.lm_slot + (size_t)b * s.LCAP;
        const size_t fo = (size_t)b * s.FCAP;
        if (!lin) {
            for (int f = tid; f < nfac; f += blockDim.x) {           // cost only: one thread per factor
                const int l = fl[f], j = fj[f], k = slot[l];
                const int i = s.f_start[fo + k];
                const double *o = S_obs(s, b, k);
                const V3 pi = v3(o[0], o[1], 1.0), pj = v3(o[2 * (j - i)], o[2 * (j - i) + 1], 1.0);
                cost += proj_eval(K, pi, pj, par + 16 * i, par + 16 * j, par[16 * s.NF + l], nullptr, nullptr, nullptr, nullptr);
            }
        } else {
            // linearisation: one thread per LANDMARK.  Everything that belongs to the landmark alone (h_ll, g_l, its coupling rows w_l)
            // and the (i,i) block / g_i of its anchor frame are reduced in registers over the landmark's factors, so only the
            // (j,j), (j,i) blocks and g_j of each factor go through atomics (63 instead of 104 per factor, and the hot anchor block
            // receives one update per landmark instead of one per factor).
            for (int l = tid; l < nl; l += blockDim.x) {
                const int k = slot[l];
                const int i = s.f_start[fo + k], no = s.f_nobs[fo + k];
                const double *o = S_obs(s, b, k);
                const V3 pi = v3(o[0], o[1], 1.0);
                const double lam = par[16 * s.NF + l];
                double Hii[21], gi[6], wi[6], hl = 0, gll = 0;
#pragma unroll
                for (int q = 0; q < 21; q++) Hii[q] = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) { gi[q] = 0; wi[q] = 0; }
                const int oi = 15 * i;
                for (int t = 1; t < no; t++) {
                    const int j = i + t, oj = 15 * j;
                    const V3 pj = v3(o[2 * t], o[2 * t + 1], 1.0);
                    double r2[2], Ji[12], Jj[12], Jl[2];
                    cost += proj_eval(K, pi, pj, par + 16 * i, par + 16 * j, lam, r2, Ji, Jj, Jl);
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            if (c <= a) {
                                Hii[q++] += Ji[a] * Ji[c] + Ji[6 + a] * Ji[6 + c];
                                atomic_add(&ws.H[(size_t)(oj + a) * NP + oj + c], Jj[a] * Jj[c] + Jj[6 + a] * Jj[6 + c]);
                            }
                            atomic_add(&ws.H[(size_t)(oj + a) * NP + oi + c], Jj[a] * Ji[c] + Jj[6 + a] * Ji[6 + c]);
                        }
                        gi[a] += Ji[a] * r2[0] + Ji[6 + a] * r2[1];
                        wi[a] += Ji[a] * Jl[0] + Ji[6 + a] * Jl[1];
                        atomic_add(&ws.g[oj + a], Jj[a] * r2[0] + Jj[6 + a] * r2[1]);
                        ws.w[(size_t)l * NPW + 6 * j + a] = Jj[a] * Jl[0] + Jj[6 + a] * Jl[1];
                    }
                    hl += Jl[0] * Jl[0] + Jl[1] * Jl[1];
                    gll += Jl[0] * r2[0] + Jl[1] * r2[1];
                }
                int q = 0;
#pragma unroll
                for (int a = 0; a < 6; a++) {
#pragma unroll
                    for (int c = 0; c <= a; c++) atomic_add(&ws.H[(size_t)(oi + a) * NP + oi + c], Hii[q++]);
                    atomic_add(&ws.g[oi + a], gi[a]);
                    ws.w[(size_t)l * NPW + 6 * i + a] = wi[a];
                }
                ws.hll[l] = hl; ws.gl[l] = gll;
            }
        }
    }
    __syncthreads();
    if (lin) BE_PROF2(pp, 18);
    if (lin) {                                            // mirror the lower triangle (the prior part is already symmetric)
        for (int e = tid; e < NP * NP; e += blockDim.x) { const int i = e / NP, j = e - i * NP; if (j > i) ws.H[e] = ws.H[(size_t)j * NP + i]; }
    }
    const double total = block_sum_d(cost, sh_red);
    __syncthreads();
    if (lin) BE_PROF2(pp, 19);
    return total;
}

// u^T H u over the full (pose/speed-bias + landmark) system
__device__ inline double quad_form(const BeState &s, const SolveWs &ws, int nl, const double *up, const double *ul, double *sh_red) {
    const int tid = threadIdx.x, NP = s.NP, NPW = s.NPW, NF = s.NF;
    double acc = 0;
    for (int i = tid; i < NP; i += blockDim.x) {
        double t = 0;
        for (int j = 0; j < NP; j++) t += ws.H[(size_t)j * NP + i] * up[j];           // H symmetric: column read = coalesced
        acc += up[i] * t;
    }
    for (int l = tid; l < nl; l += blockDim.x) {
        double t = 0;
        const double *w = ws.w + (size_t)l * NPW;
        for (int f = 0; f < NF; f++)
            for (int a = 0; a < 6; a++) t += w[6 * f + a] * up[15 * f + a];
        acc += 2.0 * ul[l] * t + ws.hll[l] * ul[l] * ul[l];
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

// Blocked RIGHT-looking Cholesky on a packed lower matrix in shared memory, BORDERED by the right-hand side (row n of the packed
// array holds rhs, so the forward substitution z = L^-1 rhs is a by-product of the factorisation).  Per 8-column panel:
//   (a) warp 0 factors the 8x8 diagonal block in place (shared memory, one __syncwarp-separated step per column);
//   (b) one thread per row below the block solves its 8 entries against the block (block reads are warp broadcasts);
//   (c) the trailing matrix gets the rank-8 update in 4x4 register tiles (0.5 shared loads per FMA), spread over the whole CTA.
// Three barriers per panel.  Backward substitution L^T y = z is blocked the same way.  A must hold (n+1)(n+2)/2 doubles.
// Returns false (in all threads) on a non-positive pivot / non-finite value (Eigen LLT info() != Success).
__device__ __noinline__ bool chol_solve_packed(double *A, int n, const double *rhs, double *y, int *sh_flag, double *sh_inv /*>= 8*/, long long *pp) {
    BE_PROF2_INIT;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    constexpr int NB = 8;
    if (tid == 0) *sh_flag = 1;
    for (int j = tid; j < n; j += T) A[pidx(n, j)] = rhs[j];
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += NB) {
        const int nb = min(NB, n - c0);
        if (tid < 32) {                                              // (a) diagonal block, in place
            bool ok = true;
            for (int k = 0; k < nb; k++) {
                const double piv = A[pidx(c0 + k, c0 + k)];
                ok &= (piv > 0) && isfinite(piv);
                const double l = sqrt(piv), il = 1.0 / l;
                __syncwarp();
                if (lane == 0) { A[pidx(c0 + k, c0 + k)] = l; sh_inv[k] = il; }
                if (lane > k && lane < nb) A[pidx(c0 + lane, c0 + k)] *= il;
                __syncwarp();
                // rank-1 update of the remaining block: lanes 0..27 own the strictly-lower pairs, lanes 0..7 then the diagonal
                if (lane < 28) {
                    int i = 1, t = lane;                             // strict-lower pair index -> (i, j): i = 1..7, j = 0..i-1
                    while (t >= i) { t -= i; i++; }
                    const int j = t;
                    if (j > k && i < nb) A[pidx(c0 + i, c0 + j)] -= A[pidx(c0 + i, c0 + k)] * A[pidx(c0 + j, c0 + k)];
                }
                if (lane > k && lane < nb) { const double v = A[pidx(c0 + lane, c0 + k)]; A[pidx(c0 + lane, c0 + lane)] -= v * v; }
                __syncwarp();
            }
            if (!ok && lane == 0) *sh_flag = 0;
        }
        __syncthreads();
        BE_PROF2(pp, 21);
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
                    v[jj] = w * sh_inv[jj];
                    ri[jj] = v[jj];
                }
            }
        }
        __syncthreads();
        BE_PROF2(pp, 22);
        {                                                            // (c) trailing update, 4x4 register tiles
            const int r0 = c0 + nb;                                  // first trailing row / column
            const int R = n + 1 - r0;                                // rows r0 .. n (the rhs row included)
            const int RT = (R + 3) >> 2;
            const int ntile = RT * (RT + 1) / 2;
            for (int e = tid; e < ntile; e += T) {
                int ti = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= e) ti++;
                while (ti * (ti + 1) / 2 > e) ti--;
                const int tj = e - ti * (ti + 1) / 2;
                const int I = r0 + 4 * ti, J = r0 + 4 * tj;
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
                        if (i <= n && j <= i && j < n) A[pidx(i, j)] -= acc[a][c];
                    }
            }
        }
        __syncthreads();
        BE_PROF2(pp, 20);
    }
    // backward: L^T y = z, z = row n of A
    for (int j = tid; j < n; j += T) y[j] = A[pidx(n, j)];
    __syncthreads();
    for (int c0 = ((n - 1) / NB) * NB; c0 >= 0; c0 -= NB) {
        const int nb = min(NB, n - c0);
        if (tid == 0) {
            for (int j = nb - 1; j >= 0; j--) {
                double v = y[c0 + j];
                for (int t = j + 1; t < nb; t++) v -= A[pidx(c0 + t, c0 + j)] * y[c0 + t];
                y[c0 + j] = v / A[pidx(c0 + j, c0 + j)];
            }
        }
        __syncthreads();
        for (int i = tid; i < c0; i += T) {
            double v = y[i];
            for (int t = 0; t < nb; t++) v -= A[pidx(c0 + t, i)] * y[c0 + t];
            y[i] = v;
        }
        __syncthreads();
    }
    bool ok = true;
    for (int i = tid; i < n; i += T) ok &= isfinite(y[i]);
    BE_PROF2(pp, 23);
    return __syncthreads_and(ok) != 0;
}

constexpr int SCHUR_CHUNK = 32;
// S (packed, shared) = S H S + mu D^2 - sum_l ws_l ws_l^T / h_l ;  rhs = S g - sum_l ws_l gs_l / h_l   (landmark blocks eliminated)
// wt = shared staging [SCHUR_CHUNK][NPW + 1]
__device__ __noinline__ void build_reduced_smem(const BeState &s, const SolveWs &ws, int nl, double mu, double *S, double *wt) {
    const int tid = threadIdx.x, T = blockDim.x, NP = s.NP, NPW = s.NPW;
    for (int e = tid; e < NP * (NP + 1) / 2; e += T) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (pidx(i + 1, 0) <= e) i++;
        while (pidx(i, 0) > e) i--;
        const int j = e - pidx(i, 0);
        double v = ws.H[(size_t)i * NP + j] * ws.sc_p[i] * ws.sc_p[j];
        if (i == j) v += mu * ws.d_p[i] * ws.d_p[i];
        S[e] = v;
    }
    for (int i = tid; i < NP; i += T) ws.rhs[i] = ws.g[i] * ws.sc_p[i];
    constexpr int EPT = 10;                                         // entries of the NPW x NPW lower triangle per thread (2211 / 256 threads)
    const int nent = NPW * (NPW + 1) / 2;
    double acc[EPT];
    int ea[EPT], ec[EPT];
#pragma unroll
    for (int q = 0; q < EPT; q++) {
        acc[q] = 0;
        const int e = tid + q * T;
        int a = 0, c = 0;
        if (e < nent) {
            a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (pidx(a + 1, 0) <= e) a++;
            while (pidx(a, 0) > e) a--;
            c = e - pidx(a, 0);
        }
        ea[q] = a; ec[q] = c;
    }
    for (int l = tid; l < nl; l += T) {                             // per-landmark weight sqrt(s_l^2 / h_l) (u_l is free scratch here)
        const double sl = ws.sc_l[l];
        ws.u_l[l] = sqrt(sl * sl / (ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l]));
    }
    double racc = 0;                                                // thread a < NPW accumulates the rhs correction
    const int ld = NPW + 1;
    for (int l0 = 0; l0 < nl; l0 += SCHUR_CHUNK) {
        const int cn = min(SCHUR_CHUNK, nl - l0);
        __syncthreads();
        for (int e = tid; e < cn * ld; e += T) {
            const int cl = e / ld, a = e - cl * ld, l = l0 + cl;
            wt[e] = ((a < NPW) ? ws.w[(size_t)l * NPW + a] : ws.gl[l]) * ws.u_l[l];      // u_l holds sqrt(s_l^2 / h_l), set below
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < EPT; q++) {
            if (tid + q * T < nent) {
                double t = 0;
                for (int cl = 0; cl < cn; cl++) t += wt[cl * ld + ea[q]] * wt[cl * ld + ec[q]];
                acc[q] += t;
            }
        }
        if (tid < NPW) { double t = 0; for (int cl = 0; cl < cn; cl++) t += wt[cl * ld + tid] * wt[cl * ld + NPW]; racc += t; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < EPT; q++) {
        if (tid + q * T < nent) {
            const int ia = 15 * (ea[q] / 6) + ea[q] % 6, ic = 15 * (ec[q] / 6) + ec[q] % 6;
            S[pidx(ia, ic)] -= acc[q] * ws.sc_p[ia] * ws.sc_p[ic];
        }
    }
    if (tid < NPW) { const int ia = 15 * (tid / 6) + tid % 6; ws.rhs[ia] -= racc * ws.sc_p[ia]; }
    __syncthreads();
}

__host__ __device__ inline size_t solve_smem_bytes(int NP, int NPW) {
    return ((size_t)(NP + 1) * (NP + 2) / 2 + 8 + (size_t)SCHUR_CHUNK * (NPW + 1)) * sizeof(double);
}

__global__ void __launch_bounds__(SOLVE_T) solve_kernel(BeState s, int use_smem) {
    extern __shared__ __align__(16) double sm_dyn[];
    const int T = blockDim.x;                              // 256 or 512 (VIO_BE_THREADS)
    __shared__ double sh_red[32];
    __shared__ int sh_flag;
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    double *dvs = S_dv(s, b);
    const SolveWs ws = carve(s, b);
    const int NP = s.NP, NPW = s.NPW, NF = s.NF, nl = iv[IV_N_LM];
    double *par = s.par + (size_t)b * (NF * 16 + s.LCAP);
    double *cand = s.cand + (size_t)b * (NF * 16 + s.LCAP);

    BE_PROF_INIT;
    double x_cost = evaluate(s, b, ws, par, 1, sh_red);
    BE_PROF(0);
    if (tid == 0) dvs[DV_COST0] = x_cost;
    // Jacobi scaling, frozen at iteration 0 (trust_region_minimizer.cc:239-254)
    for (int i = tid; i < NP; i += T) ws.sc_p[i] = 1.0 / (1.0 + sqrt(ws.H[(size_t)i * NP + i]));
    for (int l = tid; l < nl; l += T) ws.sc_l[l] = 1.0 / (1.0 + sqrt(ws.hll[l]));
    __syncthreads();

    double radius = 1e4, mu = 1e-8, alpha = 0.0, dogleg_norm = 0.0, x_norm = -1.0;
    bool reuse = false;
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
            BE_PROF(1);
            // ---- ComputeGaussNewtonStep: (S H S + mu D^2) y = S g by Schur elimination of the landmark blocks ----
            linear_ok = false;
            while (mu < 1.0) {
              bool ok;
              if (use_smem) {
                double *Ssm = sm_dyn, *wt = sm_dyn + (size_t)(NP + 1) * (NP + 2) / 2 + 8;
                build_reduced_smem(s, ws, nl, mu, Ssm, wt);
                BE_PROF(2);
                ok = chol_solve_packed(Ssm, NP, ws.rhs, ws.y, &sh_flag, sh_red, s.prof + (size_t)b * 32);
                BE_PROF(3);
              } else {
                for (int e = tid; e < NP * NP; e += T) {
                    const int i = e / NP, j = e - i * NP;
                    if (j <= i) {
                        double v = ws.H[e] * ws.sc_p[i] * ws.sc_p[j];
                        if (i == j) v += mu * ws.d_p[i] * ws.d_p[i];
                        ws.S[e] = v;
                    }
                }
                for (int i = tid; i < NP; i += T) ws.rhs[i] = ws.g[i] * ws.sc_p[i];
                __syncthreads();
                // S -= sum_l ws_l ws_l^T / h_l on the pose (6-dof) rows/cols; rhs -= ws_l gs_l / h_l
                for (int e = tid; e < NPW * NPW; e += T) {
                    const int a = e / NPW, c = e - a * NPW;
                    if (c > a) continue;
                    const int ia = 15 * (a / 6) + a % 6, ic = 15 * (c / 6) + c % 6;
                    double acc = 0;
                    for (int l = 0; l < nl; l++) {
                        const double wa = ws.w[(size_t)l * NPW + a];
                        if (wa == 0.0) continue;
                        const double sl = ws.sc_l[l];
                        const double h = ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l];
                        acc += wa * ws.w[(size_t)l * NPW + c] * (sl * sl / h);
                    }
                    ws.S[(size_t)ia * NP + ic] -= acc * ws.sc_p[ia] * ws.sc_p[ic];
                }
                for (int a = tid; a < NPW; a += T) {
                    const int ia = 15 * (a / 6) + a % 6;
                    double acc = 0;
                    for (int l = 0; l < nl; l++) {
                        const double sl = ws.sc_l[l];
                        const double h = ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l];
                        acc += ws.w[(size_t)l * NPW + a] * ws.gl[l] * (sl * sl / h);
                    }
                    ws.rhs[ia] -= acc * ws.sc_p[ia];
                }
                __syncthreads();
                ok = chol_solve(ws.S, NP, ws.rhs, ws.y, &sh_flag);
              }
                __syncthreads();
                if (ok) {
                    // back-substitution y_l = (gs_l - ws_l . y_p) / h_l ;  gauss_newton_step_ = -diagonal .* y
                    for (int l = tid; l < nl; l += T) {
                        const double sl = ws.sc_l[l];
                        const double h = ws.hll[l] * sl * sl + mu * ws.d_l[l] * ws.d_l[l];
                        double t = 0;
                        const double *w = ws.w + (size_t)l * NPW;
                        for (int f = 0; f < NF; f++)
                            for (int a = 0; a < 6; a++) t += w[6 * f + a] * ws.sc_p[15 * f + a] * ws.y[15 * f + a];
                        const double yl = (ws.gl[l] * sl - sl * t) / h;
                        ws.gn_l[l] = -ws.d_l[l] * yl;
                    }
                    for (int i = tid; i < NP; i += T) ws.gn_p[i] = -ws.d_p[i] * ws.y[i];
                    __syncthreads();
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
            const double g_norm = sqrt(dot2(ws.gr_p, ws.gr_l, ws.gr_p, ws.gr_l, NP, nl, sh_red));
            const double gn_norm = sqrt(dot2(ws.gn_p, ws.gn_l, ws.gn_p, ws.gn_l, NP, nl, sh_red));
            double ca, cb;                                            // step = ca * gradient_ + cb * gauss_newton_step_
            if (gn_norm <= radius) { ca = 0; cb = 1; dogleg_norm = gn_norm; }
            else if (g_norm * alpha >= radius) { ca = -(radius / g_norm); cb = 0; dogleg_norm = radius; }
            else {
                const double b_dot_a = -alpha * dot2(ws.gr_p, ws.gr_l, ws.gn_p, ws.gn_l, NP, nl, sh_red);
                const double a2 = (alpha * g_norm) * (alpha * g_norm);
                const double bma2 = a2 - 2 * b_dot_a + gn_norm * gn_norm;
                const double c = b_dot_a - a2;
                const double d = sqrt(c * c + bma2 * (radius * radius - a2));
                const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
                ca = -alpha * (1.0 - beta); cb = beta;
                dogleg_norm = -1.0;
            }
            for (int i = tid; i < NP; i += T) ws.st_p[i] = ca * ws.gr_p[i] + cb * ws.gn_p[i];
            for (int l = tid; l < nl; l += T) ws.st_l[l] = ca * ws.gr_l[l] + cb * ws.gn_l[l];
            __syncthreads();
            if (dogleg_norm < 0) dogleg_norm = sqrt(dot2(ws.st_p, ws.st_l, ws.st_p, ws.st_l, NP, nl, sh_red));
            // trust_region_step_ = dogleg ./ diagonal ;  delta = trust_region_step_ .* jacobian_scaling_
            for (int i = tid; i < NP; i += T) ws.u_p[i] = ws.st_p[i] / ws.d_p[i] * ws.sc_p[i];
            for (int l = tid; l < nl; l += T) ws.u_l[l] = ws.st_l[l] / ws.d_l[l] * ws.sc_l[l];
            __syncthreads();
            // model_cost_change = -(J d)^T (r + J d / 2) = -(d^T g + d^T H d / 2)
            const double dg = dot2(ws.u_p, ws.u_l, ws.g, ws.gl, NP, nl, sh_red);
            const double dHd = quad_form(s, ws, nl, ws.u_p, ws.u_l, sh_red);
            model_change = -(dg + 0.5 * dHd);
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
        const double cand_cost = evaluate(s, b, ws, cand, 0, sh_red);
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
            x_cost = evaluate(s, b, ws, par, 1, sh_red);
            BE_PROF(0);
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
