// be_marg.cuh -- K13: window marginalisation, one CTA per stream.
// Replaces MarginalizationInfo::{addResidualBlockInfo,preMarginalize,marginalize,getParameterBlocks} and the two call sites
// in VINS::solve_ceres (marginalization_factor.cpp:11-322; VINS.cpp:690-830):
//   A = sum J^T J, b = sum J^T r over {old prior, IMUFactor(0->1), ProjectionFactors of landmarks that start in frame 0} (MARGIN_OLD)
//       or over {old prior} (MARGIN_SECOND_NEW);   Amm^+ by symmetric eigendecomposition with eigenvalue threshold eps = 1e-8;
//   A_r = Arr - Arm Amm^+ Amr, b_r = brr - Arm Amm^+ bmm;   second eigendecomposition, eigenvalues <= eps dropped:
//   J0 = S^{1/2} V^T, r0 = S^{-1/2} V^T b_r.
// The new prior is kept in information form over the canonical layout: Hp = J0^T J0 = V S+ V^T, bp = J0^T r0 = V+ V+^T b_r,
// c0 = |r0|^2 -- everything MarginalizationFactor::Evaluate contributes to the next solve (cost, gradient, J^T J).
// The reference orders blocks by unordered_map iteration over pointer values (quirk Q10); here the order is canonical.
#pragma once
#include "be_solve.cuh"

namespace be {

constexpr int MARG_T = 512;
constexpr double MARG_EPS = 1e-8;      // MarginalizationInfo::eps, marginalization_factor.hpp:75
constexpr int MARG_NCAP = 104;         // A_r and its eigenvectors live in shared memory when n <= MARG_NCAP (2*104^2*8 = 173 KB)

// Cyclic two-sided Jacobi eigensolver, parallel (round-robin) ordering, A symmetric n x n (ld), destroyed: on exit its diagonal
// holds the eigenvalues and V (n x n, ld = n) the eigenvectors as columns.  Per round the n/2 disjoint rotations are computed
// (phase 1) and then every 2x2 block (pair a, pair b) of A' = J^T A J is updated by ONE thread (phase 2, together with V' = V J):
// two barriers per round.  cs/pq: shared scratch for 2*(n/2+1) doubles / ints.
__device__ inline int eig_sym_jacobi(double *A, int n, int ld, double *V, double *cs, int *pq, double *sh_red) {
    const int tid = threadIdx.x, T = blockDim.x;
    for (int e = tid; e < n * n; e += T) { const int i = e / n, j = e - i * n; V[e] = (i == j) ? 1.0 : 0.0; }
    __syncthreads();
    if (n < 2) return 0;
    const int ne = (n + 1) & ~1, np = ne / 2;
    int sweep = 0;
    for (; sweep < 40; sweep++) {
        double off = 0, dg = 0;
        for (int e = tid; e < n * n; e += T) { const int i = e / n, j = e - i * n; const double a = A[(size_t)i * ld + j]; if (i == j) dg += a * a; else off += a * a; }
        off = block_sum_d(off, sh_red);
        dg = block_sum_d(dg, sh_red);
        // Stop at the round-off floor of the MATRIX (|A|_F * eps per entry), not of the individual pivots: A_r has (gauge) eigenvalues
        // that are pure noise, and a pivot-relative test would chase them for the full 40 sweeps.
        const double normA = sqrt(off + dg);
        if (off == 0.0 || sqrt(off) <= 1e-14 * normA) break;
        const double skip = 1e-17 * normA;
        for (int round = 0; round < ne - 1; round++) {
            for (int k = tid; k < np; k += T) {
                int p = (k == 0) ? ne - 1 : (round + k) % (ne - 1);
                int q = (round + ne - 1 - k) % (ne - 1);
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, sn = 0.0;
                if (q < n) {
                    const double app = A[(size_t)p * ld + p], aqq = A[(size_t)q * ld + q], apq = A[(size_t)p * ld + q];
                    if (fabs(apq) > skip) {
                        const double zeta = (aqq - app) / (2.0 * apq);
                        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        c = 1.0 / sqrt(1.0 + t * t); sn = c * t;
                    }
                } else q = -1;                                         // odd n: p is the bye of this round
                pq[2 * k] = p; pq[2 * k + 1] = q; cs[2 * k] = c; cs[2 * k + 1] = sn;
            }
            __syncthreads();
            const int nblk = np * np;
            for (int e = tid; e < nblk + n * np; e += T) {
                if (e < nblk) {                                         // A block (pair a rows) x (pair b cols)
                    const int a = e / np, b = e - a * np;
                    const int pa = pq[2 * a], qa = pq[2 * a + 1], pb = pq[2 * b], qb = pq[2 * b + 1];
                    const double ca = cs[2 * a], sa = cs[2 * a + 1], cb = cs[2 * b], sb = cs[2 * b + 1];
                    const double x00 = A[(size_t)pa * ld + pb];
                    const double x01 = qb >= 0 ? A[(size_t)pa * ld + qb] : 0.0;
                    const double x10 = qa >= 0 ? A[(size_t)qa * ld + pb] : 0.0;
                    const double x11 = (qa >= 0 && qb >= 0) ? A[(size_t)qa * ld + qb] : 0.0;
                    const double r00 = ca * x00 - sa * x10, r01 = ca * x01 - sa * x11;
                    const double r10 = sa * x00 + ca * x10, r11 = sa * x01 + ca * x11;
                    A[(size_t)pa * ld + pb] = cb * r00 - sb * r01;
                    if (qb >= 0) A[(size_t)pa * ld + qb] = sb * r00 + cb * r01;
                    if (qa >= 0) A[(size_t)qa * ld + pb] = cb * r10 - sb * r11;
                    if (qa >= 0 && qb >= 0) A[(size_t)qa * ld + qb] = sb * r10 + cb * r11;
                } else {                                                // V columns of pair b, row i
                    const int f = e - nblk;
                    const int i = f / np, b = f - i * np;
                    const int pb = pq[2 * b], qb = pq[2 * b + 1];
                    if (qb < 0) continue;
                    const double cb = cs[2 * b], sb = cs[2 * b + 1];
                    const double x = V[(size_t)i * n + pb], y = V[(size_t)i * n + qb];
                    V[(size_t)i * n + pb] = cb * x - sb * y; V[(size_t)i * n + qb] = sb * x + cb * y;
                }
            }
            __syncthreads();
        }
    }
    return sweep;
}


// Symmetric eigensolver by Householder tridiagonalisation + implicit-shift QL (the EISPACK tred2/tql2 pair; Eigen's
// SelfAdjointEigenSolver, which the reference calls at marginalization_factor.cpp:270,290, is the same two-stage scheme).
// Same contract as eig_sym_jacobi: A symmetric n x n (ld); on exit diag(A) = eigenvalues, V (ld n) = eigenvectors as columns.
// Parallelisation: every O(n^2) inner loop of tred2 is spread over the CTA; in tql2 one thread runs the scalar rotation
// recurrence of a QL step, then every thread applies the whole rotation sequence to the rows of V it owns.
// de: shared scratch for 2n doubles (d, e); rot: shared scratch for 2n doubles.  Returns the number of QL steps.
__device__ inline int eig_sym_ql(double *A, int n, int ld, double *V, double *de, double *rot, double *sh_red, int *sh_i, long long *pp) {
    BE_PROF2_INIT;
    const int tid = threadIdx.x, T = blockDim.x;
    double *d = de, *e = de + n;
    for (int x = tid; x < n * n; x += T) { const int i = x / n, j = x - i * n; V[x] = A[(size_t)i * ld + j]; }
    __syncthreads();
    double *a = V;                                          // a[i][k] = a[i * n + k]
    // ---- tred2 ------------------------------------------------------------------------------------------------------------
    for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double h = 0.0;
        if (l > 0) {
            double sc = 0;
            for (int k = tid; k <= l; k += T) sc += fabs(a[i * n + k]);
            const double scale = block_sum_d(sc, sh_red);
            __syncthreads();
            if (scale == 0.0) {
                if (tid == 0) e[i] = a[i * n + l];
            } else {
                double hp = 0;
                for (int k = tid; k <= l; k += T) { const double v = a[i * n + k] / scale; a[i * n + k] = v; hp += v * v; }
                h = block_sum_d(hp, sh_red);
                __syncthreads();
                const double f0 = a[i * n + l];
                const double g0 = f0 >= 0 ? -sqrt(h) : sqrt(h);
                h -= f0 * g0;
                __syncthreads();
                if (tid == 0) { e[i] = scale * g0; a[i * n + l] = f0 - g0; }
                __syncthreads();
                // e[j] = (A u)_j / h ,  a[j][i] = u_j / h
                double fp = 0;
                for (int j = tid; j <= l; j += T) {
                    a[j * n + i] = a[i * n + j] / h;
                    double g = 0;
                    for (int k = 0; k <= j; k++) g += a[j * n + k] * a[i * n + k];
                    for (int k = j + 1; k <= l; k++) g += a[k * n + j] * a[i * n + k];
                    e[j] = g / h;
                    fp += (g / h) * a[i * n + j];
                }
                const double f = block_sum_d(fp, sh_red);
                __syncthreads();
                const double hh = f / (h + h);
                for (int j = tid; j <= l; j += T) e[j] -= hh * a[i * n + j];
                __syncthreads();
                for (int x = tid; x < (l + 1) * (l + 1); x += T) {
                    const int j = x / (l + 1), k = x - j * (l + 1);
                    if (k <= j) a[j * n + k] -= a[i * n + j] * e[k] + e[j] * a[i * n + k];
                }
            }
        } else if (tid == 0) e[i] = a[i * n + l];
        __syncthreads();
        if (tid == 0) d[i] = h;
        __syncthreads();
    }
    if (tid == 0) { d[0] = 0.0; e[0] = 0.0; }
    __syncthreads();
    BE_PROF2(pp, 24);
    for (int i = 0; i < n; i++) {                           // accumulate the transformation
        const int l = i - 1;
        if (d[i] != 0.0 && l >= 0) {
            for (int j = tid; j <= l; j += T) {
                double g = 0;
                for (int k = 0; k <= l; k++) g += a[i * n + k] * a[k * n + j];
                rot[j] = g;
            }
            __syncthreads();
            for (int x = tid; x < (l + 1) * (l + 1); x += T) {
                const int k = x / (l + 1), j = x - k * (l + 1);
                a[k * n + j] -= rot[j] * a[k * n + i];
            }
        }
        __syncthreads();
        if (tid == 0) { d[i] = a[i * n + i]; a[i * n + i] = 1.0; }
        for (int j = tid; j <= l; j += T) { a[j * n + i] = 0.0; a[i * n + j] = 0.0; }
        __syncthreads();
    }
    BE_PROF2(pp, 25);
    // ---- tql2 -------------------------------------------------------------------------------------------------------------
    if (tid == 0) { for (int i = 1; i < n; i++) e[i - 1] = e[i]; e[n - 1] = 0.0; }
    __syncthreads();
    int steps = 0;
    for (int l = 0; l < n; l++) {
        for (int iter = 0; iter < 60; iter++) {
            if (tid == 0) {
                int m = l;
                for (; m < n - 1; m++) { const double dd = fabs(d[m]) + fabs(d[m + 1]); if (fabs(e[m]) + dd == dd) break; }
                sh_i[0] = m; sh_i[1] = l;                   // rotations i = m-1 .. first (inclusive); first = l unless the sweep aborts early
                if (m != l) {
                    double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                    double r = hypot(g, 1.0);
                    g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
                    double sn = 1.0, c = 1.0, p = 0.0;
                    int i = m - 1;
                    bool brk = false;
                    double ei = e[i], di = d[i], dip1 = d[i + 1];          // software-pipelined operands of rotation i
                    for (; i >= l; i--) {
                        const double en = i > l ? e[i - 1] : 0.0, dn = i > l ? d[i - 1] : 0.0;
                        const double f = sn * ei;
                        const double bb = c * ei;
                        r = sqrt(f * f + g * g);                           // |f|,|g| are O(|A|): no overflow concern at this scale
                        e[i + 1] = r;
                        if (r == 0.0) { d[i + 1] = dip1 - p; e[m] = 0.0; brk = true; break; }
                        const double ir = 1.0 / r;
                        sn = f * ir; c = g * ir;
                        g = dip1 - p;
                        r = (di - g) * sn + 2.0 * c * bb;
                        p = sn * r;
                        d[i + 1] = g + p;
                        g = c * r - bb;
                        rot[2 * i] = c; rot[2 * i + 1] = sn;
                        dip1 = di; di = dn; ei = en;
                    }
                    sh_i[1] = brk ? i + 1 : l;
                    if (!brk) { d[l] -= p; e[l] = g; e[m] = 0.0; }
                }
            }
            __syncthreads();
            const int m = sh_i[0], first = sh_i[1];
            __syncthreads();                                // every thread has read sh_i before thread 0 may overwrite it (next l / next step)
            if (m == l) break;
            steps++;
            for (int k = tid; k < n; k += T) {              // apply the rotation sequence to row k of V
                double *zr = a + (size_t)k * n;
                for (int i = m - 1; i >= first; i--) {
                    const double c = rot[2 * i], sn = rot[2 * i + 1];
                    const double f = zr[i + 1];
                    zr[i + 1] = sn * zr[i] + c * f;
                    zr[i] = c * zr[i] - sn * f;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += T) A[(size_t)i * ld + i] = d[i];
    __syncthreads();
    BE_PROF2(pp, 26);
    return steps;
}

// ProjectionFactor Jacobian wrt para_Ex_Pose (projection_facor.cpp:73-84), corrected by sqrt(rho'): 2x6
__device__ inline void proj_jac_ex(const ProjConst &K, V3 pts_i, const double *pi, const double *pj, double inv_dep, double *Jex) {
    const V3 Pi = ld3(pi), Pj = ld3(pj);
    const M3 Ri = q2R(ldq(pi + 3)), Rj = q2R(ldq(pj + 3));
    const V3 pc_i = pts_i * (1.0 / inv_dep);
    const V3 p_imu_i = K.ric * pc_i + K.tic;
    const V3 pw = Ri * p_imu_i + Pi;
    const M3 RjT = tr(Rj), ricT = tr(K.ric);
    const V3 p_imu_j = RjT * (pw - Pj);
    const V3 pc_j = ricT * (p_imu_j - K.tic);
    const double dep = pc_j.z;
    double red[6];
    red[0] = K.sqrt_info / dep; red[1] = 0; red[2] = -K.sqrt_info * pc_j.x / (dep * dep);
    red[3] = 0; red[4] = K.sqrt_info / dep; red[5] = -K.sqrt_info * pc_j.y / (dep * dep);
    const M3 L = ricT * (RjT * Ri - eye3());
    const M3 tmp_r = ricT * RjT * Ri * K.ric;
    const M3 Rr = (-1.0 * (tmp_r * skew(pc_i))) + skew(tmp_r * pc_i) + skew(ricT * (RjT * (Ri * K.tic + Pi - Pj) - K.tic));
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) {
            double a = 0, bb = 0;
            for (int k = 0; k < 3; k++) { a += red[3 * r + k] * L.m[3 * k + c]; bb += red[3 * r + k] * Rr.m[3 * k + c]; }
            Jex[6 * r + c] = a; Jex[6 * r + 3 + c] = bb;
        }
}

struct MargSmem {
    int c2w[VIO_MAX_WIN * 15 + 15 + 6 + 8];      // canonical dof -> work index (-1 = absent)
    int kept[VIO_MAX_WIN * 15 + 15 + 6 + 8];     // kept work position -> canonical dof
    int pq[2 * 256];
    double cs[2 * 256];
    double de[2 * 400];
    double rot[2 * 400];
    int ql_i[4];
    double red[32];
    int scan[33];
    int n_eff, m, L0, mc;
    long long prof_sink[32];
};

__global__ void __launch_bounds__(MARG_T) marg_kernel(BeState s) {
    VIO_POISON(64u);
    extern __shared__ __align__(16) unsigned char smraw[];
    MargSmem &sm = *reinterpret_cast<MargSmem *>(smraw);
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    const int W = s.W, NF = s.NF, NP = s.NP, NPX = s.NPX;
    const int marg = iv[IV_MARG_FLAG];
    int *pres = s.present + (size_t)b * (2 * NF + 1);
    const int prior_valid = iv[IV_PRIOR_VALID];
    if (marg == 1 && !(prior_valid && pres[2 * (W - 1)])) return;         // VINS.cpp:779-780: nothing to do
    double *dvs = S_dv(s, b);
    double *par = s.par + (size_t)b * s.par_stride;
    const size_t fo = (size_t)b * s.FCAP;
    const int nl = iv[IV_N_LM], nfac = iv[IV_N_FAC];
    BE_PROF_INIT;
    // old2new() after new2old(): repack the (re-anchored) state; inverse depths come back from the feature table
    for (int i = tid; i < NF; i += T) {
        double *p = par + 16 * i;
        st3(p, ld3(S_Ps(s, b, i))); stq(p + 3, R2q(ldm(S_Rs(s, b, i))));
        st3(p + 7, ld3(S_Vs(s, b, i))); st3(p + 10, ld3(S_Bas(s, b, i))); st3(p + 13, ld3(S_Bgs(s, b, i)));
    }
    for (int l = tid; l < nl; l += T) par[16 * NF + l] = 1.0 / s.f_depth[fo + s.lm_slot[(size_t)b * s.LCAP + l]];
    // ---- which canonical dofs are marginalised / kept --------------------------------------------------------
    // new "present" set = blocks touched by any factor of this marginalisation
    __shared__ int touched[2 * (VIO_MAX_WIN + 1) + 1];
    for (int i = tid; i < 2 * NF + 1; i += T) touched[i] = prior_valid ? pres[i] : 0;
    __syncthreads();
    int *lm_work = (int *)(s.scratch + (size_t)b * s.scratch_stride);    // [LCAP] work index of landmark l (or -1)
    double *base = s.scratch + (size_t)b * s.scratch_stride + ((s.LCAP + 7) / 2);
    if (marg == 0) {
        if (tid == 0) { touched[0] = touched[1] = touched[2] = touched[3] = 1; }
        __syncthreads();
        // landmarks that start in frame 0, in list order
        int lbase = 0;
        for (int c0 = 0; c0 < nl; c0 += T) {
            const int l = c0 + tid;
            int is = 0;
            if (l < nl) is = s.f_start[fo + s.lm_slot[(size_t)b * s.LCAP + l]] == 0;
            int tot;
            const int r = lbase + block_excl_scan(is, sm.scan, &tot);
            __syncthreads();
            if (l < nl) lm_work[l] = is ? 15 + r : -1;
            lbase += tot;
        }
        if (tid == 0) { sm.L0 = lbase; sm.mc = 15; }
        for (int f = tid; f < nfac; f += T) {
            const int l = s.fac_lm[(size_t)b * s.PCAP + f];
            if (s.f_start[fo + s.lm_slot[(size_t)b * s.LCAP + l]] == 0) { touched[2 * s.fac_j[(size_t)b * s.PCAP + f]] = 1; touched[2 * NF] = 1; }
        }
    } else {
        if (tid == 0) { sm.L0 = 0; sm.mc = 6; }
    }
    __syncthreads();
    const int L0 = sm.L0, mc = sm.mc, m = mc + L0;
    if (tid == 0) {
        // work order: [marginalised canonical dofs | landmarks | kept canonical dofs (present only)]
        int k = 0;
        for (int c = 0; c < NPX; c++) {
            const int blk = c < NP ? 2 * (c / 15) + ((c % 15) >= 6) : 2 * NF;
            bool is_marg;
            if (marg == 0) is_marg = c < 15;
            else is_marg = (c >= 15 * (W - 1) && c < 15 * (W - 1) + 6);
            if (is_marg) { sm.c2w[c] = (marg == 0) ? c : c - 15 * (W - 1); continue; }
            if (touched[blk]) { sm.c2w[c] = m + k; sm.kept[k] = c; k++; } else sm.c2w[c] = -1;
        }
        sm.n_eff = k; sm.m = m;
    }
    __syncthreads();
    const int n = sm.n_eff, pos = m + n;
    // scratch carve: A[pos*pos] b[pos] Vm[m*m] T[m*(n+1)] Ar[n*n] br[n] Vr[n*n] dx[NPX] tvec[NPX]
    double *A = base;
    double *bv = A + (size_t)pos * pos;
    double *Vm = bv + pos;
    double *Tm = Vm + (size_t)m * m;
    double *Zm = Tm + (size_t)m * (n + 1);
    double *Ar = Zm + (size_t)m * (n + 1);
    double *br = Ar + (size_t)n * n;
    double *Vr = br + n;
    double *dx = Vr + (size_t)n * n;
    double *tv = dx + NPX;
    for (size_t e = tid; e < (size_t)pos * pos + pos; e += T) A[e] = 0.0;      // A and b are contiguous
    __syncthreads();
    BE_PROF(8);
    // ---- accumulate A, b ---------------------------------------------------------------------------------------
    if (prior_valid) {                                      // MarginalizationFactor of the previous prior, evaluated at the current point
        const double *Hp = s.Hp + (size_t)b * NPX * NPX, *bp = s.bp + (size_t)b * NPX;
        prior_dx(s, b, par, dx);
        __syncthreads();
        for (int i = tid; i < NPX; i += T) {
            double t = 0;
            for (int j = 0; j < NPX; j++) t += Hp[(size_t)i * NPX + j] * dx[j];
            const int wi = sm.c2w[i];
            if (wi >= 0) bv[wi] += t + bp[i];
        }
        for (int e = tid; e < NPX * NPX; e += T) {
            const int i = e / NPX, j = e - i * NPX;
            const int wi = sm.c2w[i], wj = sm.c2w[j];
            if (wi >= 0 && wj >= 0) A[(size_t)wi * pos + wj] += Hp[e];
        }
        __syncthreads();
    }
    if (marg == 0) {
        // IMUFactor(pre_integrations[1]) on {Pose0, SpeedBias0, Pose1, SpeedBias1}: one warp
        if (tid < 32) {
            const int lane = tid;
            const double *pr = S_pre(s, b, 1);
            double *J = tv + NPX;                           // 930 doubles of scratch after tv
            double *Jw = J + 450, *rr = J + 900, *rw = J + 915;
            imu_residual_warp(pr, s.gravity, par, par + 7, par + 16, par + 23, rr, J, lane);
            __syncwarp();
            const double *U = pr + PR_SQI;
            if (lane < 15) { double t = 0; for (int k = lane; k < 15; k++) t += U[lane * 15 + k] * rr[k]; rw[lane] = t; }
            for (int e = lane; e < 450; e += 32) {
                const int r = e / 30, c = e - r * 30;
                double t = 0;
                for (int k = r; k < 15; k++) t += U[r * 15 + k] * J[k * 30 + c];
                Jw[e] = t;
            }
            __syncwarp();
            for (int e = lane; e < 900; e += 32) {
                const int r = e / 30, c = e - r * 30;
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + r] * Jw[k * 30 + c];
                atomic_add(&A[(size_t)sm.c2w[r] * pos + sm.c2w[c]], t);
            }
            for (int c = lane; c < 30; c += 32) {
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + c] * rw[k];
                atomic_add(&bv[sm.c2w[c]], t);
            }
        }
        // ProjectionFactors of the landmarks that start in frame 0: blocks {Pose0, Pose_j, Ex_Pose, Feature}
        ProjConst K; K.ric = ldm(dvs + DV_RIC); K.tic = ld3(dvs + DV_TIC); K.sqrt_info = s.sqrt_info;
        for (int f = tid; f < nfac; f += T) {
            const int l = s.fac_lm[(size_t)b * s.PCAP + f], j = s.fac_j[(size_t)b * s.PCAP + f];
            const int lw = lm_work[l];
            if (lw < 0) continue;
            const int k = s.lm_slot[(size_t)b * s.LCAP + l];
            const double *o = S_obs(s, b, k);
            const V3 pi = v3(o[0], o[1], 1.0), pj = v3(o[2 * j], o[2 * j + 1], 1.0);
            double r2[2], J[2][19], Ji[12], Jj[12], Jl[2], Jex[12], sq;
            proj_eval(K, pi, pj, par, par + 16 * j, par[16 * NF + l], r2, Ji, Jj, Jl, &sq);
            proj_jac_ex(K, pi, par, par + 16 * j, par[16 * NF + l], Jex);
            const double sr = sqrt(fmax(2.2250738585072014e-308, 1.0 / (1.0 + sq)));
            int idx[19];
            for (int a = 0; a < 6; a++) { idx[a] = sm.c2w[a]; idx[6 + a] = sm.c2w[15 * j + a]; idx[12 + a] = sm.c2w[NP + a]; }
            idx[18] = lw;
            for (int r = 0; r < 2; r++) {
                for (int a = 0; a < 6; a++) { J[r][a] = Ji[6 * r + a]; J[r][6 + a] = Jj[6 * r + a]; J[r][12 + a] = sr * Jex[6 * r + a]; }
                J[r][18] = Jl[r];
            }
            for (int a = 0; a < 19; a++) {
                // one triangle only (row index >= column index in work order); mirrored after the accumulation
                for (int c = 0; c < 19; c++)
                    if (idx[a] > idx[c] || (idx[a] == idx[c] && a >= c)) atomic_add(&A[(size_t)idx[a] * pos + idx[c]], J[0][a] * J[0][c] + J[1][a] * J[1][c]);
                atomic_add(&bv[idx[a]], J[0][a] * r2[0] + J[1][a] * r2[1]);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < pos * pos; e += T) { const int i = e / pos, j = e - i * pos; if (j > i) A[e] = A[(size_t)j * pos + i]; }     // mirror
    __syncthreads();
    BE_PROF(9);
    // ---- Amm^+ [Amr | bmm] ------------------------------------------------------------------------------------------------
    // Reference: pseudo-inverse by eigendecomposition of 0.5*(Amm + Amm^T), eigenvalues <= eps dropped
    // (marginalization_factor.cpp:268-271).  When every eigenvalue is provably > eps the pseudo-inverse IS the inverse, and Amm has
    // the structure [[P, C], [C^T, D]] with D DIAGONAL (a projection factor touches one landmark only), so the inverse follows from
    // the mc x mc Schur complement S = P - C D^-1 C^T.  Proof obligation checked at run time: lambda_min(Amm) >= 1/||Amm^-1||_F,
    // with ||Amm^-1||_F formed from the explicit block inverse.  If the bound does not clear eps (or a pivot is not positive) the
    // kernel falls back to the faithful Jacobi eigendecomposition below.
    for (int e = tid; e < m * m; e += T) { const int i = e / m, j = e - i * m; if (j < i) { const double v = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); A[(size_t)i * pos + j] = v; A[(size_t)j * pos + i] = v; } }
    __syncthreads();
    __shared__ double Sinv[15 * 15];
    __shared__ int fast_ok;
    {
        // Gp = C D^-1 (mc x L0) in Vm scratch ; S = P - Gp C^T
        double *Gp = Vm, *Gm = Vm + (size_t)mc * L0;           // Gm = S^-1 Gp
        if (tid == 0) fast_ok = s.force_slow_marg ? 0 : 1;
        __syncthreads();
        for (int e = tid; e < mc * L0; e += T) {
            const int r = e / L0, l = e - r * L0;
            const double d = A[(size_t)(mc + l) * pos + mc + l];
            if (!(d > 0)) fast_ok = 0;
            Gp[e] = A[(size_t)r * pos + mc + l] / d;
        }
        __syncthreads();
        __shared__ double Sm[15 * 15];
        for (int e = tid; e < mc * mc; e += T) {
            const int r = e / mc, c = e - r * mc;
            double t = A[(size_t)r * pos + c];
            for (int l = 0; l < L0; l++) t -= Gp[(size_t)r * L0 + l] * A[(size_t)c * pos + mc + l];
            Sm[e] = t;
        }
        __syncthreads();
        if (tid == 0 && fast_ok) {                              // S^-1 by Cholesky (serial, mc <= 15)
            double Lc[225], Li[225];
            bool ok = true;
            for (int i = 0; i < mc * mc; i++) { Lc[i] = 0; Li[i] = 0; }
            for (int j = 0; j < mc && ok; j++) {
                double d = Sm[j * mc + j];
                for (int k = 0; k < j; k++) d -= Lc[j * mc + k] * Lc[j * mc + k];
                if (!(d > 0) || !isfinite(d)) { ok = false; break; }
                d = sqrt(d); Lc[j * mc + j] = d;
                for (int i = j + 1; i < mc; i++) { double v = Sm[i * mc + j]; for (int k = 0; k < j; k++) v -= Lc[i * mc + k] * Lc[j * mc + k]; Lc[i * mc + j] = v / d; }
            }
            if (ok) {
                for (int c = 0; c < mc; c++)
                    for (int i = c; i < mc; i++) { double v = (i == c) ? 1.0 : 0.0; for (int k = c; k < i; k++) v -= Lc[i * mc + k] * Li[k * mc + c]; Li[i * mc + c] = v / Lc[i * mc + i]; }
                for (int i = 0; i < mc; i++)
                    for (int j = 0; j <= i; j++) { double v = 0; for (int k = i; k < mc; k++) v += Li[k * mc + i] * Li[k * mc + j]; Sinv[i * mc + j] = v; Sinv[j * mc + i] = v; }
            } else fast_ok = 0;
        }
        __syncthreads();
        if (fast_ok) {
            for (int e = tid; e < mc * L0; e += T) {
                const int r = e / L0, l = e - r * L0;
                double t = 0;
                for (int k = 0; k < mc; k++) t += Sinv[r * mc + k] * Gp[(size_t)k * L0 + l];
                Gm[e] = t;
            }
            __syncthreads();
            // ||Amm^-1||_F^2 = ||S^-1||^2 + 2||S^-1 C D^-1||^2 + ||D^-1 + D^-1 C^T S^-1 C D^-1||^2
            double f2 = 0;
            for (int e = tid; e < mc * mc; e += T) f2 += Sinv[e] * Sinv[e];
            for (int e = tid; e < mc * L0; e += T) f2 += 2.0 * Gm[e] * Gm[e];
            for (int e = tid; e < L0 * L0; e += T) {
                const int a = e / L0, bq = e - a * L0;
                const double da = A[(size_t)(mc + a) * pos + mc + a];
                double t = (a == bq) ? 1.0 : 0.0;
                for (int k = 0; k < mc; k++) t += A[(size_t)k * pos + mc + a] * Gm[(size_t)k * L0 + bq];
                t /= da;
                f2 += t * t;
            }
            f2 = block_sum_d(f2, sm.red);
            if (!(f2 < 1.0 / (MARG_EPS * MARG_EPS)) || !isfinite(f2)) { __syncthreads(); if (tid == 0) fast_ok = 0; }
            __syncthreads();
        }
        if (fast_ok) {
            // Z = Amm^-1 X, X = [Amr | bmm]:  T = X_p - Gp X_l ; Y_p = S^-1 T ; Y_l = D^-1 (X_l - C^T Y_p)
            for (int e = tid; e < mc * (n + 1); e += T) {
                const int r = e / (n + 1), c = e - r * (n + 1);
                double t = (c < n) ? A[(size_t)r * pos + m + c] : bv[r];
                for (int l = 0; l < L0; l++) t -= Gp[(size_t)r * L0 + l] * ((c < n) ? A[(size_t)(mc + l) * pos + m + c] : bv[mc + l]);
                Tm[e] = t;
            }
            __syncthreads();
            for (int e = tid; e < mc * (n + 1); e += T) {
                const int r = e / (n + 1), c = e - r * (n + 1);
                double t = 0;
                for (int k = 0; k < mc; k++) t += Sinv[r * mc + k] * Tm[(size_t)k * (n + 1) + c];
                Zm[e] = t;
            }
            __syncthreads();
            for (int e = tid; e < L0 * (n + 1); e += T) {
                const int l = e / (n + 1), c = e - l * (n + 1);
                double t = (c < n) ? A[(size_t)(mc + l) * pos + m + c] : bv[mc + l];
                for (int k = 0; k < mc; k++) t -= A[(size_t)k * pos + mc + l] * Zm[(size_t)k * (n + 1) + c];
                Zm[(size_t)(mc + l) * (n + 1) + c] = t / A[(size_t)(mc + l) * pos + mc + l];
            }
            __syncthreads();
        }
    }
    if (!fast_ok) {
    if (s.eig_mode) eig_sym_ql(A, m, pos, Vm, sm.de, sm.rot, sm.red, sm.ql_i, s.prof + (size_t)blockIdx.x * 32); else eig_sym_jacobi(A, m, pos, Vm, sm.cs, sm.pq, sm.red);
    __syncthreads();
    // Tm = Lambda^+ Vm^T [Amr | bmm]      (m x (n+1))
    for (int e = tid; e < m * (n + 1); e += T) {
        const int k = e / (n + 1), c = e - k * (n + 1);
        const double lam = A[(size_t)k * pos + k];
        double t = 0;
        if (lam > MARG_EPS) {
            for (int i = 0; i < m; i++) t += Vm[(size_t)i * m + k] * (c < n ? A[(size_t)i * pos + m + c] : bv[i]);
            t /= lam;
        }
        Tm[e] = t;
    }
    __syncthreads();
    BE_PROF(10);
    // Z = Vm Tm = Amm^+ [Amr | bmm]   (m x (n+1)) ;  A_r = Arr - Amr^T Z ,  b_r = brr - Amr^T z_b
    for (int e = tid; e < m * (n + 1); e += T) {
        const int i = e / (n + 1), c = e - i * (n + 1);
        double t = 0;
        for (int k = 0; k < m; k++) t += Vm[(size_t)i * m + k] * Tm[(size_t)k * (n + 1) + c];
        Zm[e] = t;
    }
    __syncthreads();
    }   // !fast_ok
    for (int e = tid; e < n * (n + 1); e += T) {
        const int r = e / (n + 1), c = e - r * (n + 1);
        double acc = 0;
        for (int i = 0; i < m; i++) acc += A[(size_t)i * pos + m + r] * Zm[(size_t)i * (n + 1) + c];
        if (c < n) Ar[(size_t)r * n + c] = A[(size_t)(m + r) * pos + m + c] - acc;
        else br[r] = bv[m + r] - acc;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += T) { const int i = e / n, j = e - i * n; if (j < i) { const double v = 0.5 * (Ar[(size_t)i * n + j] + Ar[(size_t)j * n + i]); Ar[(size_t)i * n + j] = v; Ar[(size_t)j * n + i] = v; } }
    __syncthreads();
    BE_PROF(11);
    // ---- the new prior ---------------------------------------------------------------------------------------------------
    // The reference factors A_r = V S V^T, drops eigenvalues <= 1e-8 and hands Ceres J0 = S^1/2 V^T, r0 = S^-1/2 V^T b_r
    // (marginalization_factor.cpp:290-315).  In information form that is Hp = V S+ V^T, bp = V+ V+^T b_r, c0 = b_r^T A_r^+ b_r.
    // DIRECT mode (default): the dropped directions are the gauge null space (|lambda| at the round-off level of A_r, b_r has no
    // component along them), so Hp = A_r and bp = b_r up to that round-off, and c0 -- a constant of the cost -- comes from one
    // Cholesky solve of the Jacobi-scaled, 1e-10-shifted A_r.  If that factorisation fails the exact eigen path below runs.
    int sweeps = 0;
    bool direct_ok = false;
    double c0 = 0.0;
    if (s.marg_direct) {
        // packed matrix in shared memory when it fits the kernel's dynamic allocation (2 MARG_NCAP^2 doubles), else in the (free) Vr scratch
        double *sP = ((size_t)(n + 1) * (n + 2) / 2 <= (size_t)2 * MARG_NCAP * MARG_NCAP) ? reinterpret_cast<double *>(smraw + ((sizeof(MargSmem) + 15) & ~(size_t)15)) : Vr;
        double *dinv = sm.de, *dsc = sm.de + 400, *rv = sm.rot, *yv = sm.rot + 400;        // n <= 400 (VIO_MAX_WIN = 24: NPX = 381)
        for (int i = tid; i < n; i += T) { const double d = Ar[(size_t)i * n + i]; dsc[i] = d > 0 ? rsqrt(d) : 1.0; }
        __syncthreads();
        for (int i = tid >> 5; i < n; i += T >> 5)
            for (int j = tid & 31; j <= i; j += 32) sP[pidx(i, j)] = Ar[(size_t)i * n + j] * dsc[i] * dsc[j] + (i == j ? 1e-10 : 0.0);
        for (int i = tid; i < n; i += T) rv[i] = br[i] * dsc[i];
        __syncthreads();
        direct_ok = chol_solve_packed(sP, n, rv, yv, &sm.ql_i[0], dinv, sm.prof_sink);
        if (direct_ok) {
            // two steps of iterative refinement against the UNSHIFTED matrix: the component of the solution along an eigenvalue lambda of
            // the scaled A_r converges like (1e-10 / lambda)^k, so c0 = b_r^T A_r^+ b_r is exact for every direction the reference keeps
            double *ev = sP + pidx(n, 0);                              // the border row is free after the solve
            for (int it = 0; it < 2; it++) {
                for (int i = tid; i < n; i += T) {
                    double t = 0;
                    for (int j = 0; j < n; j++) t += Ar[(size_t)j * n + i] * (dsc[j] * yv[j]);      // A_r symmetric: column walk, coalesced
                    ev[i] = rv[i] - dsc[i] * t;
                }
                __syncthreads();
                chol_forward_packed(sP, n, ev, dinv);
                chol_backward_packed(sP, n, ev, dinv);
                for (int i = tid; i < n; i += T) yv[i] += ev[i];
                __syncthreads();
            }
            double t = 0;
            for (int i = tid; i < n; i += T) t += rv[i] * yv[i];
            c0 = block_sum_d(t, sm.red);
            direct_ok = isfinite(c0);
        }
        sweeps = direct_ok ? -1 : 0;
        __syncthreads();
    }
    if (!direct_ok) {
    if (n <= MARG_NCAP) {                                      // A_r and V in shared memory
        double *sA = reinterpret_cast<double *>(smraw + ((sizeof(MargSmem) + 15) & ~(size_t)15));
        double *sV = sA + (size_t)n * n;
        for (int e = tid; e < n * n; e += T) sA[e] = Ar[e];
        __syncthreads();
        sweeps = s.eig_mode ? eig_sym_ql(sA, n, n, sV, sm.de, sm.rot, sm.red, sm.ql_i, s.prof + (size_t)blockIdx.x * 32) : eig_sym_jacobi(sA, n, n, sV, sm.cs, sm.pq, sm.red);
        Ar = sA; Vr = sV;
    } else {
        sweeps = s.eig_mode ? eig_sym_ql(Ar, n, n, Vr, sm.de, sm.rot, sm.red, sm.ql_i, s.prof + (size_t)blockIdx.x * 32) : eig_sym_jacobi(Ar, n, n, Vr, sm.cs, sm.pq, sm.red);
    }
    }
    if (tid == 0) { iv[IV_MARG_FAST] = fast_ok; iv[IV_MARG_SWEEPS] = sweeps; iv[IV_MARG_M] = m; }
    __syncthreads();
    BE_PROF(12);
    if (!direct_ok) {
    // tv[k] = v_k . b_r ;  c0 = sum_{lam>eps} tv^2 / lam
    for (int k = tid; k < n; k += T) {
        double t = 0;
        for (int i = 0; i < n; i++) t += Vr[(size_t)i * n + k] * br[i];
        tv[k] = t;
    }
    __syncthreads();
    double c0p = 0;
    for (int k = tid; k < n; k += T) { const double lam = Ar[(size_t)k * n + k]; if (lam > MARG_EPS) c0p += tv[k] * tv[k] / lam; }
    c0 = block_sum_d(c0p, sm.red);
    }
    // ---- write the new prior in canonical layout, shifted like addr_shift (VINS.cpp:759-774 / 806-829) ---------------
    double *Hp = s.Hp + (size_t)b * NPX * NPX, *bp = s.bp + (size_t)b * NPX;
    for (int e = tid; e < NPX * NPX; e += T) Hp[e] = 0.0;
    for (int e = tid; e < NPX; e += T) bp[e] = 0.0;
    __syncthreads();
    auto shift = [&](int c) {
        if (c >= NP) return c;                                            // para_Ex_Pose stays
        if (marg == 0) return c - 15;                                     // frame i -> i-1
        return c >= 15 * W ? c - 15 : c;                                  // frame W -> W-1
    };
    if (direct_ok) {
        for (int e = tid; e < n * n; e += T) {
            const int i = e / n, j = e - i * n;
            Hp[(size_t)shift(sm.kept[i]) * NPX + shift(sm.kept[j])] = Ar[e];
        }
        for (int i = tid; i < n; i += T) bp[shift(sm.kept[i])] = br[i];
    } else {
    for (int e = tid; e < n * n; e += T) {
        const int i = e / n, j = e - i * n;
        double t = 0;
        for (int k = 0; k < n; k++) { const double lam = Ar[(size_t)k * n + k]; if (lam > MARG_EPS) t += Vr[(size_t)i * n + k] * lam * Vr[(size_t)j * n + k]; }
        Hp[(size_t)shift(sm.kept[i]) * NPX + shift(sm.kept[j])] = t;
    }
    for (int i = tid; i < n; i += T) {
        double t = 0;
        for (int k = 0; k < n; k++) { const double lam = Ar[(size_t)k * n + k]; if (lam > MARG_EPS) t += Vr[(size_t)i * n + k] * tv[k]; }
        bp[shift(sm.kept[i])] = t;
    }
    }
    // linearisation point = current values of every kept block (preMarginalize memcpy), re-addressed
    double *x0 = s.x0 + (size_t)b * (NF * 16 + 7);
    __syncthreads();
    if (tid == 0) {
        int np[2 * (VIO_MAX_WIN + 1) + 1];
        for (int i = 0; i < 2 * NF + 1; i++) np[i] = 0;
        for (int blk = 0; blk < 2 * NF; blk++) {
            if (!touched[blk]) continue;
            const int fr = blk / 2;
            int nf2;
            if (marg == 0) { if (fr == 0) continue; nf2 = fr - 1; }
            else { if (blk == 2 * (W - 1)) continue; nf2 = (fr == W) ? W - 1 : fr; }
            np[2 * nf2 + (blk & 1)] = 1;
        }
        np[2 * NF] = touched[2 * NF];
        for (int fr = 0; fr < NF; fr++) {
            int src;
            if (marg == 0) src = fr + 1; else src = (fr == W - 1) ? W : fr;
            if (src > W) continue;
            if (np[2 * fr]) for (int k = 0; k < 7; k++) x0[16 * fr + k] = par[16 * src + k];
            if (np[2 * fr + 1]) for (int k = 0; k < 9; k++) x0[16 * fr + 7 + k] = par[16 * src + 7 + k];
        }
        for (int i = 0; i < 2 * NF + 1; i++) pres[i] = np[i];
        iv[IV_PRIOR_VALID] = 1; iv[IV_PRIOR_N] = n;
        BE_PROF_ONLY({ const long long _t = clock64(); _pp[13] += _t - _pt0; })
        dvs[DV_PRIOR_C0] = c0;
    }
}

__host__ inline size_t marg_scratch_doubles(int NPX, int LCAP, int max_cnt) {
    const size_t m = 15 + (size_t)max_cnt, n = NPX, pos = m + n;
    return (LCAP + 7) / 2 + pos * pos + pos + m * m + 2 * m * (n + 1) + n * n + n + n * n + 2 * (size_t)NPX + 1024;
}

}  // namespace be
