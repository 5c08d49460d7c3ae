// be_align.cuh -- visual-inertial alignment of the initialisation (the pure linear-algebra half), f64.
// One CTA per stream.  Reference (paths under /root/reference/VINS_ios/):
//   VisualIMUAlignment     initial_aligment.cpp:222-229
//   solveGyroscopeBias     initial_aligment.cpp:10-46     (3 x 3 normal equations over consecutive frames, LDLT, repropagate)
//   SolveScale             initial_aligment.cpp:135-220   (velocities of every frame, gravity, scale: 3 n + 4 unknowns, LDLT)
//   TangentBasis           initial_aligment.cpp:49-62
//   RefineGravity          initial_aligment.cpp:64-133    (4 iterations on the gravity tangent plane; note that A and b are NOT
//                                                          cleared between the iterations -- they keep accumulating, scaled by 1000
//                                                          each time -- which is reproduced here)
//   A.ldlt().solve(b)      Eigen 3.3 Cholesky/LDLT.h:277-374 (unblocked, diagonal pivoting, lower) and :545-580 (solve)
//   IntegrationBase::repropagate                          integration_base.h:46-61 (be_factors.cuh: pre_init + pre_propagate_warp)
#pragma once
#include "be_kernels.cuh"

namespace be {

constexpr int ALIGN_THREADS = 128;

struct AlignArgs {
    int B, F, MAXIMU, NS;                       // NS = 3 F + 4: leading dimension of the normal matrix
    const int *n_frames, *counts;               // [B], [B][F]
    const double *R, *T, *imu0, *imu, *bg0;     // [B][F][9], [B][F][3], [B][F][imu0_stride] (acc_0, gyr_0 first), [B][F][MAXIMU][7], [B][bg0_stride]
    int imu0_stride, bg0_stride;                // 6 and 3 for the stand-alone entry; PR_STRIDE and 3 NF when the arrays are the back end's own
    const double *abg; int abg_stride;          // optional [B][F][abg_stride]: bias each frame's FIRST integration uses (nullptr: bg0 for all)
    double tic[3], g_norm, g_thr, noise[6];
    double *pre, *A, *Aw, *rhs, *pairs, *xs;    // scratch: [B][F][PR_STRIDE], [B][NS NS], [B][NS NS], [B][NS], [B][F][110], [B][2 NS]
    int *perm;                                  // [B][NS] transpositions
    double *bgs_out, *g_out, *x_out;            // [B][3], [B][3], [B][NS]
    int *ok;                                    // [B]
};

// ---- LDLT with diagonal pivoting, lower triangle, in place; then the solve.  Whole CTA. ---------------------------------------
__device__ inline void align_ldlt_solve(double *M, int ld, int n, const double *b, double *x, double *temp, int *trn) {
    __shared__ double red_v[ALIGN_THREADS / 32];
    __shared__ int red_i[ALIGN_THREADS / 32];
    __shared__ int sh_idx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < n; k++) {
        // index of the largest |diagonal| in the trailing corner (first one on ties, as maxCoeff does)
        double bv = -1.0; int bi = n;
        for (int i = k + tid; i < n; i += ALIGN_THREADS) { const double v = fabs(M[(size_t)i * ld + i]); if (v > bv) { bv = v; bi = i; } }
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < ALIGN_THREADS / 32; w++) if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
            sh_idx = bi; trn[k] = bi;
        }
        __syncthreads();
        const int p = sh_idx;
        if (p != k) {                                          // symmetric swap of k and p within the lower triangle
            const int s = n - p - 1;
            for (int c = tid; c < k; c += ALIGN_THREADS) { const double t = M[(size_t)k * ld + c]; M[(size_t)k * ld + c] = M[(size_t)p * ld + c]; M[(size_t)p * ld + c] = t; }
            for (int r = tid; r < s; r += ALIGN_THREADS) {
                const size_t rr = (size_t)(p + 1 + r) * ld;
                const double t = M[rr + k]; M[rr + k] = M[rr + p]; M[rr + p] = t;
            }
            for (int i = k + 1 + tid; i < p; i += ALIGN_THREADS) { const double t = M[(size_t)i * ld + k]; M[(size_t)i * ld + k] = M[(size_t)p * ld + i]; M[(size_t)p * ld + i] = t; }
            if (tid == 0) { const double t = M[(size_t)k * ld + k]; M[(size_t)k * ld + k] = M[(size_t)p * ld + p]; M[(size_t)p * ld + p] = t; }
            __syncthreads();
        }
        if (k > 0) {
            for (int c = tid; c < k; c += ALIGN_THREADS) temp[c] = M[(size_t)c * ld + c] * M[(size_t)k * ld + c];
            __syncthreads();
            for (int r = k + tid; r < n; r += ALIGN_THREADS) {
                const double *row = M + (size_t)r * ld;
                double d = 0.0;
                for (int c = 0; c < k; c++) d += row[c] * temp[c];
                M[(size_t)r * ld + k] -= d;
            }
            __syncthreads();
        }
        const double akk = M[(size_t)k * ld + k];
        if (fabs(akk) > 0.0)
            for (int r = k + 1 + tid; r < n; r += ALIGN_THREADS) M[(size_t)r * ld + k] /= akk;
        __syncthreads();
    }
    // x = P b
    for (int i = tid; i < n; i += ALIGN_THREADS) x[i] = b[i];
    __syncthreads();
    if (tid == 0) for (int k = 0; k < n; k++) { const int p = trn[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
    __syncthreads();
    for (int c = 0; c < n; c++) {                              // L^-1
        const double xc = x[c];
        for (int r = c + 1 + tid; r < n; r += ALIGN_THREADS) x[r] -= M[(size_t)r * ld + c] * xc;
        __syncthreads();
    }
    for (int i = tid; i < n; i += ALIGN_THREADS) {             // D^-1 (pseudo-inverse with Eigen's tolerance 1 / highest)
        const double d = M[(size_t)i * ld + i];
        x[i] = fabs(d) > 1.0 / 1.7976931348623157e308 ? x[i] / d : 0.0;
    }
    __syncthreads();
    for (int c = n - 1; c >= 0; c--) {                         // L^-T
        const double xc = x[c];
        for (int r = tid; r < c; r += ALIGN_THREADS) x[r] -= M[(size_t)c * ld + r] * xc;
        __syncthreads();
    }
    if (tid == 0) for (int k = n - 1; k >= 0; k--) { const int p = trn[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
    __syncthreads();
}

// pre-integrate (or re-propagate) every frame's interval with gyroscope bias bg; warps take frames round-robin
__device__ inline void align_integrate(const AlignArgs &a, int b, int n, V3 bg, bool per_frame_bias, PreScratch *scr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < n; k += ALIGN_THREADS / 32) {
        double *pr = a.pre + ((size_t)b * a.F + k) * PR_STRIDE;
        const double *i0 = a.imu0 + ((size_t)b * a.F + k) * a.imu0_stride;
        const V3 bgk = (per_frame_bias && a.abg) ? ld3(a.abg + ((size_t)b * a.F + k) * a.abg_stride) : bg;
        if (lane == 0) pre_init(pr, ld3(i0), ld3(i0 + 3), v3(0, 0, 0), bgk);
        __syncwarp();
        const int cnt = min(a.counts[(size_t)b * a.F + k], a.MAXIMU);
        const double *e = a.imu + ((size_t)b * a.F + k) * a.MAXIMU * 7;
        for (int i = 0; i < cnt; i++) pre_propagate_warp(pr, e[7 * i], ld3(e + 7 * i + 1), ld3(e + 7 * i + 4), a.noise, scr[warp]);
    }
    __syncthreads();
}

// r_A = tmp_A^T tmp_A (nc x nc, nc = 9 or 10) and r_b = tmp_A^T tmp_b of one consecutive pair, into pb[110] (r_A row-major with stride
// 10, r_b at 100).  lxly == nullptr: SolveScale's 6 x 10 block;  else RefineGravity's 6 x 9 block around g0.
__device__ inline void align_pair_block(const AlignArgs &a, int b, int i, const double *lxly, V3 g0, double *pb) {
    const size_t fi = (size_t)b * a.F + i, fj = fi + 1;
    const M3 Ri = ldm(a.R + fi * 9), Rj = ldm(a.R + fj * 9);
    const V3 Ti = ld3(a.T + fi * 3), Tj = ld3(a.T + fj * 3);
    const double *pj = a.pre + fj * PR_STRIDE;
    const double dt = pj[PR_SUMDT];
    const M3 RiT = tr(Ri);
    const M3 Rij = RiT * Rj;
    const V3 tic = v3(a.tic[0], a.tic[1], a.tic[2]);
    const int nc = lxly ? 9 : 10;
    double A[6][10], bb[6];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 10; c++) A[r][c] = 0.0;
    M3 Rh, Rd;                                                     // R_i^T dt dt / 2  and  R_i^T dt
    for (int e = 0; e < 9; e++) { Rh.m[e] = RiT.m[e] * dt * dt / 2; Rd.m[e] = RiT.m[e] * dt; }
    const V3 dT = RiT * (Tj - Ti);
    V3 b0 = ld3(pj + PR_DP) + Rij * tic - tic, b1 = ld3(pj + PR_DV);
    for (int r = 0; r < 3; r++) {
        A[r][r] = -dt;
        A[3 + r][r] = -1.0;
        for (int c = 0; c < 3; c++) A[3 + r][3 + c] = Rij.m[3 * r + c];
    }
    if (!lxly) {
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A[r][6 + c] = Rh.m[3 * r + c]; A[3 + r][6 + c] = Rd.m[3 * r + c]; }
        A[0][9] = dT.x / 100.0; A[1][9] = dT.y / 100.0; A[2][9] = dT.z / 100.0;
    } else {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 2; c++) {
                A[r][6 + c] = Rh.m[3 * r] * lxly[c] + Rh.m[3 * r + 1] * lxly[2 + c] + Rh.m[3 * r + 2] * lxly[4 + c];
                A[3 + r][6 + c] = Rd.m[3 * r] * lxly[c] + Rd.m[3 * r + 1] * lxly[2 + c] + Rd.m[3 * r + 2] * lxly[4 + c];
            }
        A[0][8] = dT.x / 100.0; A[1][8] = dT.y / 100.0; A[2][8] = dT.z / 100.0;
        b0 = b0 - Rh * g0;
        b1 = b1 - Rd * g0;
    }
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b1.x; bb[4] = b1.y; bb[5] = b1.z;
    for (int r = 0; r < nc; r++) {
        for (int c = 0; c < nc; c++) { double s = 0.0; for (int k = 0; k < 6; k++) s += A[k][r] * A[k][c]; pb[r * 10 + c] = s; }
        double s = 0.0;
        for (int k = 0; k < 6; k++) s += A[k][r] * bb[k];
        pb[100 + r] = s;
    }
}

// A (lower triangle) += sum over pairs in ascending order of the scattered r_A blocks, b likewise; then both times 1000.
// m = number of shared tail unknowns (4 or 3); ns = 3 n + m.
__device__ inline void align_gather(const AlignArgs &a, int b, int n, int m, int ns, double *A, double *rhs) {
    const double *pairs = a.pairs + (size_t)b * a.F * 110;
    auto loc = [&](int g, int i) -> int {          // local index of global unknown g in pair i's block, or -1
        if (g >= ns - m) return 6 + g - (ns - m);
        const int o = g - 3 * i;
        return (o >= 0 && o < 6) ? o : -1;
    };
    for (int e = threadIdx.x; e < ns * (ns + 1) / 2 + ns; e += ALIGN_THREADS) {
        int r, c;
        const bool is_b = e >= ns * (ns + 1) / 2;
        if (is_b) { r = e - ns * (ns + 1) / 2; c = r; }
        else { r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5); while (r * (r + 1) / 2 > e) r--; while ((r + 1) * (r + 2) / 2 <= e) r++; c = e - r * (r + 1) / 2; }
        const int lowg = min(r, c);
        int ilo = 0, ihi = n - 2;
        if (lowg < ns - m) { ilo = max(0, lowg / 3 - 1); ihi = min(n - 2, lowg / 3); }
        double acc = is_b ? rhs[r] : A[(size_t)r * a.NS + c];
        for (int i = ilo; i <= ihi; i++) {
            const int lr = loc(r, i), lc = loc(c, i);
            if (lr < 0 || lc < 0) continue;
            acc += is_b ? pairs[(size_t)i * 110 + 100 + lr] : pairs[(size_t)i * 110 + lr * 10 + lc];
        }
        acc = acc * 1000.0;
        if (is_b) rhs[r] = acc; else A[(size_t)r * a.NS + c] = acc;
    }
    __syncthreads();
}

// VisualIMUAlignment for stream b with n frames; whole CTA; results in a.bgs_out / g_out / x_out / ok.  Returns ok.
__device__ inline int align_core(const AlignArgs &a, int b, int n, PreScratch *scr, double *sh) {
    const int tid = threadIdx.x;
    double *A = a.A + (size_t)b * a.NS * a.NS, *Aw = a.Aw + (size_t)b * a.NS * a.NS, *rhs = a.rhs + (size_t)b * a.NS;
    double *x = a.xs + (size_t)b * 2 * a.NS, *temp = x + a.NS;
    int *trn = a.perm + (size_t)b * a.NS;
    double *pairs = a.pairs + (size_t)b * a.F * 110;
    if (tid == 0) a.ok[b] = 0;
    if (n < 2) return 0;
    V3 bg = ld3(a.bg0 + (size_t)b * a.bg0_stride);
    align_integrate(a, b, n, bg, true, scr);

    // ---- solveGyroscopeBias
    if (tid == 0) {
        double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gb[3] = {0, 0, 0};
        for (int i = 0; i + 1 < n; i++) {
            const size_t fi = (size_t)b * a.F + i, fj = fi + 1;
            const double *pj = a.pre + fj * PR_STRIDE;
            const Q4 qij = R2q(tr(ldm(a.R + fi * 9)) * ldm(a.R + fj * 9));
            const Q4 e = qmul(qinv(ldq(pj + PR_DQ)), qij);
            const double tb[3] = {2 * e.x, 2 * e.y, 2 * e.z};
            double J[9];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) J[3 * r + c] = pj[PR_JAC + (3 + r) * 15 + 12 + c];
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) { double s = 0; for (int k = 0; k < 3; k++) s += J[3 * k + r] * J[3 * k + c]; G[3 * r + c] += s; }
                double s = 0; for (int k = 0; k < 3; k++) s += J[3 * k + r] * tb[k];
                gb[r] += s;
            }
        }
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Aw[(size_t)r * a.NS + c] = G[3 * r + c]; rhs[r] = gb[r]; }
    }
    __syncthreads();
    align_ldlt_solve(Aw, a.NS, 3, rhs, x, temp, trn);
    bg = bg + v3(x[0], x[1], x[2]);
    __syncthreads();
    if (tid == 0) st3(a.bgs_out + 3 * b, bg);
    align_integrate(a, b, n, bg, false, scr);

    // ---- SolveScale
    int ns = 3 * n + 4;
    for (int e = tid; e < ns * a.NS; e += ALIGN_THREADS) A[e] = 0.0;
    for (int e = tid; e < ns; e += ALIGN_THREADS) rhs[e] = 0.0;
    for (int i = tid; i + 1 < n; i += ALIGN_THREADS) align_pair_block(a, b, i, nullptr, v3(0, 0, 0), pairs + (size_t)i * 110);
    __syncthreads();
    align_gather(a, b, n, 4, ns, A, rhs);
    for (int e = tid; e < ns * a.NS; e += ALIGN_THREADS) Aw[e] = A[e];
    __syncthreads();
    align_ldlt_solve(Aw, a.NS, ns, rhs, x, temp, trn);
    double *xo = a.x_out + (size_t)b * a.NS;
    for (int e = tid; e < a.NS; e += ALIGN_THREADS) xo[e] = e < ns ? x[e] : 0.0;
    V3 g = v3(x[ns - 4], x[ns - 3], x[ns - 2]);
    double sc = x[ns - 1] / 100.0;
    __syncthreads();
    if (fabs(norm(g) - a.g_norm) > a.g_thr || sc < 0) { if (tid == 0) st3(a.g_out + 3 * b, g); __syncthreads(); return 0; }

    // ---- RefineGravity
    ns = 3 * n + 3;
    V3 g0;                                   // g.normalized() * G_NORM
    {
        const double ng = norm(g);
        g0 = v3(g.x / ng * a.g_norm, g.y / ng * a.g_norm, g.z / ng * a.g_norm);
    }
    for (int e = tid; e < ns * a.NS; e += ALIGN_THREADS) A[e] = 0.0;
    for (int e = tid; e < ns; e += ALIGN_THREADS) rhs[e] = 0.0;
    __syncthreads();
    for (int it = 0; it < 4; it++) {
        if (tid == 0) {                            // TangentBasis(g0) -> sh[0..5] = lxly (3 x 2 row-major)
            const double ng = norm(g0);
            const V3 av = v3(g0.x / ng, g0.y / ng, g0.z / ng);
            V3 t = v3(0, 0, 1);
            if (av.x == 0.0 && av.y == 0.0 && av.z == 1.0) t = v3(1, 0, 0);
            V3 bvec = t - dot(av, t) * av;
            const double nb = norm(bvec);
            bvec = v3(bvec.x / nb, bvec.y / nb, bvec.z / nb);
            const V3 cv = cross(av, bvec);
            sh[0] = bvec.x; sh[1] = cv.x; sh[2] = bvec.y; sh[3] = cv.y; sh[4] = bvec.z; sh[5] = cv.z;
        }
        __syncthreads();
        for (int i = tid; i + 1 < n; i += ALIGN_THREADS) align_pair_block(a, b, i, sh, g0, pairs + (size_t)i * 110);
        __syncthreads();
        align_gather(a, b, n, 3, ns, A, rhs);
        for (int e = tid; e < ns * a.NS; e += ALIGN_THREADS) Aw[e] = A[e];
        __syncthreads();
        align_ldlt_solve(Aw, a.NS, ns, rhs, x, temp, trn);
        const double d0 = x[ns - 3], d1 = x[ns - 2];
        V3 gn = g0 + v3(sh[0] * d0 + sh[1] * d1, sh[2] * d0 + sh[3] * d1, sh[4] * d0 + sh[5] * d1);
        const double ng = norm(gn);
        g0 = v3(gn.x / ng * a.g_norm, gn.y / ng * a.g_norm, gn.z / ng * a.g_norm);
        __syncthreads();
    }
    sc = x[ns - 1] / 100.0;
    for (int e = tid; e < a.NS; e += ALIGN_THREADS) xo[e] = e < ns - 1 ? x[e] : (e == ns - 1 ? sc : 0.0);
    if (tid == 0) { st3(a.g_out + 3 * b, g0); a.ok[b] = sc > 0.0 ? 1 : 0; }
    __syncthreads();
    return sc > 0.0 ? 1 : 0;
}

__global__ void __launch_bounds__(ALIGN_THREADS) visual_imu_align_kernel(AlignArgs a) {
    __shared__ PreScratch scr[ALIGN_THREADS / 32];
    __shared__ double sh[16];
    align_core(a, blockIdx.x, min(a.n_frames[blockIdx.x], a.F), scr, sh);
}

// ---- VINS::visualInitialAlign (VINS.cpp:1022-1102) on the back end's own state --------------------------------------------------
// Runs before triangulate_kernel when the host has supplied the SfM poses of the window's frames (vio_backend_set_init_sfm,
// IV_INIT_PENDING == 2) and the feature kernel decided ACT_INIT_SOLVE.  The alignment runs over all_image_frame as the back end keeps it
// (the af_* records of be_state.cuh: every camera frame since the stream started, keyframe or not, with its own IMU interval); the
// window's frames are the map's keyframes, found by their headers.
//   alignment fails  -> Bgs keep the corrected bias (solveGyroscopeBias has already added it), action becomes ACT_SLIDE_ONLY
//   alignment passes -> Ps / Rs from the SfM, depths re-triangulated on the camera poses (tic = 0), pre-integrations re-propagated with the
//                       new Bgs, metric scale, velocities, gravity-aligned frame; IV_INIT_PENDING = 3 tells triangulate_kernel that the window
//                       is ready and the solve follows as in the reference (VINS.cpp:415-447).
__device__ inline M3 g2R_dev(V3 g) {                  // Utility::g2R, utility.cpp:8-19 (Quaterniond::FromTwoVectors(ng1, e_z))
    const double ng = norm(g);
    const V3 v0 = v3(g.x / ng, g.y / ng, g.z / ng), v1 = v3(0, 0, 1);
    const double c = dot(v1, v0);
    Q4 q;
    if (c < -1.0 + 1e-12) q = q4(1, 0, 0, 0);         // antiparallel: Eigen picks an orthogonal axis by SVD; any half turn about one maps g to +z
    else {
        const V3 axis = cross(v0, v1);
        const double sq = sqrt((1.0 + c) * 2.0), invs = 1.0 / sq;
        q = q4(axis.x * invs, axis.y * invs, axis.z * invs, sq * 0.5);
    }
    M3 R0 = q2R(q);
    const double yaw = R2ypr(R0).x;
    R0 = ypr2R(v3(-yaw, 0, 0)) * R0;
    R0 = ypr2R(v3(-90, 0, 0)) * R0;
    return R0;
}

__global__ void __launch_bounds__(ALIGN_THREADS) init_align_kernel(BeState s, AlignArgs a) {
    __shared__ PreScratch scr[ALIGN_THREADS / 32];
    __shared__ double sh[16];
    __shared__ int sh_bad;
    __shared__ int key[VIO_MAX_WIN + 1];              // all_image_frame index of window frame i (the keyframes of the map)
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int *iv = S_iv(s, b);
    if (iv[IV_ACTION] != ACT_INIT_SOLVE || iv[IV_INIT_PENDING] != 2) return;
    const int n = iv[IV_AF_N], W = s.W, NF = s.NF;
    if (tid == 0) {
        const int *host_n = reinterpret_cast<const int *>(s.init_sfm + (size_t)s.B * s.FA * 12);
        int bad = 0;
        if (n < 2 || n > s.FA || host_n[b] != n) bad = 1;          // the caller's frame list is not the map's (or the map overflowed)
        for (int i = 0, k = 0; i <= W && !bad; i++) {               // Headers ascend, so does the map: one forward scan
            while (k < n && s.af_hdr[(size_t)b * s.FA + k] != s.Headers[(size_t)b * NF + i]) k++;
            if (k >= n) bad = 1; else key[i] = k;
        }
        for (int i = 0; i < NF && !bad; i++) if (S_pre(s, b, i)[PR_VALID] == 0.0) bad = 2;   // a frame never saw an IMU sample
        sh_bad = bad;
    }
    __syncthreads();
    if (sh_bad) {
        if (tid == 0) { if (sh_bad == 1) iv[IV_ERR] = VIO_ERR_STATE; iv[IV_ALIGN_OK] = 0; iv[IV_ACTION] = ACT_SLIDE_ONLY; iv[IV_INIT_PENDING] = 0; }
        return;
    }
    const int ok = align_core(a, b, n, scr, sh);
    const V3 bgs = ld3(a.bgs_out + 3 * b);
    const V3 dbg = bgs - ld3(S_Bgs(s, b, 0));
    __syncthreads();
    for (int i = tid; i < NF; i += ALIGN_THREADS) st3(S_Bgs(s, b, i), ld3(S_Bgs(s, b, i)) + dbg);        // Bgs[i] += delta_bg
    for (int k = tid; k < n; k += ALIGN_THREADS) st3(s.af_abg + ((size_t)b * s.FA + k) * 3, bgs);       // the map's pre-integrations now sit at Bgs[0]
    if (tid == 0) iv[IV_ALIGN_OK] = ok;
    if (!ok) {
        if (tid == 0) { iv[IV_ACTION] = ACT_SLIDE_ONLY; iv[IV_INIT_PENDING] = 0; double *dvo = S_dv(s, b); dvo[DV_INIT_SCALE] = 0.0; st3(dvo + DV_INIT_G, ld3(a.g_out + 3 * b)); }
        return;
    }
    const size_t fo = (size_t)b * s.FCAP;
    const int nf = iv[IV_NFEAT];
    for (int i = tid; i < NF; i += ALIGN_THREADS) {
        st3(S_Ps(s, b, i), ld3(a.T + ((size_t)b * a.F + key[i]) * 3));
        stm(S_Rs(s, b, i), ldm(a.R + ((size_t)b * a.F + key[i]) * 9));
    }
    for (int k = tid; k < nf; k += ALIGN_THREADS) s.f_depth[fo + k] = -1.0;          // clearDepth(-1)
    __syncthreads();
    triangulate_stream(s, b, v3(0, 0, 0), tid, ALIGN_THREADS);                        // "triangulat on cam pose, no tic"
    // pre_integrations[i]->repropagate(0, Bgs[i]) (integration_base.h:46-61) from the window's own sample buffers
    for (int i = warp; i < NF; i += ALIGN_THREADS / 32) {
        double *pr = S_pre(s, b, i);
        const V3 la = ld3(pr + PR_LIN_ACC), lg = ld3(pr + PR_LIN_GYR), bgi = ld3(S_Bgs(s, b, i));
        __syncwarp();
        if (lane == 0) pre_init(pr, la, lg, v3(0, 0, 0), bgi);
        __syncwarp();
        const int cnt = min(s.imu_cnt[(size_t)b * NF + i], s.MAXIMU);
        const double *e = S_imu(s, b, i);
        for (int k = 0; k < cnt; k++) pre_propagate_warp(pr, e[7 * k], ld3(e + 7 * k + 1), ld3(e + 7 * k + 4), s.noise, scr[warp]);
    }
    __syncthreads();
    const double *x = a.x_out + (size_t)b * a.NS;
    const double sc = x[3 * n + 2];
    const V3 tic = ld3(S_dv(s, b) + DV_TIC);
    if (tid == 0) {
        const V3 P0 = ld3(S_Ps(s, b, 0));
        const V3 off = sc * P0 - ldm(S_Rs(s, b, 0)) * tic;
        for (int i = W; i >= 0; i--) st3(S_Ps(s, b, i), sc * ld3(S_Ps(s, b, i)) - ldm(S_Rs(s, b, i)) * tic - off);
    }
    // Vs[kv] = R_keyframe * x.segment<3>(kv * 3): the reference indexes x by the KEYFRAME counter, not by the frame's place in the map
    // (VINS.cpp:1066-1075); identical when every frame is a keyframe, reproduced as written otherwise
    for (int i = tid; i < NF; i += ALIGN_THREADS) st3(S_Vs(s, b, i), ldm(S_Rs(s, b, i)) * v3(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
    for (int k = tid; k < nf; k += ALIGN_THREADS)
        if (in_solve(s, s.f_nobs[fo + k], s.f_start[fo + k])) s.f_depth[fo + k] *= sc;
    __syncthreads();
    const V3 g = ld3(a.g_out + 3 * b);
    M3 R0 = g2R_dev(g);
    const double yaw0 = R2ypr(R0).x;
    R0 = ypr2R(v3(-yaw0, 0, 0)) * R0;
    for (int i = tid; i < NF; i += ALIGN_THREADS) {
        st3(S_Ps(s, b, i), R0 * ld3(S_Ps(s, b, i)));
        stm(S_Rs(s, b, i), R0 * ldm(S_Rs(s, b, i)));
        st3(S_Vs(s, b, i), R0 * ld3(S_Vs(s, b, i)));
    }
    __syncthreads();
    if (tid == 0) { double *dvo = S_dv(s, b); dvo[DV_INIT_SCALE] = sc; st3(dvo + DV_INIT_G, R0 * g); iv[IV_INIT_PENDING] = 3; }
}

}  // namespace be
