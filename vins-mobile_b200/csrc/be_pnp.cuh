// be_pnp.cuh -- motion-only PnP tracker: FeatureTracker::solveVinsPnP (feature_tracker.cpp:107-160)
// -> vinsPnP (vins_pnp.cpp), batched, one CTA per stream.  A window of PNP_SIZE + 1 = 7 camera frames; landmarks are FIXED at the
// positions the sliding-window estimator solved, so a frame contributes one PerspectiveFactor (perspective_factor.cpp:16-67) per
// tracked landmark and consecutive frames are tied by IMUFactorPnP (imu_factor_pnp.h: IMUFactor with speed and bias as separate
// blocks; biases and the extrinsic pose are constant, frames whose estimator result has arrived (find_solved) are constant too).
// ceres::Solve (DENSE_SCHUR + DOGLEG, 5 iterations, CauchyLoss(1) on the perspective factors) is restated on the normal equations of
// the 63 tangent dofs [pose_i(6) speed_i(3)] exactly as be_solve.cuh restates it for the big window; the reference's 10 ms wall-time
// cap on the solve is not reproduced (it makes the result machine dependent; the oracle drops it too, oracle/pnp_ref.cpp).
#pragma once
#include "be_solve.cuh"

namespace be {

constexpr int PNP_N = 7;                 // PNP_SIZE + 1 (global_param.hpp:29)
constexpr int PNP_D = 9 * PNP_N;         // tangent dofs
constexpr int PNP_T = 256;

struct PnpState {
    int B, MAXF, max_iters;
    double gravity, sqrt_info;
    double noise[6], tic[3], ric[9];
    double *Ps, *Rs, *Vs, *Bas, *Bgs, *Headers;      // [B][7][3|9|3|3|3|1]
    int *iv;                                          // [B][8]: 0 frame_count, 1 first_imu, 2 err, 3 iterations of the last solve
    int *solved;                                      // [B][7] find_solved
    double *dv;                                       // [B][8]: acc_0 (3), gyr_0 (3), initial cost, final cost
    double *pre;                                      // [B][7][PR_STRIDE] pre-integrations
    int *f_n, *f_id, *f_track;                        // [B][7], [B][7][MAXF] x2
    double *f_obs, *f_pos;                            // [B][7][MAXF][2], [B][7][MAXF][3]
};
__device__ __forceinline__ double *P_pre(const PnpState &s, int b, int i) { return s.pre + ((size_t)b * PNP_N + i) * PR_STRIDE; }
__device__ __forceinline__ size_t P_f(const PnpState &s, int b, int i) { return ((size_t)b * PNP_N + i) * s.MAXF; }

// vinsPnP::setInit (vins_pnp.cpp:63-83)
__global__ void pnp_set_init_kernel(PnpState s, const double *header, const double *P, const double *R, const double *V, const double *Ba, const double *Bg) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    for (int i = 0; i < PNP_N; i++) {
        const size_t k = (size_t)b * PNP_N + i;
        st3(s.Bas + 3 * k, ld3(Ba + 3 * b)); st3(s.Bgs + 3 * k, ld3(Bg + 3 * b));
        if (s.Headers[k] == header[b]) {
            s.solved[k] = 1;
            st3(s.Ps + 3 * k, ld3(P + 3 * b)); stm(s.Rs + 9 * k, ldm(R + 9 * b)); st3(s.Vs + 3 * k, ld3(V + 3 * b));
        }
    }
}

// vinsPnP::processIMU (vins_pnp.cpp:197-233), n samples; one warp per stream
__global__ void __launch_bounds__(128) pnp_imu_kernel(PnpState s, int n_samples, const double *__restrict__ dts, const double *__restrict__ accs,
                                                      const double *__restrict__ gyrs) {
    __shared__ PreScratch scr[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= s.B) return;
    int *iv = s.iv + (size_t)b * 8;
    double *dv = s.dv + (size_t)b * 8;
    for (int n = 0; n < n_samples; n++) {
        const double dt = dts[(size_t)n * s.B + b];
        const V3 acc = ld3(accs + ((size_t)n * s.B + b) * 3), gyr = ld3(gyrs + ((size_t)n * s.B + b) * 3);
        const int fc = iv[0];
        if (lane == 0 && !iv[1]) { iv[1] = 1; st3(dv, acc); st3(dv + 3, gyr); }
        __syncwarp();
        const size_t k = (size_t)b * PNP_N + fc;
        double *pr = P_pre(s, b, fc);
        if (lane == 0 && pr[PR_VALID] == 0.0) pre_init(pr, ld3(dv), ld3(dv + 3), ld3(s.Bas + 3 * k), ld3(s.Bgs + 3 * k));
        __syncwarp();
        if (fc != 0) {
            pre_propagate_warp(pr, dt, acc, gyr, s.noise, scr[warp]);
            if (lane == 0) {
                const V3 g = v3(0, 0, s.gravity);
                const V3 a0 = ld3(dv), g0 = ld3(dv + 3), ba = ld3(s.Bas + 3 * k), bg = ld3(s.Bgs + 3 * k);
                M3 R = ldm(s.Rs + 9 * k);
                V3 P = ld3(s.Ps + 3 * k), V = ld3(s.Vs + 3 * k);
                const V3 un_acc_0 = R * (a0 - ba) - g;
                const V3 un_gyr = 0.5 * (g0 + gyr) - bg;
                R = R * q2R(deltaQ(un_gyr * dt));
                const V3 un_acc_1 = R * (acc - ba) - g;
                const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
                P = P + dt * V + (0.5 * dt * dt) * un_acc;
                V = V + dt * un_acc;
                stm(s.Rs + 9 * k, R); st3(s.Ps + 3 * k, P); st3(s.Vs + 3 * k, V);
            }
        }
        __syncwarp();
        if (lane == 0) { st3(dv, acc); st3(dv + 3, gyr); }
        __syncwarp();
    }
}

struct PnpWs {
    double *H, *Sx, *g, *sc, *dd, *gr, *gn, *st, *up, *rhs, *y, *par, *cand, *imu;
};
__host__ __device__ inline size_t pnp_smem_doubles() { return 2 * (size_t)PNP_D * PNP_D + 9 * 64 + 2 * 16 * PNP_N + 930 * (PNP_N - 1) + 8; }

// local column of the 15x30 IMU Jacobian [pose_i(6) v_i ba_i bg_i pose_j(6) v_j ba_j bg_j] -> tangent dof of frame f / f + 1 (-1: constant bias)
__device__ __forceinline__ int pnp_dof(int f, int c) {
    if (c < 9) return c < 6 ? 9 * f + c : 9 * f + c;             // pose_i 0..5 -> 9f + 0..5, v_i 6..8 -> 9f + 6..8
    if (c < 15) return -1;
    if (c < 24) return 9 * (f + 1) + (c - 15);                   // pose_j, v_j
    return -1;
}

// cost (and, when lin != 0, H = J^T J and g = J^T r over the 63 dofs; rows / columns of constant frames are reset by the caller)
__device__ __noinline__ double pnp_evaluate(const PnpState &s, int b, const PnpWs &ws, const double *par, int lin, double *sh_red) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    double cost = 0.0;
    if (lin) {
        for (int e = tid; e < PNP_D * PNP_D; e += T) ws.H[e] = 0.0;
        for (int e = tid; e < PNP_D; e += T) ws.g[e] = 0.0;
    }
    __syncthreads();
    // A residual block whose parameter blocks are ALL constant is removed from the program Ceres minimises (Program::RemoveFixedBlocks:
    // its cost becomes summary.fixed_cost and does not enter the cost the trust-region loop and its tolerances see).  Biases and the
    // extrinsic pose are always constant, so that is an IMU factor between two solved frames and every perspective factor of a solved frame.
    const int *svf = s.solved + (size_t)b * PNP_N;
    for (int f = warp; f < PNP_N - 1; f += nwarp) {              // IMUFactorPnP(pre_integrations[f + 1]) on frames f, f + 1
        if (svf[f] && svf[f + 1]) continue;
        const double *pr = P_pre(s, b, f + 1);
        double *J = ws.imu + (size_t)f * 930, *Jw = J + 450, *rr = J + 900, *rw = J + 915;
        imu_residual_warp(pr, s.gravity, par + 16 * f, par + 16 * f + 7, par + 16 * (f + 1), par + 16 * (f + 1) + 7, rr, lin ? J : nullptr, lane);
        __syncwarp();
        const double *U = pr + PR_SQI;
        if (lane < 15) { double r_w = 0.0; for (int k = lane; k < 15; k++) r_w += U[lane * 15 + k] * rr[k]; cost += 0.5 * r_w * r_w; rw[lane] = r_w; }
        __syncwarp();
        if (lin) {
            for (int e = lane; e < 450; e += 32) {
                const int r = e / 30, c = e - r * 30;
                double t = 0;
                for (int k = r; k < 15; k++) t += U[r * 15 + k] * J[k * 30 + c];
                Jw[e] = t;
            }
            __syncwarp();
            for (int e = lane; e < 900; e += 32) {
                const int r = e / 30, c = e - r * 30;
                const int dr = pnp_dof(f, r), dc = pnp_dof(f, c);
                if (dr < 0 || dc < 0) continue;
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + r] * Jw[k * 30 + c];
                atomicAdd(&ws.H[dr * PNP_D + dc], t);
            }
            for (int c = lane; c < 30; c += 32) {
                const int dc = pnp_dof(f, c);
                if (dc < 0) continue;
                double t = 0;
                for (int k = 0; k < 15; k++) t += Jw[k * 30 + c] * rw[k];
                atomicAdd(&ws.g[dc], t);
            }
        }
    }
    __syncthreads();
    const M3 ric = ldm(s.ric), ricT = tr(ric);
    const V3 tic = ld3(s.tic);
    for (int i = 0; i < PNP_N; i++) {                            // PerspectiveFactor per (frame, landmark), CauchyLoss(1)
        if (svf[i]) continue;
        const int n = s.f_n[(size_t)b * PNP_N + i];
        const size_t fo = P_f(s, b, i);
        const V3 Pi = ld3(par + 16 * i);
        const Q4 Qi = ldq(par + 16 * i + 3);
        const M3 RiT = tr(q2R(Qi));
        for (int e = tid; e < n; e += T) {
            const V3 X = ld3(s.f_pos + 3 * (fo + e));
            const double w = s.sqrt_info * (double)s.f_track[fo + e] / 10.0;
            const V3 p_imu = qrot(qinv(Qi), X - Pi);
            const V3 pc = ricT * (p_imu - tic);
            const double dep = pc.z;
            const double rx = w * (pc.x / dep - s.f_obs[2 * (fo + e)]), ry = w * (pc.y / dep - s.f_obs[2 * (fo + e) + 1]);
            const double sq = rx * rx + ry * ry;
            cost += 0.5 * log(1.0 + sq);
            if (lin) {
                const double sr = sqrt(fmax(2.2250738585072014e-308, 1.0 / (1.0 + sq)));      // corrector.cc:113-118 (rho'' < 0: alpha = 0)
                const double r0 = w / dep * sr, r2x = -w * pc.x / (dep * dep) * sr, r2y = -w * pc.y / (dep * dep) * sr;
                // jaco = [ -ric^T Ri^T | ric^T skew(p_imu) ] (3x6); J = reduce * jaco (2x6)
                const M3 A = (-1.0) * (ricT * RiT), Bm = ricT * skew(p_imu);
                double Ji[12];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    Ji[c] = r0 * A.m[c] + r2x * A.m[6 + c];          Ji[6 + c] = r0 * A.m[3 + c] + r2y * A.m[6 + c];
                    Ji[3 + c] = r0 * Bm.m[c] + r2x * Bm.m[6 + c];    Ji[9 + c] = r0 * Bm.m[3 + c] + r2y * Bm.m[6 + c];
                }
                const double c0r = sr * rx, c1r = sr * ry;
                for (int a = 0; a < 6; a++) {
                    for (int c = 0; c < 6; c++) atomicAdd(&ws.H[(9 * i + a) * PNP_D + 9 * i + c], Ji[a] * Ji[c] + Ji[6 + a] * Ji[6 + c]);
                    atomicAdd(&ws.g[9 * i + a], Ji[a] * c0r + Ji[6 + a] * c1r);
                }
            }
        }
    }
    __syncthreads();
    if (lin) {                                                   // constant frames: decoupled unit rows, zero gradient => zero step
        const int *sv = s.solved + (size_t)b * PNP_N;
        for (int e = tid; e < PNP_D * PNP_D; e += T) {
            const int r = e / PNP_D, c = e - r * PNP_D;
            if (sv[r / 9] || sv[c / 9]) ws.H[e] = (r == c) ? 1.0 : 0.0;
        }
        for (int e = tid; e < PNP_D; e += T) if (sv[e / 9]) ws.g[e] = 0.0;
    }
    const double total = block_sum_d(cost, sh_red);
    __syncthreads();
    return total;
}

__device__ inline double pnp_dot(const double *a, const double *b, double *sh_red) {
    double acc = 0;
    for (int i = threadIdx.x; i < PNP_D; i += blockDim.x) acc += a[i] * b[i];
    return block_sum_d(acc, sh_red);
}

// vinsPnP::solve_ceres (vins_pnp.cpp:258-331) + old2new / new2old (:104-195)
__device__ void pnp_solve(const PnpState &s, int b, double *sm, double *sh_red, int *sh_flag) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    PnpWs ws;
    double *p = sm;
    auto take = [&](size_t n) { double *r = p; p += n; return r; };
    ws.H = take((size_t)PNP_D * PNP_D); ws.Sx = take((size_t)PNP_D * PNP_D);
    ws.g = take(64); ws.sc = take(64); ws.dd = take(64); ws.gr = take(64); ws.gn = take(64); ws.st = take(64); ws.up = take(64); ws.rhs = take(64); ws.y = take(64);
    ws.par = take(16 * PNP_N); ws.cand = take(16 * PNP_N); ws.imu = take(930 * (PNP_N - 1));
    const int *sv = s.solved + (size_t)b * PNP_N;
    int *iv = s.iv + (size_t)b * 8;
    double *par = ws.par, *cand = ws.cand;
    for (int i = tid; i < PNP_N; i += T) {                       // old2new
        const size_t k = (size_t)b * PNP_N + i;
        double *q = par + 16 * i;
        st3(q, ld3(s.Ps + 3 * k)); stq(q + 3, R2q(ldm(s.Rs + 9 * k)));
        st3(q + 7, ld3(s.Vs + 3 * k)); st3(q + 10, ld3(s.Bas + 3 * k)); st3(q + 13, ld3(s.Bgs + 3 * k));
    }
    for (int f = warp; f < PNP_N - 1; f += T >> 5) {             // IMUFactorPnP::Evaluate: sqrt_info = LLT(cov^-1).matrixL()^T
        double *pr = P_pre(s, b, f + 1);
        if (pr[PR_SQI_OK] == 0.0) {
            double *scr = ws.imu + (size_t)f * 930;
            if (!imu_sqrt_info_warp(pr + PR_COV, pr + PR_SQI, scr, scr + 225, scr + 450, lane) && lane == 0) iv[2] = VIO_ERR_STATE;
            if (lane == 0) pr[PR_SQI_OK] = 1.0;
        }
    }
    __syncthreads();
    double x_cost = pnp_evaluate(s, b, ws, par, 1, sh_red);
    if (tid == 0) s.dv[(size_t)b * 8 + 6] = x_cost;
    for (int i = tid; i < PNP_D; i += T) ws.sc[i] = 1.0 / (1.0 + sqrt(ws.H[i * PNP_D + i]));        // Jacobi scaling, frozen at iteration 0
    __syncthreads();
    double radius = 1e4, mu = 1e-8, alpha = 0.0, dogleg_norm = 0.0, x_norm = -1.0, mu_gn = 0.0, g2_lin = 0.0, gHg_lin = 0.0;
    bool reuse = false, step_ok = true;
    int iter = 0, invalid_run = 0;
    while (true) {
        if (iter >= s.max_iters) break;
        if (step_ok) {
            double gm = 0;
            for (int i = tid; i < PNP_D; i += T) gm = fmax(gm, fabs(ws.g[i]));
            if (block_max_d(gm, sh_red) <= 1e-10) break;
        }
        if (radius <= 1e-32) break;
        iter++;
        bool linear_ok = true;
        if (!reuse) {
            reuse = true;
            for (int i = tid; i < PNP_D; i += T) {
                const double sc = ws.sc[i];
                const double d = sqrt(fmin(fmax(ws.H[i * PNP_D + i] * sc * sc, 1e-6), 1e32));
                ws.dd[i] = d;
                ws.gr[i] = ws.g[i] * sc / d;
                ws.up[i] = sc * (ws.g[i] * sc / (d * d));
            }
            __syncthreads();
            const double g2 = pnp_dot(ws.gr, ws.gr, sh_red);
            double q = 0;                                        // u^T H u
            for (int e = tid; e < PNP_D * PNP_D; e += T) { const int r = e / PNP_D, c = e - r * PNP_D; q += ws.up[r] * ws.H[e] * ws.up[c]; }
            const double jg2 = block_sum_d(q, sh_red);
            alpha = g2 / jg2;
            g2_lin = g2; gHg_lin = jg2;
            linear_ok = false;
            while (mu < 1.0) {
                for (int e = tid; e < PNP_D * PNP_D; e += T) {
                    const int r = e / PNP_D, c = e - r * PNP_D;
                    double v = ws.H[e] * ws.sc[r] * ws.sc[c];
                    if (r == c) v += mu * ws.dd[r] * ws.dd[r];
                    ws.Sx[e] = v;
                }
                for (int i = tid; i < PNP_D; i += T) ws.rhs[i] = ws.g[i] * ws.sc[i];
                __syncthreads();
                const bool ok = chol_solve(ws.Sx, PNP_D, ws.rhs, ws.y, sh_flag);
                __syncthreads();
                if (ok) {
                    for (int i = tid; i < PNP_D; i += T) ws.gn[i] = -ws.dd[i] * ws.y[i];
                    __syncthreads();
                    mu_gn = mu; linear_ok = true;
                    break;
                }
                mu *= 10.0;
            }
        }
        bool valid = false;
        double model_change = 0.0;
        if (linear_ok) {                                         // ComputeTraditionalDoglegStep (dogleg_strategy.cc:199-255)
            const double g_norm = sqrt(g2_lin);
            const double gn2 = pnp_dot(ws.gn, ws.gn, sh_red);
            const double gngr = pnp_dot(ws.gr, ws.gn, sh_red);
            const double gn_norm = sqrt(gn2);
            double ca, cb;
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
            __syncthreads();
            for (int i = tid; i < PNP_D; i += T) {
                const double st = ca * ws.gr[i] + cb * ws.gn[i];
                ws.st[i] = st;
                ws.up[i] = st / ws.dd[i] * ws.sc[i];
            }
            __syncthreads();
            if (dogleg_norm < 0) dogleg_norm = sqrt(pnp_dot(ws.st, ws.st, sh_red));
            const double dg = ca * g2_lin + cb * gngr;
            const double dMd = ca * ca * gHg_lin + 2.0 * ca * cb * (-g2_lin - mu_gn * gngr) + cb * cb * (-gngr - mu_gn * gn2);
            model_change = -(dg + 0.5 * dMd);
            valid = model_change > 0.0;
        }
        if (!valid) {
            if (++invalid_run >= 5) break;
            mu *= 10.0; reuse = false; step_ok = false;
            continue;
        }
        invalid_run = 0;
        for (int i = tid; i < PNP_N; i += T) {                   // candidate = Plus(x, delta)
            pose_plus(par + 16 * i, ws.up + 9 * i, cand + 16 * i);
            for (int k = 0; k < 3; k++) cand[16 * i + 7 + k] = par[16 * i + 7 + k] + ws.up[9 * i + 6 + k];
            for (int k = 0; k < 6; k++) cand[16 * i + 10 + k] = par[16 * i + 10 + k];
        }
        __syncthreads();
        const double cand_cost = pnp_evaluate(s, b, ws, cand, 0, sh_red);
        double sn = 0;                                           // step / parameter norms over the variable blocks (pose 7 + speed 3)
        for (int i = tid; i < PNP_N * 10; i += T) { const int f = i / 10; if (!sv[f]) { const double d = par[16 * f + (i - 10 * f)] - cand[16 * f + (i - 10 * f)]; sn += d * d; } }
        const double step_norm = sqrt(block_sum_d(sn, sh_red));
        if (step_norm <= 1e-8 * (x_norm + 1e-8)) break;
        if (fabs(x_cost - cand_cost) <= 1e-6 * x_cost) break;
        const double quality = (x_cost - cand_cost) / model_change;
        if (quality > 1e-3) {
            __syncthreads();
            for (int i = tid; i < 16 * PNP_N; i += T) par[i] = cand[i];
            __syncthreads();
            double xn = 0;
            for (int i = tid; i < PNP_N * 10; i += T) { const int f = i / 10; if (!sv[f]) { const double v = par[16 * f + (i - 10 * f)]; xn += v * v; } }
            x_norm = sqrt(block_sum_d(xn, sh_red));
            x_cost = pnp_evaluate(s, b, ws, par, 1, sh_red);
            step_ok = true;
            if (quality < 0.25) radius *= 0.5;
            if (quality > 0.75) radius = fmax(radius, 3.0 * dogleg_norm);
            mu = fmax(1e-8, 2.0 * mu / 10.0);
            reuse = false;
        } else {
            step_ok = false;
            radius *= 0.5;
            reuse = true;
        }
    }
    __syncthreads();
    if (tid == 0) { s.dv[(size_t)b * 8 + 7] = x_cost; iv[3] = iter; }
    for (int i = tid; i < PNP_N; i += T) {                       // new2old (vins_pnp.cpp:141-195): every frame, from the parameter blocks
        const size_t k = (size_t)b * PNP_N + i;
        const double *q = par + 16 * i;
        st3(s.Ps + 3 * k, ld3(q)); stm(s.Rs + 9 * k, q2R(qnormalized(ldq(q + 3))));
        st3(s.Vs + 3 * k, ld3(q + 7)); st3(s.Bas + 3 * k, ld3(q + 10)); st3(s.Bgs + 3 * k, ld3(q + 13));
    }
    __syncthreads();
}

// vinsPnP::processImage (vins_pnp.cpp:236-256) + updateFeatures (:85-102) + slideWindow (:335-381)
__global__ void __launch_bounds__(PNP_T) pnp_image_kernel(PnpState s, const int *__restrict__ counts, const int *__restrict__ ids, const double *__restrict__ obs,
                                                           const double *__restrict__ pos, const int *__restrict__ track, const double *__restrict__ headers, int use_pnp) {
    extern __shared__ __align__(16) double pnp_sm[];
    __shared__ double sh_red[32];
    __shared__ int sh_flag;
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    int *iv = s.iv + (size_t)b * 8;
    const int fc = iv[0];
    const int n = min(counts[b], s.MAXF);
    if (counts[b] > s.MAXF && tid == 0) iv[2] = VIO_ERR_CAPACITY;
    const int *mid = ids + (size_t)b * s.MAXF;
    {   // features[frame_count] = feature_msg; Headers[frame_count] = header
        const size_t fo = P_f(s, b, fc);
        for (int e = tid; e < n; e += T) {
            s.f_id[fo + e] = mid[e]; s.f_track[fo + e] = track[(size_t)b * s.MAXF + e];
            s.f_obs[2 * (fo + e)] = obs[2 * ((size_t)b * s.MAXF + e)]; s.f_obs[2 * (fo + e) + 1] = obs[2 * ((size_t)b * s.MAXF + e) + 1];
            for (int k = 0; k < 3; k++) s.f_pos[3 * (fo + e) + k] = pos[3 * ((size_t)b * s.MAXF + e) + k];
        }
        if (tid == 0) { s.f_n[(size_t)b * PNP_N + fc] = n; s.Headers[(size_t)b * PNP_N + fc] = headers[b]; }
    }
    // updateFeatures: landmarks of the older frames take the newest position / track_num of the same id (ids ascending in both lists)
    for (int i = 0; i < fc; i++) {
        const size_t fo = P_f(s, b, i);
        const int ni = s.f_n[(size_t)b * PNP_N + i];
        for (int e = tid; e < ni; e += T) {
            const int id = s.f_id[fo + e];
            int lo = 0, hi = n - 1;
            while (lo <= hi) {
                const int m = (lo + hi) >> 1, v = mid[m];
                if (v == id) {
                    for (int k = 0; k < 3; k++) s.f_pos[3 * (fo + e) + k] = pos[3 * ((size_t)b * s.MAXF + m) + k];
                    s.f_track[fo + e] = track[(size_t)b * s.MAXF + m];
                    break;
                }
                if (v < id) lo = m + 1; else hi = m - 1;
            }
        }
    }
    __syncthreads();
    if (fc < PNP_N - 1) { if (tid == 0) iv[0] = fc + 1; return; }
    if (use_pnp) pnp_solve(s, b, pnp_sm, sh_red, &sh_flag);
    __syncthreads();
    // slideWindow: frame i <- frame i + 1 (Rs, pre-integration, Headers, Ps, Vs, features, find_solved; NOT the biases), last frame duplicated
    for (int i = 0; i < PNP_N - 1; i++) {
        const size_t k = (size_t)b * PNP_N + i;
        if (tid < 9) s.Rs[9 * k + tid] = s.Rs[9 * (k + 1) + tid];
        if (tid < 3) { s.Ps[3 * k + tid] = s.Ps[3 * (k + 1) + tid]; s.Vs[3 * k + tid] = s.Vs[3 * (k + 1) + tid]; }
        if (tid == 0) { s.Headers[k] = s.Headers[k + 1]; s.solved[k] = s.solved[k + 1]; s.f_n[k] = s.f_n[k + 1]; }
        double *pd = P_pre(s, b, i);
        const double *ps = P_pre(s, b, i + 1);
        for (int e = tid; e < PR_STRIDE; e += T) pd[e] = ps[e];
        const size_t fd = P_f(s, b, i), fs = P_f(s, b, i + 1);
        const int nn = s.f_n[k + 1];
        for (int e = tid; e < nn; e += T) {
            s.f_id[fd + e] = s.f_id[fs + e]; s.f_track[fd + e] = s.f_track[fs + e];
            s.f_obs[2 * (fd + e)] = s.f_obs[2 * (fs + e)]; s.f_obs[2 * (fd + e) + 1] = s.f_obs[2 * (fs + e) + 1];
            for (int q = 0; q < 3; q++) s.f_pos[3 * (fd + e) + q] = s.f_pos[3 * (fs + e) + q];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const size_t k = (size_t)b * PNP_N + PNP_N - 1;
        st3(s.Bas + 3 * k, ld3(s.Bas + 3 * (k - 1))); st3(s.Bgs + 3 * k, ld3(s.Bgs + 3 * (k - 1)));
        s.solved[k] = 0; s.f_n[k] = 0;
        const double *dv = s.dv + (size_t)b * 8;
        pre_init(P_pre(s, b, PNP_N - 1), ld3(dv), ld3(dv + 3), ld3(s.Bas + 3 * k), ld3(s.Bgs + 3 * k));
    }
}

}  // namespace be
