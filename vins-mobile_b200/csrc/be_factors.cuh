// be_factors.cuh -- residual / Jacobian evaluation of the three factor types and IMU pre-integration (f64).
// Reference arithmetic (paths under /root/reference/VINS_ios/):
//   IntegrationBase::midPointIntegration / propagate   integration_base.h:63-169
//   IntegrationBase::evaluate                          integration_base.h:171-198
//   IMUFactor::Evaluate                                imu_factor.h:27-184
//   ProjectionFactor::Evaluate                         projection_facor.cpp:16-99
//   ceres::CauchyLoss(1.0) + Corrector                 ceres-solver/internal/ceres/loss_function.cc:72-79, corrector.cc:41-155
#pragma once
#include "be_math.cuh"

namespace be {

// ---- pre-integration record layout (doubles) ------------------------------------------------------------
constexpr int PR_DP = 0, PR_DQ = 3, PR_DV = 7, PR_LBA = 10, PR_LBG = 13, PR_SUMDT = 16, PR_ACC0 = 17, PR_GYR0 = 20, PR_VALID = 23,
              PR_SQI_OK = 24;     // 1 when PR_SQI matches PR_COV (cleared by every propagation step, set by prepare_kernel)
constexpr int PR_LIN_ACC = 25, PR_LIN_GYR = 28;   // linearized_acc / linearized_gyr: the acc_0, gyr_0 the interval started from (repropagate needs them)
constexpr int PR_JAC = 32, PR_COV = PR_JAC + 225, PR_SQI = PR_COV + 225, PR_STRIDE = PR_SQI + 225 + 7;   // 714
constexpr int PR_ABG = PR_SQI + 225;              // 3: gyroscope bias the frame's all_image_frame copy is linearised at (initialisation only)

// IntegrationBase ctor (integration_base.h:26-44): called by one lane
__device__ inline void pre_init(double *pr, V3 acc0, V3 gyr0, V3 ba, V3 bg) {
    st3(pr + PR_DP, v3(0, 0, 0)); stq(pr + PR_DQ, q4(0, 0, 0, 1)); st3(pr + PR_DV, v3(0, 0, 0));
    st3(pr + PR_LBA, ba); st3(pr + PR_LBG, bg); pr[PR_SUMDT] = 0;
    st3(pr + PR_ACC0, acc0); st3(pr + PR_GYR0, gyr0); pr[PR_VALID] = 1; pr[PR_SQI_OK] = 0;
    st3(pr + PR_LIN_ACC, acc0); st3(pr + PR_LIN_GYR, gyr0); st3(pr + PR_ABG, v3(0, 0, 0));
    for (int i = 0; i < 225; i++) { pr[PR_JAC + i] = (i % 16 == 0) ? 1.0 : 0.0; pr[PR_COV + i] = 0.0; }
}

struct PreScratch { double F[225], V[270], T1[225], T2[225]; };   // per warp

// IntegrationBase::push_back -> propagate -> midPointIntegration for ONE sample; whole warp cooperates.
__device__ inline void pre_propagate_warp(double *pr, double dt, V3 a1, V3 g1, const double noise[6], PreScratch &S) {
    const int lane = threadIdx.x & 31;
    const V3 a0 = ld3(pr + PR_ACC0), g0 = ld3(pr + PR_GYR0), ba = ld3(pr + PR_LBA), bg = ld3(pr + PR_LBG);
    const V3 dp = ld3(pr + PR_DP), dv = ld3(pr + PR_DV);
    const Q4 dq = ldq(pr + PR_DQ);
    const V3 un_acc_0 = qrot(dq, a0 - ba);
    const V3 un_gyr = 0.5 * (g0 + g1) - bg;
    const Q4 rq = qmul(dq, q4(un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2, 1.0));
    const V3 un_acc_1 = qrot(rq, a1 - ba);
    const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    const V3 rp = dp + dv * dt + 0.5 * un_acc * dt * dt;
    const V3 rv = dv + un_acc * dt;
    // F (15x15) and V (15x18), built by lane 0..: each lane fills a few entries; simplest: zero all, then lane 0 writes blocks
    for (int i = lane; i < 225; i += 32) S.F[i] = 0.0;
    for (int i = lane; i < 270; i += 32) S.V[i] = 0.0;
    __syncwarp();
    if (lane == 0) {
        const M3 R0 = q2R(dq), R1 = q2R(rq);
        const M3 Rw = skew(un_gyr), Ra0 = skew(a0 - ba), Ra1 = skew(a1 - ba);
        const M3 I = eye3();
        const M3 ImW = I - dt * Rw;
        const M3 R0A0 = R0 * Ra0, R1A1 = R1 * Ra1;
        const M3 R1A1I = R1A1 * ImW;
        const M3 R01 = R0 + R1;
        auto put = [](double *M, int ld, int r, int c, const M3 &B) {
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[(r + i) * ld + c + j] = B.m[3 * i + j];
        };
        put(S.F, 15, 0, 0, I);
        put(S.F, 15, 0, 3, (-0.25 * dt * dt) * R0A0 + (-0.25 * dt * dt) * R1A1I);
        put(S.F, 15, 0, 6, dt * I);
        put(S.F, 15, 0, 9, (-0.25 * dt * dt) * R01);
        put(S.F, 15, 0, 12, (-0.25 * dt * dt * -dt) * R1A1);
        put(S.F, 15, 3, 3, ImW);
        put(S.F, 15, 3, 12, (-1.0 * dt) * I);
        put(S.F, 15, 6, 3, (-0.5 * dt) * R0A0 + (-0.5 * dt) * R1A1I);
        put(S.F, 15, 6, 6, I);
        put(S.F, 15, 6, 9, (-0.5 * dt) * R01);
        put(S.F, 15, 6, 12, (-0.5 * dt * -dt) * R1A1);
        put(S.F, 15, 9, 9, I);
        put(S.F, 15, 12, 12, I);
        const M3 V03 = (0.25 * -1.0 * dt * dt * 0.5 * dt) * R1A1;
        const M3 V63 = (0.5 * -1.0 * dt * 0.5 * dt) * R1A1;
        put(S.V, 18, 0, 0, (0.25 * dt * dt) * R0);
        put(S.V, 18, 0, 3, V03);
        put(S.V, 18, 0, 6, (0.25 * dt * dt) * R1);
        put(S.V, 18, 0, 9, V03);
        put(S.V, 18, 3, 3, (0.5 * dt) * I);
        put(S.V, 18, 3, 9, (0.5 * dt) * I);
        put(S.V, 18, 6, 0, (0.5 * dt) * R0);
        put(S.V, 18, 6, 3, V63);
        put(S.V, 18, 6, 6, (0.5 * dt) * R1);
        put(S.V, 18, 6, 9, V63);
        put(S.V, 18, 9, 12, dt * I);
        put(S.V, 18, 12, 15, dt * I);
    }
    __syncwarp();
    double *J = pr + PR_JAC, *C = pr + PR_COV;
    // T1 = F*J ; T2 = F*C
    for (int i = lane; i < 225; i += 32) {
        const int r = i / 15, c = i % 15;
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 15; k++) { s1 += S.F[r * 15 + k] * J[k * 15 + c]; s2 += S.F[r * 15 + k] * C[k * 15 + c]; }
        S.T1[i] = s1; S.T2[i] = s2;
    }
    __syncwarp();
    // J = T1 ; C = T2*F^T + V*Q*V^T
    for (int i = lane; i < 225; i += 32) {
        const int r = i / 15, c = i % 15;
        double s = 0;
        for (int k = 0; k < 15; k++) s += S.T2[r * 15 + k] * S.F[c * 15 + k];
        for (int k = 0; k < 18; k++) s += S.V[r * 18 + k] * noise[k / 3] * S.V[c * 18 + k];
        J[i] = S.T1[i]; C[i] = s;
    }
    __syncwarp();
    if (lane == 0) {
        st3(pr + PR_DP, rp); st3(pr + PR_DV, rv);
        stq(pr + PR_DQ, qnormalized(rq));
        pr[PR_SUMDT] += dt; pr[PR_SQI_OK] = 0;
        st3(pr + PR_ACC0, a1); st3(pr + PR_GYR0, g1);
    }
    __syncwarp();
}

// sqrt_info = LLT(covariance^-1).matrixL()^T (imu_factor.h:72), computed once per solve (constant: repropagation is #if 0'd,
// imu_factor.h:60-66).  cov = Lc Lc^T, cov^-1 = Lc^-T Lc^-1, then Cholesky of cov^-1; U = L2^T stored row-major (upper).
// Serial, one thread (15x15).
__device__ inline bool imu_sqrt_info(const double *cov, double *U) {
    double L[225], Li[225], A[225];
    for (int i = 0; i < 225; i++) { L[i] = 0; Li[i] = 0; }
    for (int j = 0; j < 15; j++) {
        double d = cov[j * 15 + j];
        for (int k = 0; k < j; k++) d -= L[j * 15 + k] * L[j * 15 + k];
        if (!(d > 0)) return false;
        d = sqrt(d);
        L[j * 15 + j] = d;
        for (int i = j + 1; i < 15; i++) {
            double s = cov[i * 15 + j];
            for (int k = 0; k < j; k++) s -= L[i * 15 + k] * L[j * 15 + k];
            L[i * 15 + j] = s / d;
        }
    }
    for (int c = 0; c < 15; c++)                       // Li = Lc^-1 (lower)
        for (int i = c; i < 15; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= L[i * 15 + k] * Li[k * 15 + c];
            Li[i * 15 + c] = s / L[i * 15 + i];
        }
    for (int i = 0; i < 15; i++)                       // A = Li^T Li
        for (int j = 0; j <= i; j++) {
            double s = 0;
            for (int k = i; k < 15; k++) s += Li[k * 15 + i] * Li[k * 15 + j];
            A[i * 15 + j] = s; A[j * 15 + i] = s;
        }
    for (int i = 0; i < 225; i++) L[i] = 0;
    for (int j = 0; j < 15; j++) {
        double d = A[j * 15 + j];
        for (int k = 0; k < j; k++) d -= L[j * 15 + k] * L[j * 15 + k];
        if (!(d > 0)) return false;
        d = sqrt(d);
        L[j * 15 + j] = d;
        for (int i = j + 1; i < 15; i++) {
            double s = A[i * 15 + j];
            for (int k = 0; k < j; k++) s -= L[i * 15 + k] * L[j * 15 + k];
            L[i * 15 + j] = s / d;
        }
    }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) U[i * 15 + j] = L[j * 15 + i];
    return true;
}

// Warp-cooperative variant (shared-memory work arrays of 225 doubles each): same result up to the summation order.
//   chol15_warp: in-place right-looking Cholesky of the lower triangle of a 15x15 matrix
__device__ inline bool chol15_warp(double *A, int lane) {
    bool ok = true;
    for (int j = 0; j < 15; j++) {
        const double piv = A[j * 15 + j];
        ok &= piv > 0;
        const double d = sqrt(piv), id = 1.0 / d;
        __syncwarp();
        if (lane == 0) A[j * 15 + j] = d;
        if (lane > j && lane < 15) A[lane * 15 + j] *= id;
        __syncwarp();
        const int m = 14 - j;                                        // trailing rows j+1..14: pairs (i >= k) of the lower triangle
        for (int e = lane; e < m * (m + 1) / 2; e += 32) {
            int r = 0, t = e;
            while (t > r) { t -= r + 1; r++; }
            const int i = j + 1 + r, k = j + 1 + t;
            A[i * 15 + k] -= A[i * 15 + j] * A[k * 15 + j];
        }
        __syncwarp();
    }
    return ok;
}
__device__ inline bool imu_sqrt_info_warp(const double *cov, double *U, double *L, double *Li, double *A, int lane) {
    for (int e = lane; e < 225; e += 32) { L[e] = cov[e]; Li[e] = 0.0; }
    __syncwarp();
    bool ok = chol15_warp(L, lane);
    if (lane < 15) {                                                 // Li = L^-1, one column per lane
        const int c = lane;
        for (int i = c; i < 15; i++) {
            double sum = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) sum -= L[i * 15 + k] * Li[k * 15 + c];
            Li[i * 15 + c] = sum / L[i * 15 + i];
        }
    }
    __syncwarp();
    for (int e = lane; e < 120; e += 32) {                           // A = Li^T Li (lower)
        int i = 0, t = e;
        while (t > i) { t -= i + 1; i++; }
        const int j = t;
        double sum = 0;
        for (int k = i; k < 15; k++) sum += Li[k * 15 + i] * Li[k * 15 + j];
        A[i * 15 + j] = sum;
    }
    __syncwarp();
    ok &= chol15_warp(A, lane);
    for (int e = lane; e < 225; e += 32) { const int i = e / 15, j = e - 15 * i; U[e] = (j >= i) ? A[j * 15 + i] : 0.0; }
    __syncwarp();
    return ok;
}

// IntegrationBase::evaluate + IMUFactor::Evaluate.  Serial (one thread); res[15] is weighted by sqrt_info; when J != nullptr it
// receives the 15x30 row-major Jacobian in LOCAL coordinates [pose_i(6) sb_i(9) pose_j(6) sb_j(9)] (the 7th pose column is zero,
// PoseLocalParameterization::ComputeJacobian is [I;0]) BEFORE weighting; the caller multiplies by sqrt_info.
__device__ inline void imu_residual(const double *pr, double g, const double *pi, const double *sbi, const double *pj, const double *sbj,
                                    double *res_raw, double *J) {
    const V3 Pi = ld3(pi), Pj = ld3(pj), Vi = ld3(sbi), Vj = ld3(sbj), Bai = ld3(sbi + 3), Bgi = ld3(sbi + 6), Baj = ld3(sbj + 3), Bgj = ld3(sbj + 6);
    const Q4 Qi = ldq(pi + 3), Qj = ldq(pj + 3);
    const double *Jm = pr + PR_JAC;
    auto blk = [&](int r, int c) { M3 B; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) B.m[3 * i + j] = Jm[(r + i) * 15 + c + j]; return B; };
    const M3 dp_dba = blk(0, 9), dp_dbg = blk(0, 12), dq_dbg = blk(3, 12), dv_dba = blk(6, 9), dv_dbg = blk(6, 12);
    const V3 dba = Bai - ld3(pr + PR_LBA), dbg = Bgi - ld3(pr + PR_LBG);
    const Q4 dq = ldq(pr + PR_DQ);
    const Q4 cq = qmul(dq, deltaQ(dq_dbg * dbg));
    const V3 cv = ld3(pr + PR_DV) + dv_dba * dba + dv_dbg * dbg;
    const V3 cp = ld3(pr + PR_DP) + dp_dba * dba + dp_dbg * dbg;
    const double sdt = pr[PR_SUMDT];
    const V3 G = v3(0, 0, g);
    const Q4 Qi_inv = qinv(Qi);
    const V3 tp = qrot(Qi_inv, 0.5 * G * sdt * sdt + Pj - Pi - Vi * sdt);
    const V3 tv = qrot(Qi_inv, G * sdt + Vj - Vi);
    const Q4 qe = qmul(qinv(cq), qmul(Qi_inv, Qj));
    st3(res_raw + 0, tp - cp);
    st3(res_raw + 3, v3(2 * qe.x, 2 * qe.y, 2 * qe.z));
    st3(res_raw + 6, tv - cv);
    st3(res_raw + 9, Baj - Bai);
    st3(res_raw + 12, Bgj - Bgi);
    if (!J) return;
    for (int i = 0; i < 450; i++) J[i] = 0.0;
    auto put = [&](int r, int c, const M3 &B) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r + i) * 30 + c + j] = B.m[3 * i + j]; };
    const M3 RiT = q2R(Qi_inv);                           // Qi.inverse().toRotationMatrix()
    const M3 negI = -1.0 * eye3();
    // pose_i (cols 0..5)
    put(0, 0, -1.0 * RiT);
    put(0, 3, skew(tp));
    put(3, 3, -1.0 * qleft_qright_33(qmul(qinv(Qj), Qi), cq));
    put(6, 3, skew(tv));
    // speedbias_i (cols 6..14)
    put(0, 6, (-sdt) * RiT);
    put(0, 9, -1.0 * dp_dba);
    put(0, 12, -1.0 * dp_dbg);
    put(3, 12, -1.0 * (qleft33(qmul(qmul(qinv(Qj), Qi), cq)) * dq_dbg));
    put(6, 6, -1.0 * RiT);
    put(6, 9, -1.0 * dv_dba);
    put(6, 12, -1.0 * dv_dbg);
    put(9, 9, negI);
    put(12, 12, negI);
    // pose_j (cols 15..20)
    put(0, 15, RiT);
    put(3, 18, qleft33(qmul(qinv(cq), qmul(Qi_inv, Qj))));
    // speedbias_j (cols 21..29)
    put(6, 21, RiT);
    put(9, 24, eye3());
    put(12, 27, eye3());
}


// Warp-cooperative variant of imu_residual: every lane evaluates the (cheap) shared prelude, lane 0 stores the raw residual and
// lanes 0..17 each form ONE of the 18 non-zero 3x3 blocks of the 15x30 Jacobian (the serial version spent most of its time in
// dependent f64 chains on a single lane while 31 lanes idled).
__device__ inline void imu_residual_warp(const double *pr, double g, const double *pi, const double *sbi, const double *pj, const double *sbj,
                                         double *res_raw, double *J, int lane) {
    const V3 Pi = ld3(pi), Pj = ld3(pj), Vi = ld3(sbi), Vj = ld3(sbj), Bai = ld3(sbi + 3), Bgi = ld3(sbi + 6), Baj = ld3(sbj + 3), Bgj = ld3(sbj + 6);
    const Q4 Qi = ldq(pi + 3), Qj = ldq(pj + 3);
    const double *Jm = pr + PR_JAC;
    auto blk = [&](int r, int c) { M3 B; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) B.m[3 * i + j] = Jm[(r + i) * 15 + c + j]; return B; };
    const M3 dq_dbg = blk(3, 12);
    const V3 dba = Bai - ld3(pr + PR_LBA), dbg = Bgi - ld3(pr + PR_LBG);
    const Q4 dq = ldq(pr + PR_DQ);
    const Q4 cq = qmul(dq, deltaQ(dq_dbg * dbg));
    const double sdt = pr[PR_SUMDT];
    const V3 G = v3(0, 0, g);
    const Q4 Qi_inv = qinv(Qi);
    const V3 tp = qrot(Qi_inv, 0.5 * G * sdt * sdt + Pj - Pi - Vi * sdt);
    const V3 tv = qrot(Qi_inv, G * sdt + Vj - Vi);
    if (lane == 0) {
        const M3 dp_dba = blk(0, 9), dp_dbg = blk(0, 12), dv_dba = blk(6, 9), dv_dbg = blk(6, 12);
        const V3 cv = ld3(pr + PR_DV) + dv_dba * dba + dv_dbg * dbg;
        const V3 cp = ld3(pr + PR_DP) + dp_dba * dba + dp_dbg * dbg;
        const Q4 qe = qmul(qinv(cq), qmul(Qi_inv, Qj));
        st3(res_raw + 0, tp - cp);
        st3(res_raw + 3, v3(2 * qe.x, 2 * qe.y, 2 * qe.z));
        st3(res_raw + 6, tv - cv);
        st3(res_raw + 9, Baj - Bai);
        st3(res_raw + 12, Bgj - Bgi);
    }
    if (!J) return;
    for (int i = lane; i < 450; i += 32) J[i] = 0.0;
    __syncwarp();
    if (lane >= 18) return;
    M3 B; int r = 0, c = 0;
    const M3 RiT = q2R(Qi_inv);
    switch (lane) {
        case 0: r = 0; c = 0; B = -1.0 * RiT; break;
        case 1: r = 0; c = 3; B = skew(tp); break;
        case 2: r = 3; c = 3; B = -1.0 * qleft_qright_33(qmul(qinv(Qj), Qi), cq); break;
        case 3: r = 6; c = 3; B = skew(tv); break;
        case 4: r = 0; c = 6; B = (-sdt) * RiT; break;
        case 5: r = 0; c = 9; B = -1.0 * blk(0, 9); break;
        case 6: r = 0; c = 12; B = -1.0 * blk(0, 12); break;
        case 7: r = 3; c = 12; B = -1.0 * (qleft33(qmul(qmul(qinv(Qj), Qi), cq)) * dq_dbg); break;
        case 8: r = 6; c = 6; B = -1.0 * RiT; break;
        case 9: r = 6; c = 9; B = -1.0 * blk(6, 9); break;
        case 10: r = 6; c = 12; B = -1.0 * blk(6, 12); break;
        case 11: r = 9; c = 9; B = -1.0 * eye3(); break;
        case 12: r = 12; c = 12; B = -1.0 * eye3(); break;
        case 13: r = 0; c = 15; B = RiT; break;
        case 14: r = 3; c = 18; B = qleft33(qmul(qinv(cq), qmul(Qi_inv, Qj))); break;
        case 15: r = 6; c = 21; B = RiT; break;
        case 16: r = 9; c = 24; B = eye3(); break;
        default: r = 12; c = 27; B = eye3(); break;
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r + i) * 30 + c + j] = B.m[3 * i + j];
}

// ProjectionFactor::Evaluate + CauchyLoss(1.0) corrector.  Returns the robustified cost 0.5*rho(s).
//   r2[2]   : corrected residual  sqrt(rho') * r
//   Ji,Jj   : 2x6 corrected Jacobians wrt pose_i / pose_j (local), Jl : 2x1 wrt inverse depth      (nullptr = cost only)
struct ProjConst { M3 ric; V3 tic; double sqrt_info; };
__device__ inline double proj_eval(const ProjConst &K, V3 pts_i, V3 pts_j, const double *pi, const double *pj, double inv_dep, double *r2,
                                   double *Ji, double *Jj, double *Jl, double *raw_sq_norm = nullptr) {
    const V3 Pi = ld3(pi), Pj = ld3(pj);
    const M3 Ri = q2R(ldq(pi + 3)), Rj = q2R(ldq(pj + 3));
    const V3 pc_i = pts_i * (1.0 / inv_dep);
    const V3 p_imu_i = K.ric * pc_i + K.tic;
    const V3 pw = Ri * p_imu_i + Pi;
    const M3 RjT = tr(Rj), ricT = tr(K.ric);
    const V3 p_imu_j = RjT * (pw - Pj);
    const V3 pc_j = ricT * (p_imu_j - K.tic);
    const double dep = pc_j.z;
    const double rx = K.sqrt_info * (pc_j.x / dep - pts_j.x), ry = K.sqrt_info * (pc_j.y / dep - pts_j.y);
    const double s = rx * rx + ry * ry;
    if (raw_sq_norm) *raw_sq_norm = s;
    const double sum = 1.0 + s, inv = 1.0 / sum;
    const double rho0 = log(sum), rho1 = fmax(2.2250738585072014e-308, inv);
    const double sr = sqrt(rho1);                      // rho[2] < 0 -> residual_scaling = sqrt(rho1), alpha = 0 (corrector.cc:113-118)
    if (r2) { r2[0] = sr * rx; r2[1] = sr * ry; }
    if (Ji) {
        double red[6];                                    // 2x3
        red[0] = K.sqrt_info / dep; red[1] = 0; red[2] = -K.sqrt_info * pc_j.x / (dep * dep);
        red[3] = 0; red[4] = K.sqrt_info / dep; red[5] = -K.sqrt_info * pc_j.y / (dep * dep);
        const M3 A = ricT * RjT;                          // ric^T Rj^T
        const M3 AR = A * Ri;
        const M3 Bi = -1.0 * (AR * skew(p_imu_i));
        const M3 Bj = ricT * skew(p_imu_j);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) {
                double a = 0, b = 0, cc = 0;
                for (int k = 0; k < 3; k++) { a += red[3 * r + k] * A.m[3 * k + c]; b += red[3 * r + k] * Bi.m[3 * k + c]; cc += red[3 * r + k] * Bj.m[3 * k + c]; }
                Ji[6 * r + c] = sr * a; Ji[6 * r + 3 + c] = sr * b;
                Jj[6 * r + c] = -sr * a; Jj[6 * r + 3 + c] = sr * cc;
            }
        const V3 t = (AR * K.ric) * pts_i;
        const double k = -1.0 / (inv_dep * inv_dep);
        Jl[0] = sr * k * (red[0] * t.x + red[1] * t.y + red[2] * t.z);
        Jl[1] = sr * k * (red[3] * t.x + red[4] * t.y + red[5] * t.z);
    }
    return 0.5 * rho0;
}

// PoseLocalParameterization::Plus (pose_local_parameterization.cpp:11-26): p + dp, (q * deltaQ(dtheta)).normalized()
__device__ inline void pose_plus(const double *x, const double *d, double *out) {
    out[0] = x[0] + d[0]; out[1] = x[1] + d[1]; out[2] = x[2] + d[2];
    stq(out + 3, qnormalized(qmul(ldq(x + 3), deltaQ(v3(d[3], d[4], d[5])))));
}

}  // namespace be
