// oracle/pnp_ref.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libvio_b200.so).
//
// CPU reference for the motion-only PnP tracker (FeatureTracker::solveVinsPnP,
// feature_tracker.cpp:107-160 -> vinsPnP, vins_pnp.cpp).  The reference's own vins_pnp.cpp, perspective_factor.cpp and
// imu_factor_pnp.h are compiled UNMODIFIED by oracle/Makefile (shim/ supplies the single OpenCV header name vins_pnp.hpp pulls in);
// this file is the C API for ctypes plus ONE restated method:
//   vinsPnP::processImage  vins_pnp.cpp:236-256   (restated so that it can call the solve below)
//   vinsPnP::solve_ceres   vins_pnp.cpp:258-331   (restated WITHOUT options.max_solver_time_in_seconds = 0.01, which makes the
//                                                  reference's result depend on the speed of the machine; everything else --
//                                                  blocks, constancy flags, factors, DENSE_SCHUR + DOGLEG, 5 iterations -- as written)
// setInit / processIMU / old2new / new2old / updateFeatures / slideWindow are the reference's own (public) methods.
// PARITY PINNING: the reference has no golden vectors for this path; this library is the pin (tests/golden/pnp_golden.npz is
// generated from it by tests/golden/make_golden.py).
#include <cstdio>
#include <vector>

#include "vins_pnp.hpp"
#include "../include/vio_b200.h"

namespace {

// vins_pnp.cpp:258-331 minus the wall-time cap
void pnp_solve(vinsPnP &e) {
    ceres::Problem problem;
    ceres::LossFunction *loss_function = new ceres::CauchyLoss(1.0);
    for (int i = 0; i < PNP_SIZE + 1; i++) {
        problem.AddParameterBlock(e.para_Pose[i], SIZE_POSE, new PoseLocalParameterization());
        problem.AddParameterBlock(e.para_Speed[i], SIZE_SPEED);
        problem.AddParameterBlock(e.para_Bias[i], SIZE_BIAS);
        if (e.find_solved[i]) {
            problem.SetParameterBlockConstant(e.para_Pose[i]);
            problem.SetParameterBlockConstant(e.para_Speed[i]);
        }
        problem.SetParameterBlockConstant(e.para_Bias[i]);
    }
    for (int i = 0; i < NUM_OF_CAM; i++) {
        problem.AddParameterBlock(e.para_Ex_Pose[i], SIZE_POSE, new PoseLocalParameterization());
        problem.SetParameterBlockConstant(e.para_Ex_Pose[i]);
    }
    e.old2new();
    for (int i = 0; i < PNP_SIZE; i++) {
        const int j = i + 1;
        problem.AddResidualBlock(new IMUFactorPnP(e.pre_integrations[j]), NULL, e.para_Pose[i], e.para_Speed[i], e.para_Bias[i], e.para_Pose[j],
                                 e.para_Speed[j], e.para_Bias[j]);
    }
    for (int i = 0; i <= PNP_SIZE; i++)
        for (auto &it : e.features[i])
            problem.AddResidualBlock(new PerspectiveFactor(it.observation, it.position, it.track_num), loss_function, e.para_Pose[i], e.para_Ex_Pose[0]);
    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_SCHUR;
    options.num_threads = 1;
    options.trust_region_strategy_type = ceres::DOGLEG;
    options.use_explicit_schur_complement = true;
    options.minimizer_progress_to_stdout = false;
    options.max_num_iterations = 5;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    e.new2old();
}

}  // namespace

extern "C" {

void *vpnp_create(const vio_config *cfg) {
    TIC_X = cfg->tic[0]; TIC_Y = cfg->tic[1]; TIC_Z = cfg->tic[2];        // globals filled by setGlobalParam() in the app (global_param.cpp:37-39)
    FOCUS_LENGTH_X = cfg->fx; FOCUS_LENGTH_Y = cfg->fy; PX = cfg->cx; PY = cfg->cy;
    vinsPnP *e = new vinsPnP();
    e->setIMUModel();                       // PerspectiveFactor::sqrt_info = FOCUS_LENGTH_X / 1.5 * I  (vins_pnp.cpp:17-20)
    e->setExtrinsic();
    for (int i = 0; i <= PNP_SIZE; i++) e->Headers[i] = -1.0;      // the reference leaves Headers uninitialised (read by setInit)
    return e;
}
void vpnp_destroy(void *h) { delete (vinsPnP *)h; }
int vpnp_size() { return PNP_SIZE; }

// FeatureTracker::solveVinsPnP hands over solved_vins (ViewController.mm:734-739): header, Ba, Bg, P, R (row-major 3x3), V
void vpnp_set_init(void *h, double header, const double *P, const double *R, const double *V, const double *Ba, const double *Bg) {
    VINS_RESULT r;
    r.header = header;
    r.P = Vector3d(P[0], P[1], P[2]); r.V = Vector3d(V[0], V[1], V[2]);
    r.Ba = Vector3d(Ba[0], Ba[1], Ba[2]); r.Bg = Vector3d(Bg[0], Bg[1], Bg[2]);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R(i, j) = R[3 * i + j];
    ((vinsPnP *)h)->setInit(r);
}
void vpnp_process_imu(void *h, double dt, const double *acc, const double *gyr) {
    ((vinsPnP *)h)->processIMU(dt, Vector3d(acc[0], acc[1], acc[2]), Vector3d(gyr[0], gyr[1], gyr[2]));
}
// vins_pnp.cpp:236-256; feature_msg = (id, observation (normalised image coordinates), position (world), track_num), ids ascending
void vpnp_process_image(void *h, int n, const int *ids, const double *obs_xy, const double *pos_xyz, const int *track_num, double header, int use_pnp) {
    vinsPnP &e = *(vinsPnP *)h;
    std::vector<IMG_MSG_LOCAL> msg(n);
    for (int i = 0; i < n; i++) {
        msg[i].id = ids[i];
        msg[i].observation = Vector2d(obs_xy[2 * i], obs_xy[2 * i + 1]);
        msg[i].position = Vector3d(pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2]);
        msg[i].track_num = track_num[i];
    }
    e.features[e.frame_count] = msg;
    e.Headers[e.frame_count] = header;
    e.updateFeatures(msg);
    if (e.frame_count < PNP_SIZE) { e.frame_count++; return; }
    if (use_pnp) pnp_solve(e);
    e.slideWindow();
}
// window state after the call: P[(PNP_SIZE+1)*3], R row-major [(PNP_SIZE+1)*9], V, Ba, Bg [(PNP_SIZE+1)*3], headers, find_solved
void vpnp_get_state(void *h, double *P, double *R, double *V, double *Ba, double *Bg, double *headers, int *find_solved, int *frame_count) {
    vinsPnP &e = *(vinsPnP *)h;
    for (int i = 0; i <= PNP_SIZE; i++) {
        for (int c = 0; c < 3; c++) { P[3 * i + c] = e.Ps[i][c]; V[3 * i + c] = e.Vs[i][c]; Ba[3 * i + c] = e.Bas[i][c]; Bg[3 * i + c] = e.Bgs[i][c]; }
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[9 * i + 3 * r + c] = e.Rs[i](r, c);
        headers[i] = e.Headers[i];
        find_solved[i] = e.find_solved[i] ? 1 : 0;
    }
    *frame_count = e.frame_count;
}
// PerspectiveFactor::Evaluate (perspective_factor.cpp:16-67): residual[2], Jacobians wrt pose (2x6) and extrinsic pose (2x6)
void vpnp_perspective_factor(const double *obs_xy, const double *pos_xyz, int track_num, double fx, const double *pose7, const double *ex7,
                             double *res, double *J_pose, double *J_ex) {
    PerspectiveFactor::sqrt_info = fx / 1.5 * Matrix2d::Identity();
    PerspectiveFactor f(Vector2d(obs_xy[0], obs_xy[1]), Vector3d(pos_xyz[0], pos_xyz[1], pos_xyz[2]), track_num);
    double j0[14], j1[14];
    double *jac[2] = {j0, j1};
    const double *params[2] = {pose7, ex7};
    f.Evaluate(params, res, jac);
    for (int r = 0; r < 2; r++) for (int c = 0; c < 6; c++) { J_pose[6 * r + c] = j0[7 * r + c]; J_ex[6 * r + c] = j1[7 * r + c]; }
}

}  // extern "C"
