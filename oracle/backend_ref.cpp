// oracle/backend_ref.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libvio_b200.so).
//
// CPU reference for the sliding-window back end.  The ARITHMETIC is the reference's own, compiled
// unmodified from /root/reference by oracle/Makefile:
//   IMUFactor / IntegrationBase        VINS_ios/imu_factor.h, integration_base.h
//   ProjectionFactor                   VINS_ios/projection_facor.cpp
//   MarginalizationInfo / Factor       VINS_ios/marginalization_factor.cpp
//   PoseLocalParameterization          VINS_ios/pose_local_parameterization.cpp
//   ceres::Solve (DENSE_SCHUR, DOGLEG) VINS_ThirdPartyLib/ceres-solver 1.12.0 (+ Eigen 3.3.0)
// What is RESTATED here (VINS.cpp / feature_manager.cpp cannot be compiled: they pull in
// <opencv2/opencv.hpp>, draw_result, loop closure -- SURVEY.md "Facts") is the estimator loop that
// drives them, with WINDOW_SIZE / NUM_OF_F turned into run-time values:
//   processIMU      VINS.cpp:333-375      processImage   VINS.cpp:377-478
//   solve_ceres     VINS.cpp:480-831      old2new/new2old VINS.cpp:89-212
//   slideWindow*    VINS.cpp:1149-1273    failureDetection VINS.cpp:214-265
//   FeatureManager  feature_manager.cpp:103-155,190-406
// Deviations, all documented in DESIGN.md: (1) the wall-time cap on ceres::Solve (VINS.cpp:648-653)
// is removed (it makes the reference non-deterministic); (2) the initialisation pipeline
// (solveInitial, VINS.cpp:833-1102) is replaced by a caller-supplied window (vref_set_init_window);
// (3) loop-closure factors (VINS.cpp:571-637) are absent (retrive_pose_data empty => inert).
// PARITY PINNING: the reference has no golden vectors for this path; this library IS the pin.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <unordered_map>
#include <algorithm>
#include <chrono>
#include <sys/mman.h>

#include "global_param.hpp"
#include "utility.hpp"
#include "imu_factor.h"
#include "projection_facor.hpp"
#include "pose_local_parameterization.hpp"
#include "marginalization_factor.hpp"
#include <ceres/ceres.h>

#include "../include/vio_b200.h"

using namespace Eigen;

namespace {

struct Track {                      // FeaturePerId, feature_manager.hpp:42-68
    int id, start;
    std::vector<Vector3d> obs;      // FeaturePerFrame::point (z-normalised)
    double depth = -1.0;            // estimated_depth
    int solve_flag = 0;
    int end() const { return start + (int)obs.size() - 1; }
};

struct Est {
    vio_config c;
    int W;
    std::vector<Track> feat;        // f_manager.feature (always sorted by id: ids are monotone)
    int last_track_num = 0;
    int frame_count = 0;
    bool first_imu = false;
    int solver_flag = 0;            // 0 INITIAL, 1 NON_LINEAR
    int marg_flag = 0;              // 0 MARGIN_OLD, 1 MARGIN_SECOND_NEW
    std::vector<Vector3d> Ps, Vs, Bas, Bgs;
    std::vector<Matrix3d> Rs;
    std::vector<double> Headers;
    std::vector<IntegrationBase *> pre;
    std::vector<std::vector<double>> dt_buf;
    std::vector<std::vector<Vector3d>> acc_buf, gyr_buf;
    Vector3d acc_0, gyr_0, tic;
    Matrix3d ric;
    // para_Pose / para_SpeedBias / para_Feature / para_Ex_Pose (VINS.hpp:79-82): consecutive arrays, as in the reference object,
    // placed in an arena at a FIXED virtual address (ParaArena below)
    struct Span { double *p = nullptr; size_t n = 0; double *data() { return p; } double &operator[](size_t i) { return p[i]; }
                  const double &operator[](size_t i) const { return p[i]; } };
    Span para_Pose, para_SB, para_Feature, para_Ex;
    int arena_slot = -1;
    // loop closure: retrive_pose_data / front_pose (VINS.hpp:28-45, VINS.cpp:571-637)
    double loop_hdr = -1; std::vector<int> loop_ids; std::vector<double> loop_xy; Vector3d loop_P_old = Vector3d::Zero(); Quaterniond loop_Q_old{1, 0, 0, 0};
    double loop_pose[7] = {0, 0, 0, 0, 0, 0, 1};
    int loop_nfac = 0, loop_frame = -1; bool loop_enable = false;
    double loop_out[12] = {0};       // relative_t3, relative_q4 (xyzw), relative_yaw, drift_yaw, t_drift3
    // wall-clock per stage (bench.py cpu_baseline split): [0] processImage total, [1] ceres::Solve, [2] marginalisation, [3] processIMU
    double t_stage[4] = {0, 0, 0, 0};
    MarginalizationInfo *last_marg = nullptr;
    std::vector<double *> last_marg_blocks;
    int failure_occur = 0;
    Matrix3d last_R, last_R_old, back_R0;
    Vector3d last_P, last_P_old, back_P0;
    // initialisation from SfM poses (vref_set_init_sfm): visualInitialAlign restated around the reference's own VisualIMUAlignment
    bool sfm_pending = false;
    std::vector<Matrix3d> sR; std::vector<Vector3d> sT;      // ImageFrame::R / T of every frame of the map, map order
    // all_image_frame (VINS.hpp:141; filled while INITIAL, VINS.cpp:392-398): per frame the header and what its tmp_pre_integration was fed --
    // start values, samples, and the gyroscope bias it is currently linearised at (solveGyroscopeBias repropagates it)
    struct AllFrame { double hdr = 0; Vector3d acc0 = Vector3d::Zero(), gyr0 = Vector3d::Zero(), abg = Vector3d::Zero();
                      std::vector<double> dt; std::vector<Vector3d> acc, gyr; };
    std::vector<AllFrame> all;                               // closed records
    AllFrame tmp;                                            // the open one (tmp_pre_integration)
    Vector3d g_init = Vector3d::Zero(); int align_ok = -1; double scale_init = 0;
    // external initialisation
    bool init_pending = false;
    std::vector<Vector3d> iP, iV; std::vector<Quaterniond> iQ; Vector3d iBa, iBg;
    // diagnostics
    int n_feat_solve = 0, n_proj = 0, iters = 0;
    double cost0 = 0, cost1 = 0;
    std::vector<double> post_solve;     // (W+1)*16 state right after new2old()

    double *pose(int i) { return &para_Pose[7 * i]; }
    double *sb(int i) { return &para_SB[9 * i]; }
    double *feat_p(int i) { return &para_Feature[i]; }
};

// MarginalizationInfo keys its maps by reinterpret_cast<long>(parameter address) and fixes the block order of the prior by ITERATING an
// unordered_map over those keys (marginalization_factor.cpp:185-200): with heap-allocated parameter arrays the
// order -- and with it the round-off of every prior -- changes from one estimator object (and one process) to the next.  The oracle
// pins it: the parameter arrays of estimator slot k live at the fixed address ARENA_BASE + k * ARENA_STRIDE (lowest free slot
// first), so a fresh estimator always sees the same addresses and the reference arithmetic becomes reproducible.
namespace ParaArena {
const uintptr_t ARENA_BASE = 0x5f0000000000ull;
const size_t ARENA_STRIDE = 1u << 20;
const int SLOTS = 64;
bool used[SLOTS];
double *claim(int *slot, size_t doubles) {
    if (doubles * sizeof(double) > ARENA_STRIDE) return nullptr;
    for (int k = 0; k < SLOTS; k++) {
        if (used[k]) continue;
        void *want = (void *)(ARENA_BASE + (uintptr_t)k * ARENA_STRIDE);
        void *got = mmap(want, ARENA_STRIDE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
        if (got != want) { if (got != MAP_FAILED) munmap(got, ARENA_STRIDE); continue; }
        used[k] = true; *slot = k;
        return (double *)got;
    }
    return nullptr;
}
void release(int slot) {
    if (slot < 0) return;
    munmap((void *)(ARENA_BASE + (uintptr_t)slot * ARENA_STRIDE), ARENA_STRIDE);
    used[slot] = false;
}
}  // namespace ParaArena

bool in_solve(const Est &e, const Track &t) { return (int)t.obs.size() >= 2 && t.start < e.W - 2; }

void clear_state(Est &e) {           // VINS::clearState, VINS.cpp:35-80
    int n = e.W + 1;
    e.Ps.assign(n, Vector3d::Zero()); e.Vs.assign(n, Vector3d::Zero());
    e.Bas.assign(n, Vector3d::Zero()); e.Bgs.assign(n, Vector3d::Zero());
    e.Rs.assign(n, Matrix3d::Identity()); e.Headers.assign(n, 0.0);
    for (auto *p : e.pre) delete p;
    e.pre.assign(n, nullptr);
    e.dt_buf.assign(n, {}); e.acc_buf.assign(n, {}); e.gyr_buf.assign(n, {});
    e.tic = Vector3d(e.c.tic[0], e.c.tic[1], e.c.tic[2]);
    e.ric = Map<const Matrix<double, 3, 3, RowMajor>>(e.c.ric);
    e.frame_count = 0; e.first_imu = false; e.solver_flag = 0;
    e.all.clear(); e.tmp = Est::AllFrame(); e.sfm_pending = false;                  // all_image_frame.clear(), VINS.cpp:62-68
    e.align_ok = -1;
    delete e.last_marg; e.last_marg = nullptr; e.last_marg_blocks.clear();
    e.feat.clear();
}

void process_imu(Est &e, double dt, const Vector3d &acc, const Vector3d &gyr) {   // VINS.cpp:333-375
    if (!e.first_imu) { e.first_imu = true; e.acc_0 = acc; e.gyr_0 = gyr; }
    int j = e.frame_count;
    if (!e.pre[j]) e.pre[j] = new IntegrationBase{e.acc_0, e.gyr_0, e.Bas[j], e.Bgs[j]};
    if (j != 0) {
        e.pre[j]->push_back(dt, acc, gyr);
        if (e.solver_flag != 1) { e.tmp.dt.push_back(dt); e.tmp.acc.push_back(acc); e.tmp.gyr.push_back(gyr); }     // tmp_pre_integration->push_back
        e.dt_buf[j].push_back(dt); e.acc_buf[j].push_back(acc); e.gyr_buf[j].push_back(gyr);
        Vector3d g{0, 0, GRAVITY};
        Vector3d un_acc_0 = e.Rs[j] * (e.acc_0 - e.Bas[j]) - g;
        Vector3d un_gyr = 0.5 * (e.gyr_0 + gyr) - e.Bgs[j];
        e.Rs[j] *= Utility::deltaQ(un_gyr * dt).toRotationMatrix();
        Vector3d un_acc_1 = e.Rs[j] * (acc - e.Bas[j]) - g;
        Vector3d un_acc = 0.5 * (un_acc_0 + un_acc_1);
        e.Ps[j] += dt * e.Vs[j] + 0.5 * dt * dt * un_acc;
        e.Vs[j] += dt * un_acc;
    }
    e.acc_0 = acc; e.gyr_0 = gyr;
}

// FeatureManager::addFeatureCheckParallax, feature_manager.cpp:103-155 (+compensatedParallax2 :65-95)
bool add_feature_check_parallax(Est &e, int n, const int *ids, const double *xyz) {
    std::map<int, Vector3d> msg;                  // map<int,Vector3d> iteration = ascending id
    for (int i = 0; i < n; i++) msg[ids[i]] = Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    double parallax_sum = 0; int parallax_num = 0;
    e.last_track_num = 0;
    for (auto &kv : msg) {
        Vector3d p = kv.second / kv.second(2);
        auto it = std::find_if(e.feat.begin(), e.feat.end(), [&](const Track &t) { return t.id == kv.first; });
        if (it == e.feat.end()) { Track t; t.id = kv.first; t.start = e.frame_count; t.obs.push_back(p); e.feat.push_back(t); }
        else { it->obs.push_back(p); e.last_track_num++; }
    }
    int fc = e.frame_count;
    if (fc < 2 || e.last_track_num < 20) return true;
    for (auto &t : e.feat)
        if (t.start <= fc - 2 && t.end() >= fc - 1) {
            const Vector3d &pi = t.obs[fc - 2 - t.start], &pj = t.obs[fc - 1 - t.start];
            double du = pi(0) / pi(2) - pj(0), dv = pi(1) / pi(2) - pj(1);
            parallax_sum += std::sqrt(du * du + dv * dv);   // p_i_comp == p_i (COMPENSATE_ROTATION false)
            parallax_num++;
        }
    if (parallax_num == 0) return true;
    return parallax_sum / parallax_num >= e.c.min_parallax;
}

void triangulate(Est &e) {            // FeatureManager::triangulate, feature_manager.cpp:190-257
    for (auto &t : e.feat) {
        if (!in_solve(e, t) || t.depth > 0) continue;
        int imu_i = t.start, imu_j = imu_i - 1;
        MatrixXd A(2 * t.obs.size(), 4);
        int r = 0;
        Vector3d t0 = e.Ps[imu_i] + e.Rs[imu_i] * e.tic;
        Matrix3d R0 = e.Rs[imu_i] * e.ric;
        for (auto &o : t.obs) {
            imu_j++;
            Vector3d t1 = e.Ps[imu_j] + e.Rs[imu_j] * e.tic;
            Matrix3d R1 = e.Rs[imu_j] * e.ric;
            Vector3d tt = R0.transpose() * (t1 - t0);
            Matrix3d R = R0.transpose() * R1;
            Matrix<double, 3, 4> P;
            P.leftCols<3>() = R.transpose();
            P.rightCols<1>() = -R.transpose() * tt;
            Vector3d f = o.normalized();
            A.row(r++) = f[0] * P.row(2) - f[2] * P.row(0);
            A.row(r++) = f[1] * P.row(2) - f[2] * P.row(1);
        }
        Vector4d v = JacobiSVD<MatrixXd>(A, ComputeThinV).matrixV().rightCols<1>();
        t.depth = v[2] / v[3];
        if (t.depth < 0.1) t.depth = e.c.init_depth;
    }
}

}  // namespace
// oracle/init_ref.cpp: builds map<double, ImageFrame> from the arrays and calls the reference's VisualIMUAlignment (initial_aligment.cpp:222-229)
bool vref_align_frames(int n, const double *headers, const Eigen::Matrix3d *R, const Eigen::Vector3d *T, IntegrationBase **pre,
                       Eigen::Vector3d *Bgs, Eigen::Vector3d &g, Eigen::VectorXd &x);
namespace {

// VINS::visualInitialAlign (VINS.cpp:1022-1102).  ImageFrame::R / T of every frame of the map come from the caller's SfM (VINS.cpp:889-958).
bool visual_initial_align(Est &e) {
    const int W = e.W, n = (int)e.all.size();
    e.align_ok = 0;
    if ((int)e.sR.size() != n || n < 2) return false;            // the caller's frame list is not the map's
    std::vector<int> key(W + 1, -1);                              // map index of window frame i
    for (int i = 0, k = 0; i <= W; i++) {
        while (k < n && e.all[k].hdr != e.Headers[i]) k++;
        if (k >= n) return false;
        key[i] = k;
    }
    for (int i = 0; i <= W; i++) if (!e.pre[i]) return false;    // no IMU sample ever arrived for a frame (the reference would dereference NULL)
    TIC_X = e.tic.x(); TIC_Y = e.tic.y(); TIC_Z = e.tic.z();      // initial_aligment.cpp reads the globals (global_param.cpp:37-39 sets them per device)
    std::vector<IntegrationBase *> tmp(n);
    std::vector<double> hdrs(n);
    for (int k = 0; k < n; k++) {
        const Est::AllFrame &f = e.all[k];
        hdrs[k] = f.hdr;
        tmp[k] = new IntegrationBase{f.acc0, f.gyr0, Vector3d::Zero(), f.abg};
        for (size_t q = 0; q < f.dt.size(); q++) tmp[k]->push_back(f.dt[q], f.acc[q], f.gyr[q]);
    }
    Vector3d g; VectorXd x;
    const bool ok = vref_align_frames(n, hdrs.data(), e.sR.data(), e.sT.data(), tmp.data(), e.Bgs.data(), g, x);
    for (int k = 0; k < n; k++) { e.all[k].abg = e.Bgs[0]; delete tmp[k]; }
    e.align_ok = ok ? 1 : 0;
    if (!ok) return false;
    for (int i = 0; i <= e.frame_count; i++) { e.Ps[i] = e.sT[key[i]]; e.Rs[i] = e.sR[key[i]]; }
    for (auto &t : e.feat) t.depth = -1.0;                       // clearDepth(-1)
    { const Vector3d keep = e.tic; e.tic.setZero(); triangulate(e); e.tic = keep; }      // "triangulat on cam pose, no tic"
    const double s = (x.tail<1>())(0);
    e.scale_init = s;
    for (int i = 0; i <= W; i++) e.pre[i]->repropagate(Vector3d::Zero(), e.Bgs[i]);
    for (int i = e.frame_count; i >= 0; i--) e.Ps[i] = s * e.Ps[i] - e.Rs[i] * e.tic - (s * e.Ps[0] - e.Rs[0] * e.tic);
    for (int kv = 0; kv <= W; kv++) e.Vs[kv] = e.sR[key[kv]] * x.segment<3>(kv * 3);    // indexed by the keyframe counter, as written (VINS.cpp:1066-1075)
    for (auto &t : e.feat) { if (!in_solve(e, t)) continue; t.depth *= s; }
    Matrix3d R0 = Utility::g2R(g);
    const double yaw0 = Utility::R2ypr(R0).x();
    R0 = Utility::ypr2R(Vector3d{-yaw0, 0, 0}) * R0;
    g = R0 * g;
    e.g_init = g;
    for (int i = 0; i <= e.frame_count; i++) { e.Ps[i] = R0 * e.Ps[i]; e.Rs[i] = R0 * e.Rs[i]; e.Vs[i] = R0 * e.Vs[i]; }
    return true;
}

int feature_count(Est &e) { int s = 0; for (auto &t : e.feat) s += in_solve(e, t); return s; }

void old2new(Est &e) {               // VINS.cpp:89-129
    for (int i = 0; i <= e.W; i++) {
        double *p = e.pose(i), *s = e.sb(i);
        p[0] = e.Ps[i].x(); p[1] = e.Ps[i].y(); p[2] = e.Ps[i].z();
        Quaterniond q{e.Rs[i]};
        p[3] = q.x(); p[4] = q.y(); p[5] = q.z(); p[6] = q.w();
        for (int k = 0; k < 3; k++) { s[k] = e.Vs[i](k); s[3 + k] = e.Bas[i](k); s[6 + k] = e.Bgs[i](k); }
    }
    double *x = e.para_Ex.data();
    x[0] = e.tic.x(); x[1] = e.tic.y(); x[2] = e.tic.z();
    Quaterniond q{e.ric};
    x[3] = q.x(); x[4] = q.y(); x[5] = q.z(); x[6] = q.w();
    int k = 0;
    for (auto &t : e.feat) if (in_solve(e, t)) e.para_Feature[k++] = 1.0 / t.depth;
}

void new2old(Est &e) {               // VINS.cpp:131-212
    Vector3d origin_R0 = Utility::R2ypr(e.Rs[0]);
    Vector3d origin_P0 = e.Ps[0];
    if (e.failure_occur) { origin_R0 = Utility::R2ypr(e.last_R_old); origin_P0 = e.last_P_old; }
    double *p0 = e.pose(0);
    Vector3d origin_R00 = Utility::R2ypr(Quaterniond(p0[6], p0[3], p0[4], p0[5]).toRotationMatrix());
    double y_diff = origin_R0.x() - origin_R00.x();
    Matrix3d rot_diff = Utility::ypr2R(Vector3d(y_diff, 0, 0));
    for (int i = 0; i <= e.W; i++) {
        double *p = e.pose(i), *s = e.sb(i);
        e.Rs[i] = rot_diff * Quaterniond(p[6], p[3], p[4], p[5]).normalized().toRotationMatrix();
        e.Ps[i] = rot_diff * Vector3d(p[0] - p0[0], p[1] - p0[1], p[2] - p0[2]) + origin_P0;
        e.Vs[i] = rot_diff * Vector3d(s[0], s[1], s[2]);
        e.Bas[i] = Vector3d(s[3], s[4], s[5]);
        e.Bgs[i] = Vector3d(s[6], s[7], s[8]);
    }
    if (e.c.loop_closure && e.loop_enable) {                     // VINS.cpp:174-195: r_drift / t_drift from the re-anchored loop pose
        e.loop_enable = false;
        if (e.loop_frame >= 0) {
            Matrix3d Rs_loop = Quaterniond(e.loop_pose[6], e.loop_pose[3], e.loop_pose[4], e.loop_pose[5]).normalized().toRotationMatrix();
            Vector3d Ps_loop(e.loop_pose[0], e.loop_pose[1], e.loop_pose[2]);
            Rs_loop = rot_diff * Rs_loop;
            Ps_loop = rot_diff * (Ps_loop - Vector3d(p0[0], p0[1], p0[2])) + origin_P0;
            double drift_yaw = Utility::R2ypr(e.loop_Q_old.toRotationMatrix()).x() - Utility::R2ypr(Rs_loop).x();
            Matrix3d r_drift = Utility::ypr2R(Vector3d(drift_yaw, 0, 0));
            Vector3d t_drift = e.loop_P_old - r_drift * Ps_loop;
            e.loop_out[8] = drift_yaw;
            for (int k = 0; k < 3; k++) e.loop_out[9 + k] = t_drift(k);
        }
    }
    double *x = e.para_Ex.data();
    e.tic = Vector3d(x[0], x[1], x[2]);
    e.ric = Quaterniond(x[6], x[3], x[4], x[5]).toRotationMatrix();
    int k = 0;                        // FeatureManager::setDepth, feature_manager.cpp:331-349
    for (auto &t : e.feat) if (in_solve(e, t)) {
        t.depth = 1.0 / e.para_Feature[k++];
        t.solve_flag = t.depth < 0 ? 2 : 1;
    }
}

void marginalize_old(Est &e, ceres::LossFunction *loss) {     // VINS.cpp:690-774
    MarginalizationInfo *mi = new MarginalizationInfo();
    old2new(e);
    if (e.last_marg) {
        std::vector<int> drop;
        for (int i = 0; i < (int)e.last_marg_blocks.size(); i++)
            if (e.last_marg_blocks[i] == e.pose(0) || e.last_marg_blocks[i] == e.sb(0)) drop.push_back(i);
        mi->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(e.last_marg), NULL, e.last_marg_blocks, drop));
    }
    mi->addResidualBlockInfo(new ResidualBlockInfo(new IMUFactor(e.pre[1]), NULL,
        std::vector<double *>{e.pose(0), e.sb(0), e.pose(1), e.sb(1)}, std::vector<int>{0, 1}));
    int fi = -1;
    for (auto &t : e.feat) {
        if (!in_solve(e, t)) continue;
        ++fi;
        if (t.start != 0) continue;
        for (int k = 1; k < (int)t.obs.size(); k++)
            mi->addResidualBlockInfo(new ResidualBlockInfo(new ProjectionFactor(t.obs[0], t.obs[k]), loss,
                std::vector<double *>{e.pose(0), e.pose(k), e.para_Ex.data(), e.feat_p(fi)}, std::vector<int>{0, 3}));
    }
    mi->preMarginalize();
    mi->marginalize();
    std::unordered_map<long, double *> shift;
    for (int i = 1; i <= e.W; i++) {
        shift[reinterpret_cast<long>(e.pose(i))] = e.pose(i - 1);
        shift[reinterpret_cast<long>(e.sb(i))] = e.sb(i - 1);
    }
    shift[reinterpret_cast<long>(e.para_Ex.data())] = e.para_Ex.data();
    std::vector<double *> blocks = mi->getParameterBlocks(shift);
    delete e.last_marg;
    e.last_marg = mi; e.last_marg_blocks = blocks;
}

void marginalize_second_new(Est &e) {                          // VINS.cpp:776-830
    if (!(e.last_marg && std::count(e.last_marg_blocks.begin(), e.last_marg_blocks.end(), e.pose(e.W - 1)))) return;
    MarginalizationInfo *mi = new MarginalizationInfo();
    old2new(e);
    std::vector<int> drop;
    for (int i = 0; i < (int)e.last_marg_blocks.size(); i++)
        if (e.last_marg_blocks[i] == e.pose(e.W - 1)) drop.push_back(i);
    mi->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(e.last_marg), NULL, e.last_marg_blocks, drop));
    mi->preMarginalize();
    mi->marginalize();
    std::unordered_map<long, double *> shift;
    for (int i = 0; i <= e.W; i++) {
        if (i == e.W - 1) continue;
        int d = (i == e.W) ? i - 1 : i;
        shift[reinterpret_cast<long>(e.pose(i))] = e.pose(d);
        shift[reinterpret_cast<long>(e.sb(i))] = e.sb(d);
    }
    shift[reinterpret_cast<long>(e.para_Ex.data())] = e.para_Ex.data();
    std::vector<double *> blocks = mi->getParameterBlocks(shift);
    delete e.last_marg;
    e.last_marg = mi; e.last_marg_blocks = blocks;
}

void solve(Est &e) {                 // VINS::solve_ceres, VINS.cpp:480-831
    ceres::Problem problem;
    ceres::LossFunction *loss = new ceres::CauchyLoss(1.0);
    for (int i = 0; i <= e.W; i++) {
        problem.AddParameterBlock(e.pose(i), 7, new PoseLocalParameterization());
        problem.AddParameterBlock(e.sb(i), 9);
    }
    problem.AddParameterBlock(e.para_Ex.data(), 7, new PoseLocalParameterization());
    problem.SetParameterBlockConstant(e.para_Ex.data());
    for (int i = 0; i < e.c.num_of_f; i++) problem.AddParameterBlock(e.feat_p(i), 1);
    old2new(e);
    if (e.last_marg) problem.AddResidualBlock(new MarginalizationFactor(e.last_marg), NULL, e.last_marg_blocks);
    for (int i = 0; i < e.W; i++)
        problem.AddResidualBlock(new IMUFactor(e.pre[i + 1]), NULL, e.pose(i), e.sb(i), e.pose(i + 1), e.sb(i + 1));
    int fi = -1; e.n_proj = 0;
    for (auto &t : e.feat) {
        if (!in_solve(e, t)) continue;
        ++fi;
        for (int k = 1; k < (int)t.obs.size(); k++) {
            problem.AddResidualBlock(new ProjectionFactor(t.obs[0], t.obs[k]), loss,
                                     e.pose(t.start), e.pose(t.start + k), e.para_Ex.data(), e.feat_p(fi));
            e.n_proj++;
        }
    }
    e.n_feat_solve = fi + 1;
    // loop-closure factors, VINS.cpp:571-637 (LOOP_CLOSURE == cfg.loop_closure)
    e.loop_nfac = 0; e.loop_frame = -1;
    if (e.c.loop_closure && !e.loop_ids.empty() && e.loop_hdr >= e.Headers[0]) {
        for (int i = 0; i < e.W; i++) {
            if (e.loop_hdr != e.Headers[i]) continue;
            e.loop_frame = i;
            for (int k = 0; k < 7; k++) e.loop_pose[k] = e.pose(i)[k];
            problem.AddParameterBlock(e.loop_pose, 7, new PoseLocalParameterization());
            size_t ri = 0;
            int feature_index = -1;
            for (auto &t : e.feat) {
                if (!in_solve(e, t)) continue;
                ++feature_index;
                const int start = t.start, end = (int)(start + t.obs.size() - i - 1);
                if (start <= i && end >= 0) {
                    while (ri < e.loop_ids.size() && e.loop_ids[ri] < t.id) ri++;          // (the reference has no bound check here)
                    if (ri < e.loop_ids.size() && e.loop_ids[ri] == t.id) {
                        Vector3d pts_j(e.loop_xy[2 * ri], e.loop_xy[2 * ri + 1], 1.0);
                        problem.AddResidualBlock(new ProjectionFactor(t.obs[0], pts_j), loss, e.pose(start), e.loop_pose, e.para_Ex.data(), e.feat_p(feature_index));
                        ri++; e.loop_nfac++; e.loop_enable = true;
                    }
                }
            }
        }
    }
    ceres::Solver::Options o;
    o.linear_solver_type = ceres::DENSE_SCHUR;
    o.num_threads = 1;
    o.trust_region_strategy_type = ceres::DOGLEG;
    o.use_explicit_schur_complement = true;
    o.minimizer_progress_to_stdout = false;
    o.max_num_iterations = e.c.max_iters;
    o.logging_type = ceres::SILENT;
    ceres::Solver::Summary sum;
    {
        const auto t0 = std::chrono::steady_clock::now();
        ceres::Solve(o, &problem, &sum);
        e.t_stage[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    e.cost0 = sum.initial_cost; e.cost1 = sum.final_cost;
    e.iters = (int)sum.iterations.size() - 1;
    for (int k = 0; k < 12; k++) e.loop_out[k] = 0;
    if (e.loop_frame >= 0 && e.loop_nfac > 0) {                  // VINS.cpp:664-680
        const int i = e.loop_frame;
        const double *pp = e.pose(i);
        Matrix3d Rs_i = Quaterniond(pp[6], pp[3], pp[4], pp[5]).normalized().toRotationMatrix();
        Vector3d Ps_i(pp[0], pp[1], pp[2]);
        Matrix3d Rs_loop = Quaterniond(e.loop_pose[6], e.loop_pose[3], e.loop_pose[4], e.loop_pose[5]).normalized().toRotationMatrix();
        Vector3d Ps_loop(e.loop_pose[0], e.loop_pose[1], e.loop_pose[2]);
        Vector3d rt = Rs_loop.transpose() * (Ps_i - Ps_loop);
        Quaterniond rq(Rs_loop.transpose() * Rs_i);
        for (int k = 0; k < 3; k++) e.loop_out[k] = rt(k);
        e.loop_out[3] = rq.x(); e.loop_out[4] = rq.y(); e.loop_out[5] = rq.z(); e.loop_out[6] = rq.w();
        e.loop_out[7] = Utility::normalizeAngle(Utility::R2ypr(Rs_i).x() - Utility::R2ypr(Rs_loop).x());
    }
    new2old(e);
    e.post_solve.assign((e.W + 1) * 16, 0.0);
    for (int i = 0; i <= e.W; i++) {
        double *d = &e.post_solve[16 * i];
        Quaterniond q{e.Rs[i]};
        for (int k = 0; k < 3; k++) { d[k] = e.Ps[i](k); d[7 + k] = e.Vs[i](k); d[10 + k] = e.Bas[i](k); d[13 + k] = e.Bgs[i](k); }
        d[3] = q.x(); d[4] = q.y(); d[5] = q.z(); d[6] = q.w();
    }
    std::vector<ceres::ResidualBlockId> rs;
    problem.GetResidualBlocks(&rs);
    // Problem owns cost functions; the MarginalizationFactor added above must not delete last_marg (it does not).
    {
        const auto t0 = std::chrono::steady_clock::now();
        if (e.marg_flag == 0) marginalize_old(e, loss); else marginalize_second_new(e);
        e.t_stage[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
}

bool failure_detection(Est &e) {     // VINS.cpp:214-265
    bool f = false;
    if (e.last_track_num < 4) f = true;
    if (e.Bgs[e.W].norm() > 1) f = true;
    Vector3d tp = e.Ps[e.W];
    if ((tp - e.last_P).norm() > 1) f = true;
    if (std::abs(tp.z() - e.last_P.z()) > 0.5) f = true;
    Matrix3d dR = e.Rs[e.W].transpose() * e.last_R;
    Quaterniond dq(dR);
    double ang = std::acos(dq.w()) * 2.0 / 3.14 * 180.0;
    if (ang > 40) f = true;
    return f;
}

void slide_window(Est &e) {          // VINS.cpp:1149-1273 + feature_manager.cpp:259-287,356-406
    int W = e.W;
    if (e.marg_flag == 0) {
        e.back_R0 = e.Rs[0]; e.back_P0 = e.Ps[0];
        if (e.frame_count != W) return;
        for (int i = 0; i < W; i++) {
            e.Rs[i].swap(e.Rs[i + 1]);
            std::swap(e.pre[i], e.pre[i + 1]);
            e.dt_buf[i].swap(e.dt_buf[i + 1]); e.acc_buf[i].swap(e.acc_buf[i + 1]); e.gyr_buf[i].swap(e.gyr_buf[i + 1]);
            e.Headers[i] = e.Headers[i + 1];
            e.Ps[i].swap(e.Ps[i + 1]); e.Vs[i].swap(e.Vs[i + 1]);
        }
        e.Headers[W] = e.Headers[W - 1]; e.Ps[W] = e.Ps[W - 1]; e.Vs[W] = e.Vs[W - 1]; e.Rs[W] = e.Rs[W - 1];
        e.Bas[W] = e.Bas[W - 1]; e.Bgs[W] = e.Bgs[W - 1];          // Q9: Bas/Bgs are NOT shifted for i < W
        if (e.solver_flag == 0) {                                   // all_image_frame.erase(begin, find(Headers[0])), VINS.cpp:1186-1193
            size_t m = 0;
            while (m < e.all.size() && e.all[m].hdr != e.Headers[0]) m++;
            if (m < e.all.size()) e.all.erase(e.all.begin(), e.all.begin() + m);
        }
        delete e.pre[W];
        e.pre[W] = new IntegrationBase{e.acc_0, e.gyr_0, e.Bas[W], e.Bgs[W]};
        e.dt_buf[W].clear(); e.acc_buf[W].clear(); e.gyr_buf[W].clear();
        // slideWindowOld
        if (e.solver_flag == 1) {
            Matrix3d R0 = e.back_R0 * e.ric, R1 = e.Rs[0] * e.ric;
            Vector3d P0 = e.back_P0 + e.back_R0 * e.tic, P1 = e.Ps[0] + e.Rs[0] * e.tic;
            std::vector<Track> out;
            for (auto &t : e.feat) {
                if (t.start != 0) { t.start--; out.push_back(t); continue; }
                Vector3d uv = t.obs[0];
                t.obs.erase(t.obs.begin());
                if (t.obs.size() < 2) continue;
                Vector3d w = R0 * (uv * t.depth) + P0;
                double dep = (R1.transpose() * (w - P1))(2);
                t.depth = dep > 0 ? dep : e.c.init_depth;
                out.push_back(t);
            }
            e.feat.swap(out);
        } else {
            std::vector<Track> out;
            for (auto &t : e.feat) {
                if (t.start != 0) { t.start--; out.push_back(t); continue; }
                t.obs.erase(t.obs.begin());
                if (t.obs.size() == 0) continue;
                out.push_back(t);
            }
            e.feat.swap(out);
        }
    } else {
        if (e.frame_count != W) return;
        for (size_t i = 0; i < e.dt_buf[W].size(); i++) {
            e.pre[W - 1]->push_back(e.dt_buf[W][i], e.acc_buf[W][i], e.gyr_buf[W][i]);
            e.dt_buf[W - 1].push_back(e.dt_buf[W][i]);
            e.acc_buf[W - 1].push_back(e.acc_buf[W][i]);
            e.gyr_buf[W - 1].push_back(e.gyr_buf[W][i]);
        }
        e.Headers[W - 1] = e.Headers[W]; e.Ps[W - 1] = e.Ps[W]; e.Vs[W - 1] = e.Vs[W]; e.Rs[W - 1] = e.Rs[W];
        e.Bas[W - 1] = e.Bas[W]; e.Bgs[W - 1] = e.Bgs[W];
        delete e.pre[W];
        e.pre[W] = new IntegrationBase{e.acc_0, e.gyr_0, e.Bas[W], e.Bgs[W]};
        e.dt_buf[W].clear(); e.acc_buf[W].clear(); e.gyr_buf[W].clear();
        // slideWindowNew -> removeFront(frame_count)
        std::vector<Track> out;
        for (auto &t : e.feat) {
            if (t.start == e.frame_count) { t.start--; out.push_back(t); continue; }
            if (t.end() < e.frame_count - 1) { out.push_back(t); continue; }
            int j = W - 1 - t.start;
            t.obs.erase(t.obs.begin() + j);
            if (t.obs.size() == 0) continue;
            out.push_back(t);
        }
        e.feat.swap(out);
    }
}

void remove_failures(Est &e) {       // feature_manager.cpp:289-298
    e.feat.erase(std::remove_if(e.feat.begin(), e.feat.end(), [](const Track &t) { return t.solve_flag == 2; }), e.feat.end());
}

int process_image(Est &e, int n, const int *ids, const double *xyz, double header) {   // VINS.cpp:377-478
    e.marg_flag = add_feature_check_parallax(e, n, ids, xyz) ? 0 : 1;
    e.Headers[e.frame_count] = header;
    if (e.solver_flag == 0) {
        e.tmp.hdr = header; e.all.push_back(e.tmp);           // all_image_frame.insert; tmp_pre_integration = new IntegrationBase{acc_0, gyr_0, 0, 0}
        e.tmp = Est::AllFrame(); e.tmp.acc0 = e.acc_0; e.tmp.gyr0 = e.gyr_0;
        if (e.frame_count == e.W) {
            if (e.last_track_num < 20) { clear_state(e); return 2; }      // VINS.cpp:401-405
            if (e.sfm_pending) {                            // solveInitial() from the SfM poses on: VINS.cpp:1022-1102, then :415-447
                e.sfm_pending = false;
                if (visual_initial_align(e)) {
                    solve(e);
                    if (e.cost1 > 200) {
                        delete e.last_marg; e.last_marg = nullptr;
                        e.solver_flag = 0;
                        slide_window(e);
                    } else {
                        e.failure_occur = 0;
                        e.solver_flag = 1;
                        slide_window(e);
                        remove_failures(e);
                        e.last_R = e.Rs[e.W]; e.last_P = e.Ps[e.W]; e.last_R_old = e.Rs[0]; e.last_P_old = e.Ps[0];
                    }
                } else {
                    slide_window(e);
                }
            } else if (e.init_pending) {
                e.init_pending = false;
                for (int i = 0; i <= e.W; i++) { e.Ps[i] = e.iP[i]; e.Rs[i] = e.iQ[i].normalized().toRotationMatrix(); e.Vs[i] = e.iV[i]; e.Bas[i] = e.iBa; e.Bgs[i] = e.iBg; }
                for (auto &t : e.feat) t.depth = -1.0;       // clearDepth(-1), VINS.cpp:1047-1050
                triangulate(e);
                solve(e);
                if (e.cost1 > 200) {                         // VINS.cpp:416-425: initialisation rejected
                    delete e.last_marg; e.last_marg = nullptr;
                    e.solver_flag = 0;
                    slide_window(e);
                } else {
                    e.failure_occur = 0;
                    e.solver_flag = 1;
                    slide_window(e);
                    remove_failures(e);
                    e.last_R = e.Rs[e.W]; e.last_P = e.Ps[e.W]; e.last_R_old = e.Rs[0]; e.last_P_old = e.Ps[0];
                }
            } else {
                slide_window(e);
            }
        } else {
            e.frame_count++;
        }
    } else {
        triangulate(e);
        solve(e);
        e.failure_occur = 0;
        if (failure_detection(e)) { e.failure_occur = 1; clear_state(e); return 1; }
        slide_window(e);
        remove_failures(e);
        e.last_R = e.Rs[e.W]; e.last_P = e.Ps[e.W]; e.last_R_old = e.Rs[0]; e.last_P_old = e.Ps[0];
    }
    return 0;
}

}  // namespace

extern "C" {

void *vref_create(const vio_config *cfg) {
    if (std::fabs(cfg->acc_n - ACC_N) > 0 || std::fabs(cfg->acc_w - ACC_W) > 0 || std::fabs(cfg->gyr_n - GYR_N) > 0 ||
        std::fabs(cfg->gyr_w - GYR_W) > 0 || std::fabs(cfg->gravity - GRAVITY) > 0) {
        fprintf(stderr, "vref_create: noise/gravity are compile-time macros in the reference (global_param.hpp:42-46)\n");
        return nullptr;
    }
    Est *e = new Est();
    e->c = *cfg; e->W = cfg->window_size;
    {
        const size_t nP = 7 * (e->W + 1), nS = 9 * (e->W + 1), nF = cfg->num_of_f, nE = 7;
        double *a = ParaArena::claim(&e->arena_slot, nP + nS + nF + nE);
        if (!a) { fprintf(stderr, "vref_create: no parameter arena slot free\n"); delete e; return nullptr; }
        e->para_Pose.p = a; e->para_Pose.n = nP; e->para_SB.p = a + nP; e->para_SB.n = nS;
        e->para_Feature.p = a + nP + nS; e->para_Feature.n = nF; e->para_Ex.p = a + nP + nS + nF; e->para_Ex.n = nE;      // mmap memory is zeroed
    }
    FOCUS_LENGTH_X = cfg->fx; FOCUS_LENGTH_Y = cfg->fy; PX = cfg->cx; PY = cfg->cy;
    ProjectionFactor::sqrt_info = cfg->fx / 1.5 * Matrix2d::Identity();     // VINS::setIMUModel, VINS.cpp:29-32
    e->last_P.setZero(); e->last_R.setIdentity(); e->last_P_old.setZero(); e->last_R_old.setIdentity();
    clear_state(*e);
    return e;
}
void vref_destroy(void *h) { Est *e = (Est *)h; clear_state(*e); ParaArena::release(e->arena_slot); delete e; }
void vref_clear(void *h) { clear_state(*(Est *)h); }
void vref_process_imu(void *h, double dt, const double *a, const double *g) {
    const auto t0 = std::chrono::steady_clock::now();
    process_imu(*(Est *)h, dt, Vector3d(a[0], a[1], a[2]), Vector3d(g[0], g[1], g[2]));
    ((Est *)h)->t_stage[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// retrive_pose_data: header, n ids (ascending) + n x 2 measurements of the old keyframe, its pose P_old[3], Q_old (xyzw); n = 0 clears
void vref_set_loop_match(void *h, int n, double header, const int *ids, const double *xy, const double *pose_old) {
    Est &e = *(Est *)h;
    e.loop_hdr = header;
    e.loop_ids.assign(ids, ids + n); e.loop_xy.assign(xy, xy + 2 * n);
    e.loop_P_old = Vector3d(pose_old[0], pose_old[1], pose_old[2]);
    e.loop_Q_old = Quaterniond(pose_old[6], pose_old[3], pose_old[4], pose_old[5]);
}
int vref_get_loop_result(void *h, double *out) {
    Est &e = *(Est *)h;
    const bool valid = e.loop_frame >= 0 && e.loop_nfac > 0;
    for (int k = 0; k < 12; k++) out[k] = valid ? e.loop_out[k] : 0.0;
    return valid ? e.loop_nfac : 0;
}
// VINS::solve_ceres() alone on the current window (NON_LINEAR, full window), as vio_backend_solve does
int vref_solve(void *h) {
    Est &e = *(Est *)h;
    if (e.solver_flag != 1 || e.frame_count != e.W) return 1;
    solve(e);
    return 0;
}
// seconds spent so far in [0] processImage, [1] ceres::Solve, [2] marginalisation, [3] processIMU; reset != 0 zeroes them
void vref_stage_seconds(void *h, double *out, int reset) {
    Est &e = *(Est *)h;
    for (int i = 0; i < 4; i++) { out[i] = e.t_stage[i]; if (reset) e.t_stage[i] = 0; }
}
void vref_set_init_window(void *h, const double *P, const double *Q, const double *V, const double *Ba, const double *Bg) {
    Est &e = *(Est *)h;
    e.iP.clear(); e.iQ.clear(); e.iV.clear();
    for (int i = 0; i <= e.W; i++) {
        e.iP.emplace_back(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
        e.iQ.emplace_back(Q[4 * i + 3], Q[4 * i], Q[4 * i + 1], Q[4 * i + 2]);
        e.iV.emplace_back(V[3 * i], V[3 * i + 1], V[3 * i + 2]);
    }
    e.iBa = Vector3d(Ba[0], Ba[1], Ba[2]); e.iBg = Vector3d(Bg[0], Bg[1], Bg[2]);
    e.init_pending = true;
}
void vref_set_init_sfm_frames(void *h, int n, const double *R, const double *T) {
    Est &e = *(Est *)h;
    e.sR.clear(); e.sT.clear();
    for (int i = 0; i < n; i++) {
        e.sR.push_back(Map<const Matrix<double, 3, 3, RowMajor>>(R + 9 * i));
        e.sT.emplace_back(T[3 * i], T[3 * i + 1], T[3 * i + 2]);
    }
    e.sfm_pending = true;
}
void vref_set_init_sfm(void *h, const double *R, const double *T) { vref_set_init_sfm_frames(h, ((Est *)h)->W + 1, R, T); }
int vref_get_init_frames(void *h, int cap, double *headers) {
    Est &e = *(Est *)h;
    const int n = (int)e.all.size();
    for (int i = 0; i < n && i < cap; i++) headers[i] = e.all[i].hdr;
    return n;
}
void vref_get_init_result(void *h, int *ok, double *g, double *scale) {
    Est &e = *(Est *)h;
    *ok = e.align_ok; *scale = e.scale_init;
    for (int i = 0; i < 3; i++) g[i] = e.g_init[i];
}
int vref_process_image(void *h, int n, const int *ids, const double *xyz, double header) {
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = process_image(*(Est *)h, n, ids, xyz, header);
    ((Est *)h)->t_stage[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}
static void pack_state(Est &e, double *P, double *Q, double *V, double *Ba, double *Bg, double *H) {
    for (int i = 0; i <= e.W; i++) {
        Quaterniond q{e.Rs[i]};
        if (Q) { Q[4 * i] = q.x(); Q[4 * i + 1] = q.y(); Q[4 * i + 2] = q.z(); Q[4 * i + 3] = q.w(); }
        for (int k = 0; k < 3; k++) {
            if (P) P[3 * i + k] = e.Ps[i](k);
            if (V) V[3 * i + k] = e.Vs[i](k);
            if (Ba) Ba[3 * i + k] = e.Bas[i](k);
            if (Bg) Bg[3 * i + k] = e.Bgs[i](k);
        }
        if (H) H[i] = e.Headers[i];
    }
}
void vref_get_state(void *h, double *P, double *Q, double *V, double *Ba, double *Bg, double *H) { pack_state(*(Est *)h, P, Q, V, Ba, Bg, H); }
int vref_get_post_solve(void *h, double *out) {
    Est &e = *(Est *)h;
    if (e.post_solve.empty()) return 1;
    memcpy(out, e.post_solve.data(), sizeof(double) * e.post_solve.size());
    return 0;
}
void vref_get_info(void *h, int *info, double *dinfo) {
    Est &e = *(Est *)h;
    info[0] = e.solver_flag; info[1] = e.marg_flag; info[2] = e.frame_count; info[3] = e.failure_occur;
    info[4] = e.n_feat_solve; info[5] = e.n_proj; info[6] = e.iters; info[7] = e.last_track_num;
    dinfo[0] = e.cost0; dinfo[1] = e.cost1; dinfo[2] = e.last_marg ? e.last_marg->n : 0; dinfo[3] = 0;
}
int vref_get_features(void *h, int cap, int *n_out, int *ids, int *start, int *nobs, double *depth, int *flag) {
    Est &e = *(Est *)h;
    int n = (int)e.feat.size();
    *n_out = n;
    for (int i = 0; i < n && i < cap; i++) {
        ids[i] = e.feat[i].id; start[i] = e.feat[i].start; nobs[i] = (int)e.feat[i].obs.size();
        depth[i] = e.feat[i].depth; flag[i] = e.feat[i].solve_flag;
    }
    return n <= cap ? 0 : 1;
}
// Prior in information form over the canonical layout [pose_i(6) sb_i(9)]_{i=0..W} ex(6).
int vref_get_prior(void *h, double *H, double *b, int *present, double *c0) {
    Est &e = *(Est *)h;
    int N = 15 * (e.W + 1) + 6;
    std::fill(H, H + N * N, 0.0); std::fill(b, b + N, 0.0);
    for (int i = 0; i < 2 * (e.W + 1) + 1; i++) present[i] = 0;
    *c0 = 0;
    if (!e.last_marg) return 1;
    MarginalizationInfo *mi = e.last_marg;
    int n = mi->n, m = mi->m;
    std::vector<int> col(n, -1);
    for (size_t k = 0; k < e.last_marg_blocks.size(); k++) {
        double *a = e.last_marg_blocks[k];
        int base = -1, blk = -1, ls = mi->keep_block_size[k] == 7 ? 6 : mi->keep_block_size[k];
        for (int i = 0; i <= e.W; i++) {
            if (a == e.pose(i)) { base = 15 * i; blk = 2 * i; }
            if (a == e.sb(i)) { base = 15 * i + 6; blk = 2 * i + 1; }
        }
        if (a == e.para_Ex.data()) { base = 15 * (e.W + 1); blk = 2 * (e.W + 1); }
        if (base < 0) return 2;
        present[blk] = 1;
        for (int j = 0; j < ls; j++) col[mi->keep_block_idx[k] - m + j] = base + j;
    }
    MatrixXd A = mi->linearized_jacobians.transpose() * mi->linearized_jacobians;
    VectorXd g = mi->linearized_jacobians.transpose() * mi->linearized_residuals;
    for (int i = 0; i < n; i++) {
        if (col[i] < 0) return 3;
        b[col[i]] = g(i);
        for (int j = 0; j < n; j++) H[col[i] * N + col[j]] = A(i, j);
    }
    *c0 = mi->linearized_residuals.squaredNorm();
    return 0;
}

// ---- factor-level primitives (reference code, unmodified) -------------------------------------------------
void vref_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *acc0, const double *gyr0,
                       const double *ba, const double *bg, double *pqv, double *jac, double *cov, double *sum_dt) {
    IntegrationBase ib{Vector3d(acc0[0], acc0[1], acc0[2]), Vector3d(gyr0[0], gyr0[1], gyr0[2]),
                       Vector3d(ba[0], ba[1], ba[2]), Vector3d(bg[0], bg[1], bg[2])};
    for (int i = 0; i < n; i++)
        ib.push_back(dt[i], Vector3d(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]), Vector3d(gyr[3 * i], gyr[3 * i + 1], gyr[3 * i + 2]));
    for (int k = 0; k < 3; k++) { pqv[k] = ib.delta_p(k); pqv[7 + k] = ib.delta_v(k); }
    pqv[3] = ib.delta_q.x(); pqv[4] = ib.delta_q.y(); pqv[5] = ib.delta_q.z(); pqv[6] = ib.delta_q.w();
    Map<Matrix<double, 15, 15, RowMajor>> jm(jac), cm(cov);
    jm = ib.jacobian;
    cm = ib.covariance;
    *sum_dt = ib.sum_dt;
}
void vref_imu_factor(const double *pqv, const double *jac, const double *cov, double sum_dt, const double *lba, const double *lbg,
                     const double *pi, const double *sbi, const double *pj, const double *sbj, double *res, double *J) {
    IntegrationBase ib{Vector3d::Zero(), Vector3d::Zero(), Vector3d(lba[0], lba[1], lba[2]), Vector3d(lbg[0], lbg[1], lbg[2])};
    ib.delta_p = Vector3d(pqv[0], pqv[1], pqv[2]);
    ib.delta_q = Quaterniond(pqv[6], pqv[3], pqv[4], pqv[5]);
    ib.delta_v = Vector3d(pqv[7], pqv[8], pqv[9]);
    ib.jacobian = Map<const Matrix<double, 15, 15, RowMajor>>(jac);
    ib.covariance = Map<const Matrix<double, 15, 15, RowMajor>>(cov);
    ib.sum_dt = sum_dt;
    IMUFactor f(&ib);
    const double *params[4] = {pi, sbi, pj, sbj};
    double j0[15 * 7], j1[15 * 9], j2[15 * 7], j3[15 * 9];
    double *jac_out[4] = {j0, j1, j2, j3};
    f.Evaluate(params, res, jac_out);
    for (int r = 0; r < 15; r++) {
        for (int c = 0; c < 6; c++) { J[r * 30 + c] = j0[r * 7 + c]; J[r * 30 + 15 + c] = j2[r * 7 + c]; }
        for (int c = 0; c < 9; c++) { J[r * 30 + 6 + c] = j1[r * 9 + c]; J[r * 30 + 21 + c] = j3[r * 9 + c]; }
    }
}
// sqrt_info exactly as IMUFactor::Evaluate forms it (imu_factor.h:72), row-major 15 x 15 (upper triangular)
void vref_imu_sqrt_info(const double *cov, double *U) {
    Matrix<double, 15, 15> covariance = Map<const Matrix<double, 15, 15, RowMajor>>(cov);
    Matrix<double, 15, 15> sqrt_info = LLT<Matrix<double, 15, 15>>(covariance.inverse()).matrixL().transpose();
    Map<Matrix<double, 15, 15, RowMajor>> out(U);
    out = sqrt_info;
}
void vref_projection_factor(double fx, const double *tic, const double *ric, const double *pts_i, const double *pts_j,
                            const double *pi, const double *pj, double inv_dep, double *res, double *J) {
    ProjectionFactor::sqrt_info = fx / 1.5 * Matrix2d::Identity();
    ProjectionFactor f(Vector3d(pts_i[0], pts_i[1], pts_i[2]), Vector3d(pts_j[0], pts_j[1], pts_j[2]));
    Matrix3d R = Map<const Matrix<double, 3, 3, RowMajor>>(ric);
    Quaterniond q{R};
    double ex[7] = {tic[0], tic[1], tic[2], q.x(), q.y(), q.z(), q.w()};
    const double *params[4] = {pi, pj, ex, &inv_dep};
    double j0[14], j1[14], j2[14], j3[2];
    double *jac_out[4] = {j0, j1, j2, j3};
    f.Evaluate(params, res, jac_out);
    for (int r = 0; r < 2; r++) {
        for (int c = 0; c < 6; c++) { J[r * 13 + c] = j0[r * 7 + c]; J[r * 13 + 6 + c] = j1[r * 7 + c]; }
        J[r * 13 + 12] = j3[r];
    }
}

}  // extern "C"
