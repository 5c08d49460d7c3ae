// oracle/init_ref.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
// C API (ctypes) over the reference's UNMODIFIED VisualIMUAlignment() (VINS_ios/initial_aligment.cpp:222-229: solveGyroscopeBias :10-46,
// SolveScale :135-220, RefineGravity :64-133), compiled from /root/reference where it lies.  This file only builds the
// map<double, ImageFrame> the way VINS::processImage does (VINS.cpp:412-416: one ImageFrame per camera frame carrying the IMU
// pre-integration since the previous frame, IntegrationBase{acc_0, gyr_0, Bas, Bgs} + push_back per sample, VINS.cpp:340-352) and hands it over.
#include <map>
#include <vector>
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>
#include "initial_aligment.hpp"

// Used by oracle/backend_ref.cpp (visual_initial_align): the window's frames as all_image_frame, then the reference's VisualIMUAlignment.
// pre[k] are the all_image_frame pre-integration objects (owned by the caller); Bgs has WINDOW_SIZE + 1 entries.
bool vref_align_frames(int n, const double *headers, const Eigen::Matrix3d *R, const Eigen::Vector3d *T, IntegrationBase **pre,
                       Eigen::Vector3d *Bgs, Eigen::Vector3d &g, Eigen::VectorXd &x) {
    std::map<double, ImageFrame> all;
    for (int k = 0; k < n; k++) {
        std::map<int, Eigen::Vector3d> pts;
        ImageFrame f(pts, headers[k]);
        f.R = R[k]; f.T = T[k]; f.pre_integration = pre[k];
        all.insert(std::make_pair(headers[k], f));
    }
    fflush(stdout);
    const int saved = dup(1), nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    const bool ok = VisualIMUAlignment(all, Bgs, g, x);
    std::cout.flush(); fflush(stdout);
    dup2(saved, 1); close(nul); close(saved);
    return ok;
}

extern "C" {

// n frames; R [n][9] row-major (ImageFrame::R), T [n][3] (ImageFrame::T); counts [n] samples of the interval ENDING at frame k
// (k = 0: the pre-integration exists but is never read); imu0 [n][6] acc_0, gyr_0 the interval starts from;
// imu [n][max_imu][7] (dt, acc, gyr); bg0 [3] = Bgs[*] before the call, tic [3] = TIC_X/Y/Z.
// out: bgs [3] = Bgs[0] after, g [3], x [3 n + 3] (the caller's VectorXd; 3 n + 3 entries are valid after RefineGravity), returns ok.
int vref_visual_imu_align(int n, const double *R, const double *T, const int *counts, const double *imu0, int max_imu, const double *imu,
                          const double *bg0, const double *tic, double *bgs_out, double *g_out, double *x_out) {
    TIC_X = tic[0]; TIC_Y = tic[1]; TIC_Z = tic[2];
    std::map<double, ImageFrame> all;
    Eigen::Vector3d Bgs[WINDOW_SIZE + 1];
    for (int i = 0; i <= WINDOW_SIZE; i++) Bgs[i] = Eigen::Vector3d(bg0[0], bg0[1], bg0[2]);
    std::vector<IntegrationBase *> own;
    for (int k = 0; k < n; k++) {
        std::map<int, Eigen::Vector3d> pts;
        ImageFrame f(pts, (double)k);
        f.R = Eigen::Map<const Eigen::Matrix<double, 3, 3, Eigen::RowMajor>>(R + 9 * k);
        f.T = Eigen::Map<const Eigen::Vector3d>(T + 3 * k);
        const double *a0 = imu0 + 6 * k;
        IntegrationBase *p = new IntegrationBase(Eigen::Vector3d(a0[0], a0[1], a0[2]), Eigen::Vector3d(a0[3], a0[4], a0[5]),
                                                 Eigen::Vector3d::Zero(), Bgs[0]);
        for (int i = 0; i < counts[k]; i++) {
            const double *e = imu + ((size_t)k * max_imu + i) * 7;
            p->push_back(e[0], Eigen::Vector3d(e[1], e[2], e[3]), Eigen::Vector3d(e[4], e[5], e[6]));
        }
        f.pre_integration = p;
        own.push_back(p);
        all.insert(std::make_pair((double)k, f));
    }
    Eigen::Vector3d g;
    Eigen::VectorXd x;
    fflush(stdout);
    const int saved = dup(1), nul = open("/dev/null", O_WRONLY);      // the reference prints through both printf and cout
    dup2(nul, 1);
    const bool ok = VisualIMUAlignment(all, Bgs, g, x);
    std::cout.flush(); fflush(stdout);
    dup2(saved, 1); close(nul); close(saved);
    for (int i = 0; i < 3; i++) { bgs_out[i] = Bgs[0][i]; g_out[i] = g[i]; }
    for (int i = 0; i < 3 * n + 3; i++) x_out[i] = i < x.size() ? x[i] : 0.0;
    for (auto *p : own) delete p;
    return ok ? 1 : 0;
}

}  // extern "C"
