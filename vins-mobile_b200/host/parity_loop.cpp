// The estimator loop of ViewController.mm:364-494 / :701-882 over the host mirror (vio_host.hpp), fed from a file of recorded inputs, so
// that the C++ classes are EXECUTED on a GPU and compared with the ctypes path (tests/test_pipeline_gpu.py::test_cpp_host_mirror_runs).
//   file: int32 {n_frames, rows, cols, imu_per_kf, window, max_cnt}, then per frame rows*cols u8, then per keyframe interval (one more
//         than the frames need: it feeds the stand-alone solve_ceres at the end) imu_per_kf x {dt, acc[3], gyr[3]} f64, then the initial window P[(W+1)*3], Q[(W+1)*4] (xyzw), V[(W+1)*3] f64.
//   stdout per published frame: "kf <frame> <frame_count> <solver_flag> <n_ids> <sum ids> <P of the newest frame> <final_cost>"
#include <cstdio>
#include <cstdint>
#include <vector>

#include "vio_host.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 3;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hd[6];
    if (std::fread(hd, sizeof(int32_t), 6, f) != 6) return 3;
    const int n_frames = hd[0], rows = hd[1], cols = hd[2], per = hd[3], W = hd[4], max_cnt = hd[5];
    std::vector<uint8_t> frames((size_t)n_frames * rows * cols);
    if (std::fread(frames.data(), 1, frames.size(), f) != frames.size()) return 3;
    const int n_kf = (n_frames + 2) / 3;
    std::vector<double> imu((size_t)n_kf * per * 7), P((W + 1) * 3), Q((W + 1) * 4), V((W + 1) * 3);
    if (std::fread(imu.data(), sizeof(double), imu.size(), f) != imu.size()) return 3;
    if (std::fread(P.data(), sizeof(double), P.size(), f) != P.size() || std::fread(Q.data(), sizeof(double), Q.size(), f) != Q.size() ||
        std::fread(V.data(), sizeof(double), V.size(), f) != V.size()) return 3;
    std::fclose(f);
    vio_config cfg;
    vio_config_default(&cfg);
    cfg.rows = rows; cfg.cols = cols; cfg.window_size = W; cfg.max_cnt = max_cnt;
    try {
        vio::FeatureTracker featuretracker(cfg);
        vio::VINS vins(cfg);
        vins.setIMUModel();
        vins.setExtrinsic();
        vio::Vector3d Pd{0, 0, 0};
        vio::Matrix3d Rd{};
        const double zero[3] = {0, 0, 0};
        int kf = 0;
        for (int k = 0; k < n_frames; k++) {
            vio::Mat img{frames.data() + (size_t)k * rows * cols, rows, cols}, result{nullptr, 0, 0};
            std::vector<vio::Point2f> good_pts;
            std::vector<double> track_len;
            const bool publish = featuretracker.img_cnt == 0;
            featuretracker.readImage(img, result, k, good_pts, track_len, k / 30.0, Pd, Rd, false);
            if (!publish) continue;
            if (kf > 0)
                for (int j = 0; j < per; j++) {
                    const double *m = &imu[((size_t)(kf - 1) * per + j) * 7];
                    vins.processIMU(m[0], vio::Vector3d{m[1], m[2], m[3]}, vio::Vector3d{m[4], m[5], m[6]});
                }
            if (kf == W) vins.setInitialWindow(P.data(), Q.data(), V.data(), zero, zero);
            vins.processImage(featuretracker.image_msg, k / 30.0, 0);
            long long sum = 0;
            for (int id : featuretracker.ids) sum += id;
            const vio::Vector3d &pw = vins.Ps[W];
            std::printf("kf %d %d %d %zu %lld %.17g %.17g %.17g %.17g\n", k, vins.frame_count, (int)vins.solver_flag, featuretracker.ids.size(), sum, pw.x, pw.y, pw.z,
                        vins.final_cost);
            kf++;
        }
        // VINS::solve_ceres on its own: the IMU samples of the next interval, then a solve of the window as it stands
        for (int j = 0; j < per; j++) {
            const double *m = &imu[((size_t)(kf - 1) * per + j) * 7];
            vins.processIMU(m[0], vio::Vector3d{m[1], m[2], m[3]}, vio::Vector3d{m[4], m[5], m[6]});
        }
        vins.solve_ceres(0);
        std::printf("resolve %.17g %.17g\n", vins.Ps[W].x, vins.final_cost);
    } catch (const std::exception &e) {
        std::printf("no CUDA device / error: %s\n", e.what());
        return 2;
    }
    return 0;
}
