// Minimal estimator loop over the host mirror, shaped like ViewController.mm:364-494 (camera thread) + :701-882 (mainLoop).
// Also the link test for libvio_b200.so (tests/test_cpu_abi_and_host.py builds it; it needs a GPU to run).
#include <cstdio>
#include <vector>

#include "vio_host.hpp"

int main() {
    vio_config cfg;
    vio_config_default(&cfg);
    try {
        vio::FeatureTracker featuretracker(cfg);
        vio::VINS vins(cfg);
        vins.setIMUModel();
        vins.setExtrinsic();
        std::vector<uint8_t> frame((size_t)cfg.rows * cfg.cols, 128);
        vio::Mat img{frame.data(), cfg.rows, cfg.cols}, result{nullptr, 0, 0};
        vio::Vector3d P{0, 0, 0};
        vio::Matrix3d R{};
        for (int k = 0; k < 6; k++) {
            std::vector<vio::Point2f> good_pts;
            std::vector<double> track_len;
            const bool publish = featuretracker.img_cnt == 0;
            featuretracker.readImage(img, result, k, good_pts, track_len, k / 30.0, P, R, false);
            if (publish) {
                for (int j = 0; j < 20; j++) vins.processIMU(0.005, vio::Vector3d{0, 0, 9.805}, vio::Vector3d{0, 0, 0});
                vins.processImage(featuretracker.image_msg, k / 30.0, 0);
            }
        }
        std::printf("frame_count %d features %zu\n", vins.frame_count, featuretracker.ids.size());
    } catch (const std::exception &e) {
        std::printf("no CUDA device / error: %s\n", e.what());
        return 2;
    }
    return 0;
}
