// vio_host.hpp -- host-side C++ mirror of the reference call surface over the C-ABI (include/vio_b200.h).
//
// The reference owns the hot path through two global objects, `FeatureTracker featuretracker; VINS vins;`
// (/root/reference/VINS_ios/ViewController.mm:107,109).  These two classes keep the member names, argument order and the public
// fields a caller reads (feature_tracker.hpp:52-90, VINS.hpp:51-172) so that the estimator loop of ViewController.mm:364-494 and
// :701-882 compiles against them with the OpenCV / Eigen types swapped for the plain structs below.  One object = one stream
// (batch 1); batched use goes through the C-ABI directly.  Everything numerical happens in libvio_b200.so on the GPU.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vio_b200.h"

namespace vio {

struct Point2f { float x, y; };
struct Vector3d { double x, y, z; double &operator()(int i) { return (&x)[i]; } };
struct Matrix3d { double m[9]; };                          // row-major
struct Mat { const uint8_t *data; int rows, cols; bool empty() const { return data == nullptr; } };   // CV_8UC1, stride == cols

inline void check(int rc, const char *what) {
    if (rc != VIO_OK) throw std::runtime_error(std::string(what) + " failed with code " + std::to_string(rc));
}

// vins_pnp.hpp:27-39
struct VINS_RESULT { double header; Vector3d Ba, Bg, P; Matrix3d R; Vector3d V; };
struct Vector2d { double x, y; };
struct IMG_MSG_LOCAL { int id; Vector2d observation; Vector3d position; int track_num; };
struct IMU_MSG_LOCAL { double header; Vector3d acc, gyr; };            // feature_tracker.hpp:30-35

// vins_pnp.hpp:40-91 -- the motion-only tracker behind FeatureTracker::solveVinsPnP (7-frame window, landmarks fixed)
class vinsPnP {
  public:
    static constexpr int PNP_SIZE = 6;                                   // global_param.hpp:29
    explicit vinsPnP(const vio_config &cfg) : cfg_(cfg), frame_count(0) {
        cfg_.batch = 1;
        check(vio_pnp_create(&cfg_, &h_), "vio_pnp_create");
    }
    ~vinsPnP() { vio_pnp_destroy(h_); }
    vinsPnP(const vinsPnP &) = delete;
    void setIMUModel() {}                                                // PerspectiveFactor::sqrt_info = fx / 1.5 comes from vio_config
    void setExtrinsic() {}
    void setInit(const VINS_RESULT &r) {                                 // vins_pnp.cpp:63-83
        check(vio_pnp_set_init(h_, &r.header, &r.P.x, r.R.m, &r.V.x, &r.Ba.x, &r.Bg.x), "vio_pnp_set_init");
    }
    void processIMU(double dt, const Vector3d &acc, const Vector3d &gyr) {     // vins_pnp.cpp:197-233
        check(vio_pnp_process_imu(h_, 1, &dt, &acc.x, &gyr.x), "vio_pnp_process_imu");
    }
    void processImage(std::vector<IMG_MSG_LOCAL> &feature_msg, double header, bool use_pnp) {      // vins_pnp.cpp:236-256
        const int cap = cfg_.max_cnt;
        if ((int)feature_msg.size() > cap) throw std::length_error("vinsPnP::processImage: more than max_cnt features");
        std::vector<int32_t> ids(cap, 0), tn(cap, 0);
        std::vector<double> obs(2 * cap, 0.0), pos(3 * cap, 0.0);
        int32_t n = 0;
        for (auto &f : feature_msg) {
            ids[n] = f.id; tn[n] = f.track_num; obs[2 * n] = f.observation.x; obs[2 * n + 1] = f.observation.y;
            pos[3 * n] = f.position.x; pos[3 * n + 1] = f.position.y; pos[3 * n + 2] = f.position.z; n++;
        }
        check(vio_pnp_process_image(h_, &n, ids.data(), obs.data(), pos.data(), tn.data(), &header, use_pnp ? 1 : 0), "vio_pnp_process_image");
        double P[21], R[63], V[21], H[7];
        int32_t fs[7], info[3];
        check(vio_pnp_get_state(h_, 0, P, R, V, H, fs, info, nullptr), "vio_pnp_get_state");
        for (int i = 0; i <= PNP_SIZE; i++) {
            Ps[i] = Vector3d{P[3 * i], P[3 * i + 1], P[3 * i + 2]}; Vs[i] = Vector3d{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
            for (int k = 0; k < 9; k++) Rs[i].m[k] = R[9 * i + k];
            Headers[i] = H[i]; find_solved[i] = fs[i] != 0;
        }
        frame_count = info[0];
    }
    int frame_count;
    Vector3d Ps[PNP_SIZE + 1], Vs[PNP_SIZE + 1];
    Matrix3d Rs[PNP_SIZE + 1];
    double Headers[PNP_SIZE + 1];
    bool find_solved[PNP_SIZE + 1];

  private:
    vio_config cfg_;
    vio_pnp *h_ = nullptr;
};

// feature_tracker.hpp:52-90
class FeatureTracker {
  public:
    explicit FeatureTracker(const vio_config &cfg) : cfg_(cfg), img_cnt(0), update_finished(false), use_pnp(false), vins_pnp(cfg), current_time(-1.0) {
        cfg_.batch = 1;
        check(vio_frontend_create(&cfg_, &h_), "vio_frontend_create");
        ids.reserve(cfg_.max_cnt);
    }
    ~FeatureTracker() { vio_frontend_destroy(h_); }
    // ViewController.mm:438-441 equalises every camera frame (cv::createCLAHE(); setClipLimit(3); apply()) before readImage: with this
    // switched on the caller hands the raw gray frame to readImage and the equalisation runs on the device
    void setClahe(bool enable, double clip_limit = 3.0, int tiles_x = 8, int tiles_y = 8) {
        check(vio_frontend_set_clahe(h_, enable ? 1 : 0, clip_limit, tiles_x, tiles_y), "vio_frontend_set_clahe");
    }
    FeatureTracker(const FeatureTracker &) = delete;

    // void readImage(const cv::Mat &_img, cv::Mat &result, int _frame_cnt, vector<Point2f> &good_pts, vector<double> &track_len,
    //                double header, Vector3d &P, Matrix3d &R, bool vins_normal)               feature_tracker.hpp:59
    // P / R are outputs of the motion-only PnP tracker (solveVinsPnP below; use_pnp defaults to false, ViewController.mm:144):
    // written when vins_normal is true, as in feature_tracker.cpp:203-208.
    void readImage(const Mat &_img, Mat &result, int _frame_cnt, std::vector<Point2f> &good_pts, std::vector<double> &track_len, double header,
                   Vector3d &P, Matrix3d &R, bool vins_normal) {
        (void)_frame_cnt;
        if (_img.rows != cfg_.rows || _img.cols != cfg_.cols) throw std::invalid_argument("readImage: image size differs from vio_config");
        result = _img;
        int published = 0;
        check(vio_frontend_read_images(h_, _img.data, &published), "vio_frontend_read_images");
        const int cap = cfg_.max_cnt;
        int n = 0;
        std::vector<float> g(4 * cap);                 // good_pts: up to 2 * max_cnt points (tracked before setMask + new corners)
        std::vector<double> tl(2 * cap);
        check(vio_frontend_get_ui(h_, 0, &n, g.data(), tl.data()), "vio_frontend_get_ui");
        for (int i = 0; i < n; i++) { good_pts.push_back(Point2f{g[2 * i], g[2 * i + 1]}); track_len.push_back(tl[i]); }
        std::vector<int32_t> id(cap), cnt(cap);
        std::vector<float> pts(2 * cap);
        std::vector<double> xyz(3 * cap);
        check(vio_frontend_get_stream(h_, 0, &n, id.data(), pts.data(), cnt.data(), xyz.data()), "vio_frontend_get_stream");
        ids.assign(id.begin(), id.begin() + n);
        track_cnt.assign(cnt.begin(), cnt.begin() + n);
        cur_pts.resize(n);
        for (int i = 0; i < n; i++) cur_pts[i] = Point2f{pts[2 * i], pts[2 * i + 1]};
        if (published) {                                                    // feature_tracker.cpp:287-307
            image_msg.clear();
            for (int i = 0; i < n; i++) image_msg[ids[i]] = Vector3d{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        }
        solveVinsPnP(header, P, R, vins_normal);                            // feature_tracker.cpp:203-208 (after the tracking step)
        update_finished = true;
        img_cnt = (img_cnt + 1) % cfg_.freq;                                // the caller does this at ViewController.mm:494
    }

    // bool solveVinsPnP(double header, Vector3d &P, Matrix3d &R, bool vins_normal)                 feature_tracker.cpp:107-160
    bool solveVinsPnP(double header, Vector3d &P, Matrix3d &R, bool vins_normal) {
        if (!vins_normal) return false;
        // The reference matches solved_features against ids with a merge that assumes ascending ids and runs between the first
        // RANSAC rejection and setMask (feature_tracker.cpp:119-132,207).  The C-ABI exposes the tracker state after the whole
        // frame, so the match is done by id on that state: on a detecting frame the points setMask dropped are not offered.
        std::vector<IMG_MSG_LOCAL> feature_msg;
        std::map<int, size_t> where;
        for (size_t i = 0; i < ids.size(); i++) where[ids[i]] = i;
        for (auto &it : solved_features) {
            auto f = where.find(it.id);
            if (f == where.end()) continue;
            IMG_MSG_LOCAL tmp = it;
            tmp.observation = Vector2d{(cur_pts[f->second].x - cfg_.cx) / cfg_.fx, (cur_pts[f->second].y - cfg_.cy) / cfg_.fy};
            feature_msg.push_back(tmp);
        }
        std::sort(feature_msg.begin(), feature_msg.end(), [](const IMG_MSG_LOCAL &a, const IMG_MSG_LOCAL &b) { return a.id < b.id; });
        vins_pnp.setInit(solved_vins);
        for (auto &it : imu_msgs) {
            if (current_time < 0) current_time = it.header;
            const double dt = it.header - current_time;
            current_time = it.header;
            vins_pnp.processIMU(dt, it.acc, it.gyr);
        }
        vins_pnp.processImage(feature_msg, header, use_pnp);
        P = vins_pnp.Ps[vinsPnP::PNP_SIZE - 1];
        R = vins_pnp.Rs[vinsPnP::PNP_SIZE - 1];
        return true;
    }

    std::vector<Point2f> cur_pts;
    std::vector<int> ids, track_cnt;
    int img_cnt;
    std::map<int, Vector3d> image_msg;
    bool update_finished;
    bool use_pnp;
    std::vector<IMG_MSG_LOCAL> solved_features;          // feature_tracker.hpp:77-79: filled by the caller from the estimator (ViewController.mm:445-447)
    VINS_RESULT solved_vins{};
    std::vector<IMU_MSG_LOCAL> imu_msgs;
    vinsPnP vins_pnp;
    double current_time;
    vio_frontend *handle() { return h_; }

  private:
    vio_config cfg_;
    vio_frontend *h_ = nullptr;
};

// VINS.hpp:51-172
// VINS.hpp:28-45 (the fields the window solve reads and writes)
struct RetriveData {
    double header = -1;
    Vector3d P_old{0, 0, 0};
    std::array<double, 4> Q_old{0, 0, 0, 1};               // x,y,z,w
    std::vector<Point2f> measurements;                     // normalised image coordinates in the old keyframe
    std::vector<int> features_ids;                         // ascending
    Vector3d relative_t{0, 0, 0};
    std::array<double, 4> relative_q{0, 0, 0, 1};
    double relative_yaw = 0;
};

class VINS {
  public:
    enum SolverFlag { INITIAL, NON_LINEAR };
    enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };

    explicit VINS(const vio_config &cfg) : cfg_(cfg), frame_count(0), solver_flag(INITIAL), marginalization_flag(MARGIN_OLD), failure_occur(false),
                                           final_cost(0), feature_num(0) {
        cfg_.batch = 1;
        const int n = cfg_.window_size + 1;
        Ps.resize(n); Vs.resize(n); Bas.resize(n); Bgs.resize(n); Qs.resize(n); Rs.resize(n); Headers.resize(n);
        check(vio_backend_create(&cfg_, &h_), "vio_backend_create");
    }
    ~VINS() { vio_backend_destroy(h_); }
    VINS(const VINS &) = delete;

    void setIMUModel() {}                      // ProjectionFactor::sqrt_info = fx/1.5 is derived from vio_config (VINS.cpp:29-32)
    void setExtrinsic() {}                     // tic/ric come from vio_config (VINS.cpp:82-88)
    void clearState() { check(vio_backend_clear(h_), "vio_backend_clear"); refresh(); }

    // void processIMU(double dt, const Vector3d &linear_acceleration, const Vector3d &angular_velocity)        VINS.hpp:164
    void processIMU(double dt, const Vector3d &acc, const Vector3d &gyr) {
        check(vio_backend_process_imu(h_, 1, &dt, &acc.x, &gyr.x), "vio_backend_process_imu");
    }

    // caller-supplied result of solveInitial() (VINS.cpp:833-1102, out of scope): P,Q(xyzw),V for frames 0..WINDOW_SIZE
    void setInitialWindow(const double *P, const double *Q, const double *V, const double Ba[3], const double Bg[3]) {
        check(vio_backend_set_init_window(h_, P, Q, V, Ba, Bg), "vio_backend_set_init_window");
    }

    // vector<pair<Vector3d, Vector3d>> FeatureManager::getCorresponding(int frame_count_l, int frame_count_r)          feature_manager.cpp:157-176
    // (f_manager lives on the device: this reads its table back) -- the correspondences relativePose / the SfM set-up work from
    std::vector<std::pair<Vector3d, Vector3d>> getCorresponding(int frame_count_l, int frame_count_r) {
        const int cap = cfg_.num_of_f, nfr = cfg_.window_size + 1;
        std::vector<int32_t> ids(cap), st(cap), no(cap), fl(cap);
        std::vector<double> dep(cap), obs((size_t)cap * nfr * 2);
        int n = 0, n2 = 0;
        check(vio_backend_get_features(h_, 0, cap, &n, ids.data(), st.data(), no.data(), dep.data(), fl.data()), "vio_backend_get_features");
        check(vio_backend_get_observations(h_, 0, cap, &n2, obs.data()), "vio_backend_get_observations");
        std::vector<std::pair<Vector3d, Vector3d>> corres;
        for (int i = 0; i < n && i < n2; i++)
            if (st[i] <= frame_count_l && st[i] + no[i] - 1 >= frame_count_r) {
                const double *a = &obs[((size_t)i * nfr + (frame_count_l - st[i])) * 2], *b = &obs[((size_t)i * nfr + (frame_count_r - st[i])) * 2];
                Vector3d pa, pb;
                pa.x = a[0]; pa.y = a[1]; pa.z = 1.0; pb.x = b[0]; pb.y = b[1]; pb.z = 1.0;
                corres.push_back(std::make_pair(pa, pb));
            }
        return corres;
    }

    // The other way in: hand over what solveInitial() holds after the global SfM (VINS.cpp:889-905) -- ImageFrame::R (row-major 3x3) and
    // ImageFrame::T of the window's WINDOW_SIZE + 1 frames -- and let the device run visualInitialAlign (VINS.cpp:1022-1102:
    // VisualIMUAlignment, scale, gravity frame, velocities, depths) inside the processImage call that fills the window.
    void setInitialSfm(const double *R, const double *T) { check(vio_backend_set_init_sfm(h_, R, T), "vio_backend_set_init_sfm"); }
    // The general form: poses of EVERY frame of all_image_frame (keyframes and the frames MARGIN_SECOND_NEW dropped from the window), in
    // time order, the frame about to be processed last; initialFrames() lists the headers the device holds so far.
    void setInitialSfmFrames(int n_frames, const double *R, const double *T) {
        int32_t n = n_frames;
        check(vio_backend_set_init_sfm_frames(h_, &n, n_frames, R, T), "vio_backend_set_init_sfm_frames");
    }
    std::vector<double> initialFrames() {
        std::vector<double> h(3 * (cfg_.window_size + 1));
        int32_t n = 0;
        check(vio_backend_get_init_frames(h_, 0, (int)h.size(), &n, h.data()), "vio_backend_get_init_frames");
        h.resize(n);
        return h;
    }
    // bool result of the last VisualIMUAlignment (-1: none yet), vins.g after it and the metric scale of the SfM
    int initialAlignment(Vector3d *g = nullptr, double *scale = nullptr) {
        int32_t ok = -1; double gg[3] = {0, 0, 0}, sc = 0;
        check(vio_backend_get_init_result(h_, 0, &ok, gg, &sc), "vio_backend_get_init_result");
        if (g) { g->x = gg[0]; g->y = gg[1]; g->z = gg[2]; }
        if (scale) *scale = sc;
        return ok;
    }

    // void processImage(map<int, Vector3d> &image_msg, double header, int buf_num)                             VINS.hpp:163
    // buf_num only scaled the wall-time cap of ceres::Solve (VINS.cpp:648-653); the cap is removed (it made the reference
    // timing-dependent), so buf_num is accepted and ignored.  solve_ceres() runs inside, on the device.
    void processImage(std::map<int, Vector3d> &image_msg, double header, int buf_num) {
        (void)buf_num;
        const int cap = cfg_.max_cnt;
        std::vector<int32_t> ids(cap, 0);
        std::vector<double> xyz(3 * cap, 1.0);
        int32_t n = 0;
        for (auto &kv : image_msg) {
            if (n >= cap) throw std::length_error("processImage: more than max_cnt features");
            ids[n] = kv.first; xyz[3 * n] = kv.second.x; xyz[3 * n + 1] = kv.second.y; xyz[3 * n + 2] = kv.second.z; n++;
        }
        push_loop_match();
        check(vio_backend_process_image(h_, &n, ids.data(), xyz.data(), &header), "vio_backend_process_image");
        refresh();
    }
    // VINS.hpp:153.  processImage() already runs the solve on the device; called on its own it re-solves the current window
    // (vio_backend_solve).  buf_num only shortens the reference's wall-time cap, which is not reproduced.
    void solve_ceres(int buf_num) { (void)buf_num; check(vio_backend_solve(h_), "vio_backend_solve"); refresh(); }

    int frame_count;
    SolverFlag solver_flag;
    MarginalizationFlag marginalization_flag;
    std::vector<Vector3d> Ps, Vs, Bas, Bgs;
    std::vector<std::array<double, 4>> Qs;                 // Rs as quaternions (x,y,z,w)
    std::vector<Matrix3d> Rs;                              // VINS.hpp:74 (row-major), refreshed together with Qs
    std::vector<double> Headers;
    bool failure_occur;
    double final_cost;
    int feature_num;
    // loop closure (needs vio_config::loop_closure = 1): the caller (loop-closure thread, ViewController.mm:964) writes retrive_pose_data;
    // the next solves add the loop factors (VINS.cpp:571-637) and publish relative_t / relative_q / relative_yaw in it, and the drift
    // correction r_drift = ypr2R(drift_yaw, 0, 0), t_drift (VINS.hpp:117-118)
    RetriveData retrive_pose_data;
    double drift_yaw = 0;
    Vector3d t_drift{0, 0, 0};
    vio_backend *handle() { return h_; }

  private:
    void push_loop_match() {
        if (!cfg_.loop_closure || retrive_pose_data.header == pushed_header_) return;
        const RetriveData &r = retrive_pose_data;
        const int cap = cfg_.max_cnt;
        const int32_t n = (int32_t)std::min<size_t>(r.features_ids.size(), (size_t)cap);
        std::vector<int32_t> ids(cap, 0);
        std::vector<double> xy(2 * cap, 0.0);
        for (int i = 0; i < n; i++) { ids[i] = r.features_ids[i]; xy[2 * i] = r.measurements[i].x; xy[2 * i + 1] = r.measurements[i].y; }
        const double pose[7] = {r.P_old.x, r.P_old.y, r.P_old.z, r.Q_old[0], r.Q_old[1], r.Q_old[2], r.Q_old[3]};
        check(vio_backend_set_loop_match(h_, &n, &r.header, ids.data(), xy.data(), pose), "vio_backend_set_loop_match");
        pushed_header_ = r.header;
    }
    void refresh() {
        if (cfg_.loop_closure) {
            double lo[12]; int32_t nf = 0;
            check(vio_backend_get_loop_result(h_, 0, lo, &nf), "vio_backend_get_loop_result");
            if (nf > 0) {
                retrive_pose_data.relative_t = Vector3d{lo[0], lo[1], lo[2]};
                retrive_pose_data.relative_q = {lo[3], lo[4], lo[5], lo[6]};
                retrive_pose_data.relative_yaw = lo[7];
                drift_yaw = lo[8]; t_drift = Vector3d{lo[9], lo[10], lo[11]};
            }
        }
        const int n = cfg_.window_size + 1;
        std::vector<double> P(3 * n), Q(4 * n), V(3 * n), Ba(3 * n), Bg(3 * n);
        check(vio_backend_get_state(h_, 0, P.data(), Q.data(), V.data(), Ba.data(), Bg.data(), Headers.data()), "vio_backend_get_state");
        for (int i = 0; i < n; i++) {
            Ps[i] = Vector3d{P[3 * i], P[3 * i + 1], P[3 * i + 2]}; Vs[i] = Vector3d{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
            Bas[i] = Vector3d{Ba[3 * i], Ba[3 * i + 1], Ba[3 * i + 2]}; Bgs[i] = Vector3d{Bg[3 * i], Bg[3 * i + 1], Bg[3 * i + 2]};
            Qs[i] = {Q[4 * i], Q[4 * i + 1], Q[4 * i + 2], Q[4 * i + 3]};
            const double x = Q[4 * i], y = Q[4 * i + 1], z = Q[4 * i + 2], w = Q[4 * i + 3];
            Rs[i] = Matrix3d{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                              2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
        }
        int32_t info[8]; double dinfo[4];
        check(vio_backend_get_info(h_, 0, info, dinfo), "vio_backend_get_info");
        solver_flag = info[0] ? NON_LINEAR : INITIAL;
        marginalization_flag = info[1] ? MARGIN_SECOND_NEW : MARGIN_OLD;
        frame_count = info[2]; failure_occur = info[3] != 0; feature_num = info[4]; final_cost = dinfo[1];
    }
    vio_config cfg_;
    vio_backend *h_ = nullptr;
    double pushed_header_ = -1;
};

}  // namespace vio
