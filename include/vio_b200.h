/*
 * vio_b200.h -- C-ABI of the B200-native VIO hot path (libvio_b200.so).
 *
 * Drop-in boundary for the two reference classes that own the hot path
 * (SURVEY.md section 8(b)); the reference has no FFI layer of its own, so every
 * entry point cites the C++ member it replaces (paths relative to
 * /root/reference/VINS_ios/):
 *
 *   FeatureTracker::readImage            feature_tracker.hpp:59,  feature_tracker.cpp:162-310
 *   FeatureTracker::{setMask,rejectWithF,addPoints,updateID}
 *                                        feature_tracker.cpp:36-103,311-321
 *   VINS::processIMU                     VINS.hpp:164, VINS.cpp:333-375
 *   VINS::processImage                   VINS.hpp:163, VINS.cpp:377-478
 *   VINS::solve_ceres                    VINS.hpp:153, VINS.cpp:480-831
 *   VINS::{clearState,setIMUModel,setExtrinsic}   VINS.hpp:157-159
 *   setGlobalParam                       global_param.hpp:83, global_param.cpp:24-137
 *
 * Everything is BATCHED: a handle owns `batch` independent streams that advance in
 * lock-step (one FeatureTracker + one VINS object per stream in reference terms);
 * batch == 1 is the single-stream drop-in.  Plain pointers and sizes only; caller
 * allocates every output; no ownership crosses the ABI; int return codes
 * (0 = VIO_OK).  "dev" variants take/return DEVICE pointers and never touch the
 * host; the plain variants take HOST pointers and do their own H2D/D2H copies.
 *
 * The library contains no CPU fallback: every entry point that computes needs a
 * CUDA device and returns VIO_ERR_CUDA otherwise.
 */
#ifndef VIO_B200_H
#define VIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIO_OK 0
#define VIO_ERR_ARG 1
#define VIO_ERR_CUDA 2
#define VIO_ERR_STATE 3
#define VIO_ERR_CAPACITY 4

/* Compile-time macros of the reference turned into run-time parameters
 * (global_param.hpp:23-53, feature_tracker.hpp:24-29, feature_manager.hpp:22-25). */
typedef struct vio_config {
    int32_t rows, cols;          /* ROW 640, COL 480           feature_tracker.hpp:26-27 */
    double fx, fy, cx, cy;       /* FOCUS_LENGTH_X/Y, PX, PY   global_param.cpp:29-32    */
    double tic[3];               /* TIC_X/Y/Z                  global_param.cpp:37-39    */
    double ric[9];               /* row-major; ypr2R(RIC_y,p,r) VINS.cpp:55               */
    int32_t max_cnt;             /* MAX_CNT 70                 feature_tracker.hpp:24    */
    int32_t min_dist;            /* MIN_DIST 30                feature_tracker.hpp:25    */
    double f_threshold;          /* F_THRESHOLD 1.0            feature_tracker.hpp:28    */
    int32_t freq;                /* FREQ 3                     global_param.cpp:35       */
    int32_t window_size;         /* WINDOW_SIZE 10             global_param.hpp:28       */
    int32_t num_of_f;            /* NUM_OF_F 1000              global_param.hpp:37       */
    double acc_n, acc_w, gyr_n, gyr_w;  /* global_param.hpp:43-46 */
    double gravity;              /* GRAVITY 9.805              global_param.hpp:42       */
    int32_t max_iters;           /* options.max_num_iterations VINS.cpp:646              */
    double min_parallax;         /* MIN_PARALLAX 10/549        feature_manager.hpp:24    */
    double init_depth;           /* INIT_DEPTH 5.0             feature_manager.hpp:25    */
    int32_t max_imu_per_frame;   /* capacity of dt_buf[] etc.  VINS.hpp:103-105          */
    int32_t batch;               /* number of independent streams in the handle          */
    int32_t device;              /* CUDA device ordinal                                   */
    /* How the back end computes (all 0 by default; results agree within the parity tolerance whichever is chosen): */
    int32_t marg_mode;           /* 0: new prior formed directly in information form (H = A_r, b, c0).  1: the reference's route,
                                    eigendecomposition of A_r -> linearized_jacobians / residuals (marginalization_factor.cpp:270-294) */
    int32_t marg_eig;            /* eigen-solver used by marg_mode 1 and by the pseudo-inverse of Amm: 0 Householder tridiagonalisation
                                    + implicit QL, 1 parallel Jacobi */
    int32_t marg_amm_eig;        /* 1: Amm^+ always through its eigendecomposition with the 1e-8 cut (marginalization_factor.cpp:270-276);
                                    0: structured inverse guarded by an eigenvalue bound, falling back to the eigendecomposition */
    int32_t solve_path;          /* reduced-system solve: 0 auto (FP64 tensor-pipe tiles when the window fits one SM, else global memory),
                                    1 packed Cholesky in shared memory without DMMA tiles */
    int32_t be_threads;          /* threads per stream in the solve / marginalisation kernels: 0 = 512, or 256 */
    int32_t loop_closure;        /* LOOP_CLOSURE (VINS.cpp:11): 1 reserves the loop-closure pose in the window solve; the factors are added when a
                                    match has been supplied with vio_backend_set_loop_match (VINS.cpp:571-637).  0 (default): never. */
} vio_config;

/* Fills `cfg` with the reference's iPhone7P defaults (global_param.cpp:26-39) and the
 * BASELINE.json bench values (max_cnt 150, window 10, freq 3). */
void vio_config_default(vio_config *cfg);

/* ------------------------------------------------------------------ front end */
typedef struct vio_frontend vio_frontend;

int vio_frontend_create(const vio_config *cfg, vio_frontend **out);
void vio_frontend_destroy(vio_frontend *fe);

/* FeatureTracker::readImage for every stream of the batch.
 *   images : batch contiguous rows*cols u8 images (row-major, stride = cols)
 * Detection / ID assignment / image_msg publication happen when the internal img_cnt == 0,
 * then img_cnt = (img_cnt+1) % freq (ViewController.mm:494).  Returns VIO_OK; *published is
 * set to 1 when this call produced a new image_msg. */
int vio_frontend_read_images(vio_frontend *fe, const uint8_t *images_host, int *published);
int vio_frontend_read_images_dev(vio_frontend *fe, const uint8_t *images_dev, int *published);
/* Device buffer that the NEXT read_images_dev call will consume if the caller passes it back
 * (zero-copy producer path): batch * rows * cols bytes. */
uint8_t *vio_frontend_next_image_buffer(vio_frontend *fe);

/* Tracker state after the last call, HOST outputs, per stream `s`:
 *   n_out            number of tracked points
 *   ids[n]           FeatureTracker::ids            (-1 between detect frames for fresh points never occurs: ids are assigned on detect frames)
 *   pts_xy[2n]       FeatureTracker::cur_pts  (pixels)
 *   track_cnt[n]     FeatureTracker::track_cnt
 *   norm_xyz[3n]     image_msg values ((u-PX)/fx,(v-PY)/fy,1), only meaningful after a publishing call
 * Any output pointer may be NULL.  Arrays must hold max_cnt entries. */
int vio_frontend_get_stream(vio_frontend *fe, int s, int *n_out, int32_t *ids, float *pts_xy,
                            int32_t *track_cnt, double *norm_xyz);
/* readImage's UI outputs good_pts / track_len (feature_tracker.cpp:209-226,237-250,276-280).  On a detecting frame the
 * reference lists every point tracked BEFORE setMask plus the new corners, so the arrays must hold 2 * max_cnt entries. */
int vio_frontend_get_ui(vio_frontend *fe, int s, int *n_out, float *good_pts_xy, double *track_len);
/* Stage counters of the last call for stream s: [0] lk_in [1] lk_ok [2] f1_ok [3] f2_ok [4] kept [5] new [6] n_candidates [7] ransac iters */
int vio_frontend_get_stats(vio_frontend *fe, int s, int32_t stats[8]);
/* Device-resident SoA views (valid until the next call): counts[batch], ids[batch*max_cnt],
 * norm_xyz[batch*max_cnt*3] -- what vio_backend_process_image_dev consumes. */
int vio_frontend_image_msg_dev(vio_frontend *fe, const int32_t **counts, const int32_t **ids, const double **norm_xyz);
/* Optional pre-processing of every frame handed to vio_frontend_read_images[_dev]: cv::createCLAHE() + setClipLimit(clip_limit) +
 * apply(), the step the reference runs on the camera frame right before FeatureTracker::readImage (ViewController.mm:438-441:
 * clip_limit 3.0, default 8x8 tiles).  Bit-identical to OpenCV 4.13's CLAHE.  rows / cols must be divisible by tiles_y / tiles_x
 * (OpenCV pads otherwise; VIO_ERR_ARG here).  enable = 0 switches it off (default). */
int vio_frontend_set_clahe(vio_frontend *fe, int enable, double clip_limit, int tiles_x, int tiles_y);
/* Number of CUDA kernels launched by this handle since creation. */
int64_t vio_frontend_launch_count(const vio_frontend *fe);
/* Primitive entry points used by the parity tests (single image, host in/out).  */
int vio_prim_pyramid(const vio_config *cfg, const uint8_t *img, uint8_t *l1, uint8_t *l2, uint8_t *l3);
int vio_prim_clahe(const vio_config *cfg, const uint8_t *img, double clip_limit, int tiles_x, int tiles_y, uint8_t *out);
int vio_prim_min_eig_candidates(const vio_config *cfg, const uint8_t *img, const float *kept_xy, int n_kept,
                                int max_corners, float *corners_xy, int *n_corners, float *max_val);
int vio_prim_lk(const vio_config *cfg, const uint8_t *prev, const uint8_t *next, const float *pts_xy, int n,
                float *next_xy, uint8_t *status);
int vio_prim_ransac_f(const vio_config *cfg, const float *p1_xy, const float *p2_xy, int n, uint8_t *mask, int *iters);

/* ------------------------------------------------------------------- back end */
typedef struct vio_backend vio_backend;

int vio_backend_create(const vio_config *cfg, vio_backend **out);
void vio_backend_destroy(vio_backend *be);
/* VINS::clearState for every stream. */
int vio_backend_clear(vio_backend *be);

/* VINS::processIMU, n_samples consecutive samples for every stream:
 *   dt[n_samples*batch], acc[n_samples*batch*3], gyr[n_samples*batch*3]   (sample-major) */
int vio_backend_process_imu(vio_backend *be, int n_samples, const double *dt, const double *acc, const double *gyr);
int vio_backend_process_imu_dev(vio_backend *be, int n_samples, const double *dt, const double *acc, const double *gyr);

/* External initialisation: the output of VINS::solveInitial/visualInitialAlign (VINS.cpp:833-1102,
 * out of scope) supplied by the caller for frames 0..window_size:
 * P[b][W+1][3], Q[b][W+1][4] (x,y,z,w), V[b][W+1][3], Ba[b][3], Bg[b][3].  Consumed by the
 * process_image call that fills the window (frame_count == window_size, solver_flag INITIAL). */
int vio_backend_set_init_window(vio_backend *be, const double *P, const double *Q, const double *V,
                                const double *Ba, const double *Bg);

/* VINS::processImage for every stream:  counts[batch], ids[batch*max_cnt], norm_xyz[batch*max_cnt*3],
 * headers[batch].  Runs addFeatureCheckParallax, triangulate, solve_ceres (dogleg, <= max_iters),
 * marginalisation, failureDetection and slideWindow on the device. */
int vio_backend_process_image(vio_backend *be, const int32_t *counts, const int32_t *ids, const double *norm_xyz,
                              const double *headers);
int vio_backend_process_image_dev(vio_backend *be, const int32_t *counts, const int32_t *ids, const double *norm_xyz,
                                  const double *headers_host);

/* VINS::solve_ceres(buf_num) on its own (VINS.hpp:153, VINS.cpp:480-831) for every stream that is NON_LINEAR with a full window: problem
 * build, <= max_iters dogleg iterations, new2old and the marginalisation chosen by the current marginalization_flag, on the window as it
 * stands -- none of the processImage steps around it.  Read the result with vio_backend_get_post_solve / _get_state / _get_prior.
 * (buf_num only shortens the reference's wall-time cap, which is not reproduced.) */
int vio_backend_solve(vio_backend *be);

/* Loop-closure factors in the window solve (VINS.cpp:571-637, 664-680, 174-195; needs vio_config::loop_closure = 1).
 * vio_backend_set_loop_match = retrive_pose_data (VINS.hpp:28-45): per stream the header of the window keyframe that was matched, the ids
 * (ascending) and normalised image measurements [batch][max_cnt][2] of the shared features in the OLD keyframe, and the old keyframe's pose
 * pose_old[batch][7] = P_old, Q_old (x,y,z,w).  counts[b] = 0: no match for stream b.  A match stays in force until replaced (front_pose).
 * While the matched header is inside the window, every solve adds ProjectionFactor(first observation, old measurement) between the
 * landmark's anchor pose and a free loop pose initialised from the matched frame.
 * vio_backend_get_loop_result: out = relative_t[3], relative_q[4] (xyzw), relative_yaw, drift yaw (r_drift = ypr2R(yaw,0,0)), t_drift[3]
 * of the last solve; *n_factors = 0 when that solve had no loop constraint. */
int vio_backend_set_loop_match(vio_backend *be, const int32_t *counts, const double *headers, const int32_t *ids, const double *xy,
                               const double *pose_old);
int vio_backend_get_loop_result(vio_backend *be, int s, double out[12], int32_t *n_factors);

/* processImage fed from a front end's device-resident image_msg; event-ordered hand-over when the two handles use different
 * CUDA streams (front end of the next frames overlaps the solve). */
int vio_backend_process_image_from_frontend(vio_backend *be, vio_frontend *fe, const double *headers_host);

/* Window state of stream s (VINS::Ps/Rs/Vs/Bas/Bgs/Headers, VINS.hpp:73-77,107):
 * P[(W+1)*3], Q[(W+1)*4] x,y,z,w, V[(W+1)*3], Ba[(W+1)*3], Bg[(W+1)*3], headers[W+1]. NULL = skip */
int vio_backend_get_state(vio_backend *be, int s, double *P, double *Q, double *V, double *Ba, double *Bg, double *headers);
/* Whole-batch device pointer to the packed state [batch][W+1][16] = P3,Q4(xyzw),V3,Ba3,Bg3 (for NCCL gathers). */
int vio_backend_state_dev(vio_backend *be, const double **state, int64_t *n_doubles);
/* Per-stream error flag.  A capacity overflow inside a kernel (IMU buffer beyond max_imu_per_frame, feature table beyond
 * (window_size + 2) * max_cnt entries) cannot fail the call that enqueued the kernel, so it is latched per stream: the state getters
 * (vio_backend_get_state / _get_post_solve) return it, VIO_ERR_CAPACITY, until it is cleared here (clear != 0) or by vio_backend_clear(). */
int vio_backend_get_error(vio_backend *be, int s, int clear, int32_t *code);
/* info[0] solver_flag (0 INITIAL,1 NON_LINEAR) [1] marginalization_flag (0 OLD,1 SECOND_NEW) [2] frame_count
 * [3] failure_occur [4] feature count in solve [5] projection factors [6] iterations run [7] last_track_num
 * dinfo[0] initial cost [1] final cost [2] prior size n */
int vio_backend_get_info(vio_backend *be, int s, int32_t info[8], double dinfo[4]);
/* f_manager.feature of stream s: ids, start_frame, n_obs, estimated_depth, solve_flag; arrays sized `cap`. */
int vio_backend_get_features(vio_backend *be, int s, int cap, int *n_out, int32_t *ids, int32_t *start_frame,
                             int32_t *n_obs, double *depth, int32_t *solve_flag);
/* FeaturePerId::feature_per_frame[k].point (x, y; z = 1) of the same features in the same order: obs [cap][window_size + 1][2], entry
 * k < n_obs valid.  What FeatureManager::getCorresponding (feature_manager.cpp:157-176) and the SfM set-up of VINS::solveInitial
 * (VINS.cpp:857-886) read from f_manager -- for a caller that keeps relativePose / GlobalSFM on the host. */
int vio_backend_get_observations(vio_backend *be, int s, int cap, int *n_out, double *obs);
/* Prior in information form over the canonical local layout [pose0(6) sb0(9) ... poseW sbW ex(6)],
 * n = 15*(W+1)+6:  H[n*n] = J0^T J0, b[n] = J0^T r0, present[2*(W+1)+1] block mask.  For parity tests. */
int vio_backend_get_prior(vio_backend *be, int s, double *H, double *b, int32_t *present, double *c0);
/* Packed state of the whole batch [batch][W+1][16] (P3,Q4 xyzw,V3,Ba3,Bg3) copied to caller memory.  dst_is_device = 0: host,
 * synchronous; 1: device, stream-ordered (e.g. the send buffer of an NCCL gather); 2: PINNED host memory, stream-ordered -- the caller
 * reads it after vio_backend_sync() or an event recorded on the back-end stream (keeps a pipelined caller from stalling on the solve). */
int vio_backend_copy_state(vio_backend *be, double *dst, int dst_is_device);
/* State right after new2old() of the last solve, before marginalisation / slideWindow: [W+1][16] of stream s (parity tests). */
int vio_backend_get_post_solve(vio_backend *be, int s, double *out);
int64_t vio_backend_launch_count(const vio_backend *be);
int vio_backend_sync(vio_backend *be);
/* Run the handle's kernels on a caller-owned cudaStream_t (passed as void*), e.g. to chain front end -> back end without a host
 * synchronisation or to time with the caller's CUDA events. */
int vio_backend_use_stream(vio_backend *be, void *cuda_stream);
int vio_frontend_use_stream(vio_frontend *fe, void *cuda_stream);
int vio_frontend_sync(vio_frontend *fe);
void *vio_frontend_stream(vio_frontend *fe);
/* Per-kernel CUDA-event timing: returns "name:launches:total_ms;..." accumulated since the previous call and switches the
 * timer on/off for subsequent launches. */
int vio_frontend_profile(vio_frontend *fe, int enable, char *out, int cap);
/* Per-stream clock64 cycle counters of the solve (slots 0-7) and marginalisation (8-15) phases: out[batch*32]. */
int vio_backend_phase_cycles(vio_backend *be, long long *out, int reset);
int vio_backend_profile(vio_backend *be, int enable, char *out, int cap);

/* Factor-level primitives for parity tests (host in/out, one factor each). */
int vio_prim_preintegrate(const vio_config *cfg, int n, const double *dt, const double *acc, const double *gyr,
                          const double acc0[3], const double gyr0[3], const double ba[3], const double bg[3],
                          double *delta_pqv /*10: p3 q4(xyzw) v3*/, double *jacobian /*225*/, double *covariance /*225*/,
                          double *sum_dt);
int vio_prim_imu_factor(const vio_config *cfg, const double *delta_pqv, const double *jacobian, const double *covariance,
                        double sum_dt, const double lin_ba[3], const double lin_bg[3],
                        const double pose_i[7], const double sb_i[9], const double pose_j[7], const double sb_j[9],
                        double *residual /*15*/, double *J /*15x30 row-major, local: [pi6 sbi9 pj6 sbj9]*/);
/* Same factor with the 15x15 weighting matrix made explicit: sqrt_info_in (row-major, upper triangular; NULL = form it on the device
 * from `covariance` as imu_factor.h:72 does) and sqrt_info_out (NULL or 225 doubles: the matrix that was used).  The covariance has
 * condition ~1e8, so the parity tests check the weighting matrix on its own (entry-wise and through U^T U cov = I) and the rest of
 * IMUFactor::Evaluate (imu_factor.h:27-184) with the reference's matrix passed in. */
int vio_prim_imu_factor_sqi(const vio_config *cfg, const double *delta_pqv, const double *jacobian, const double *covariance,
                            double sum_dt, const double lin_ba[3], const double lin_bg[3],
                            const double pose_i[7], const double sb_i[9], const double pose_j[7], const double sb_j[9],
                            const double *sqrt_info_in, double *sqrt_info_out, double *residual /*15*/, double *J /*15x30*/);
int vio_prim_projection_factor(const vio_config *cfg, const double pts_i[3], const double pts_j[3],
                               const double pose_i[7], const double pose_j[7], double inv_dep,
                               double *residual /*2*/, double *J /*2x13 row-major: [pi6 pj6 lambda1]*/);

/* ------------------------------------------------------------------------------------------------------------------
 * Visual-inertial alignment of the initialisation (the linear-algebra half).
 * Replaces VisualIMUAlignment (initial_aligment.cpp:222-229) = solveGyroscopeBias (:10-46) + SolveScale (:135-220) +
 * RefineGravity (:64-133) over the map<double, ImageFrame> VINS::solveInitial hands it (VINS.cpp:889-905), for `batch` streams
 * at once, one CTA per stream; the IMU pre-integration of every frame interval (IntegrationBase ctor + push_back,
 * VINS.cpp:340-352) and its re-propagation with the corrected gyroscope bias (integration_base.h:46-61) run on the device.
 * Host in / out (an initialisation-time call, not on the per-frame path):
 *   n_frames   [batch]                         frames in all_image_frame (2 .. max_frames)
 *   R, T       [batch][max_frames][9 | 3]      ImageFrame::R (row-major) and ImageFrame::T of every frame, in map order
 *   imu_counts [batch][max_frames]             samples of the interval ENDING at frame k (0 .. max_imu); frame 0's are never read
 *   imu0       [batch][max_frames][6]          acc_0, gyr_0 each interval starts from (the previous sample)
 *   imu        [batch][max_frames][max_imu][7] dt, acc xyz, gyr xyz
 *   bg0        [batch][3]                      Bgs[*] before the call
 *   bgs        [batch][3]                      Bgs[*] after (bg0 + delta_bg)
 *   g          [batch][3]                      refined gravity in the frame R / T are expressed in
 *   x          [batch][3 max_frames + 4]       the reference's VectorXd x: body-frame velocity of every frame, then the 2 tangent-plane
 *                                              steps, then the scale s (already divided by 100) -- 3 n + 3 entries; when SolveScale
 *                                              itself rejects (|norm(g) - G_NORM| > G_THRESHOLD or s < 0) its 3 n + 4 entries instead
 *   ok         [batch]                         the bool VisualIMUAlignment returns
 * VIO_ERR_CAPACITY when a stream has more frames / samples than the maxima. */
int vio_visual_imu_align(const vio_config *cfg, int batch, int max_frames, int max_imu, const int32_t *n_frames,
                         const double *R, const double *T, const int32_t *imu_counts, const double *imu0, const double *imu,
                         const double *bg0, double *bgs, double *g, double *x, int32_t *ok);

/* The same alignment inside the estimator: VINS::visualInitialAlign (VINS.cpp:1022-1102) on the back end's own window.
 * vio_backend_set_init_sfm hands over what VINS::solveInitial has after the global SfM (VINS.cpp:889-905) -- ImageFrame::R [batch][W+1][9]
 * (row-major) and ImageFrame::T [batch][W+1][3] of the window's frames -- and the vio_backend_process_image call that fills the window
 * then runs, per stream and on the device: VisualIMUAlignment over the window's own IMU buffers; on success Ps / Rs from the SfM,
 * clearDepth + triangulate on the camera poses, repropagate with the new Bgs, metric scale, Vs, gravity-aligned frame, then the
 * first solve exactly as VINS.cpp:415-447; on failure Bgs keep the corrected bias and the window only slides (solveInitial == false).
 * The alignment runs over all_image_frame as the back end keeps it: every camera frame since the stream (re)started, keyframe or not
 * (VINS.cpp:392-398), each with its own IMU interval; frames older than the window are dropped as the reference does (VINS.cpp:1186-1193).
 * vio_backend_get_init_frames returns their headers (at most 3 (W + 1) - 1 are kept; beyond that the list is abandoned until the next
 * clearState and attempts are refused with VIO_ERR_STATE); vio_backend_set_init_sfm_frames takes R [batch][max_frames][9] / T
 * [batch][max_frames][3] for exactly those frames (n_frames[b] of them; a mismatch with the device's count refuses the attempt with
 * VIO_ERR_STATE); vio_backend_set_init_sfm is the common case where every frame is a keyframe (no MARGIN_SECOND_NEW slide since the start):
 * R / T of the window's W + 1 frames.  The SfM itself (relativePose / GlobalSFM / solvePnP) is NOT part of this library.
 * vio_backend_get_init_result: ok = 1 / 0 of the stream's last alignment (-1: none yet), g = vins.g after it, scale = the metric scale. */
int vio_backend_set_init_sfm(vio_backend *be, const double *R, const double *T);
int vio_backend_set_init_sfm_frames(vio_backend *be, const int32_t *n_frames, int max_frames, const double *R, const double *T);
int vio_backend_get_init_frames(vio_backend *be, int s, int cap, int32_t *n, double *headers);
int vio_backend_get_init_result(vio_backend *be, int s, int32_t *ok, double g[3], double *scale);

/* ------------------------------------------------------------------------------------------------------------------
 * Motion-only PnP tracker: FeatureTracker::solveVinsPnP (feature_tracker.cpp:107-160) and the
 * vinsPnP object it drives (vins_pnp.hpp:40-91, vins_pnp.cpp).  `batch` independent 7-frame windows advancing in lock-step.
 * Replaces: vinsPnP::setInit (:63-83), processIMU (:197-233), processImage (:236-256) incl. updateFeatures / solve_ceres /
 * slideWindow, PerspectiveFactor::Evaluate (perspective_factor.cpp:16-67), IMUFactorPnP::Evaluate (imu_factor_pnp.h).
 * The 10 ms wall-time cap of the reference's solve is not reproduced (machine dependent).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct vio_pnp vio_pnp;
int vio_pnp_create(const vio_config *cfg, vio_pnp **out);
void vio_pnp_destroy(vio_pnp *p);
/* solved_vins of every stream (ViewController.mm:734-739): header[batch], P[batch][3], R[batch][9] row-major, V, Ba, Bg [batch][3] */
int vio_pnp_set_init(vio_pnp *p, const double *header, const double *P, const double *R, const double *V, const double *Ba, const double *Bg);
/* n consecutive IMU samples: dt[n][batch], acc[n][batch][3], gyr[n][batch][3] */
int vio_pnp_process_imu(vio_pnp *p, int n, const double *dt, const double *acc, const double *gyr);
/* the frame's landmarks with a solved position (ids ascending per stream): counts[batch], ids[batch][max_cnt], obs_xy[batch][max_cnt][2]
 * (normalised image coordinates), pos_xyz[batch][max_cnt][3] (world), track_num[batch][max_cnt], headers[batch] */
int vio_pnp_process_image(vio_pnp *p, const int32_t *counts, const int32_t *ids, const double *obs_xy, const double *pos_xyz,
                          const int32_t *track_num, const double *headers, int use_pnp);
/* window of stream s: P[7][3], R[7][9], V[7][3], headers[7], find_solved[7]; info = {frame_count, error, iterations}; cost = {initial, final}.
 * FeatureTracker::solveVinsPnP returns index PNP_SIZE - 1 = 5. */
int vio_pnp_get_state(vio_pnp *p, int s, double *P, double *R, double *V, double *headers, int32_t *find_solved, int32_t info[3], double cost[2]);
int64_t vio_pnp_launch_count(const vio_pnp *p);

#ifdef __cplusplus
}
#endif
#endif /* VIO_B200_H */
