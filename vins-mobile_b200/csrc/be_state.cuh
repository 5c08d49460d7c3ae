// be_state.cuh -- device-resident state of the batched estimator (one VINS object per stream, VINS.hpp:51-172).
// All arrays are [batch][...] with the per-stream extent given by the stride fields; a kernel gets the struct by value and
// indexes with blockIdx.x = stream.
#pragma once
#include "be_factors.cuh"

namespace be {

// per-stream integer scalars (BeState::iv)
enum { IV_FRAME_COUNT = 0, IV_FIRST_IMU, IV_SOLVER_FLAG, IV_MARG_FLAG, IV_FAILURE, IV_NFEAT, IV_LAST_TRACK, IV_ACTION, IV_INIT_PENDING,
       IV_PRIOR_VALID, IV_N_LM, IV_N_FAC, IV_ITERS, IV_PRIOR_N, IV_ERR, IV_MARG_FAST, IV_MARG_SWEEPS, IV_MARG_M, IV_CHOL_RETRY,
       IV_N_FAC_ALL /* IV_N_FAC + loop-closure factors */, IV_LOOP_FRAME /* window frame the loop pose is tied to, -1: none */, IV_LOOP_NFAC,
       IV_ALLKEY /* no MARGIN_SECOND_NEW slide since the stream (re)started: all_image_frame == the window's frames */,
       IV_ALIGN_OK /* result of the last VisualIMUAlignment, -1: none yet */,
       IV_AF_N /* frames in all_image_frame (closed records of the af_* list; record IV_AF_N is the open tmp_pre_integration) */, IV_COUNT = 32 };
// per-stream double scalars (BeState::dv)
enum { DV_ACC0 = 0, DV_GYR0 = 3, DV_LAST_P = 6, DV_LAST_P_OLD = 9, DV_BACK_P0 = 12, DV_LAST_R = 15, DV_LAST_R_OLD = 24, DV_BACK_R0 = 33,
       DV_COST0 = 42, DV_COST1 = 43, DV_PRIOR_C0 = 44, DV_TIC = 45, DV_RIC = 48,
       DV_INIT_SCALE = 57 /* metric scale of the last accepted alignment */, DV_INIT_G = 58 /* 3: vins.g after the last alignment */, DV_COUNT = 64 };
// IV_ACTION values decided by the feature kernel (VINS::processImage control flow, VINS.cpp:377-478)
enum { ACT_ACCUMULATE = 0, ACT_INIT_SOLVE = 1, ACT_SLIDE_ONLY = 2, ACT_NL_SOLVE = 3, ACT_NONE = 4, ACT_CLEAR = 5 /* track_num < 20 at frame_count == W: clearState(), VINS.cpp:401-405 */ };

struct BeState {
    int B, W, NF;            // NF = W + 1 frames
    int NP;                  // 15 * NF   pose+speed-bias local size
    int NPX;                 // NP + 6    (+ ex_pose)
    int NPW;                 // 6 * NF    pose-only local size (landmark coupling rows)
    // The solve may carry one more pose than the window: the loop-closure pose ("12th pose", VINS.cpp:571-637), stored as solve frame NF.
    // NFS = NF + loop_on; the solve's dimensions are NPS = 15 NFS (the extra speed-bias dofs have no factor: they stay exactly 0) and
    // NPWS = 6 NFS; everything that belongs to the WINDOW (prior, IMU factors, marginalisation, slide) keeps NF / NP / NPX.
    int loop_on, NFS, NPS, NPWS;
    size_t par_stride;       // doubles per stream in par / cand: 16 * NFS + LCAP
    int FCAP, LCAP, PCAP, MAXIMU, MAXCNT;
    double gravity, min_parallax, init_depth, sqrt_info;
    double noise[6];         // acc_n^2, gyr_n^2, acc_n^2, gyr_n^2, acc_w^2, gyr_w^2   (integration_base.h:37-43)
    int max_iters;
    int eig_mode;            // 0 = parallel Jacobi, 1 = Householder tridiagonalisation + implicit QL (default)
    int force_slow_marg;     // test hook: always take the eigendecomposition path for Amm^+
    int marg_direct;         // 1 (default): new prior in information form straight from A_r; 0: the reference's eigendecomposition of A_r
    // window state
    double *Ps, *Rs, *Vs, *Bas, *Bgs, *Headers;       // [B][NF][3|9|3|3|3|1]
    double *pre;                                      // [B][NF][PR_STRIDE]
    double *imu_buf; int *imu_cnt;                    // [B][NF][MAXIMU][7], [B][NF]
    int *iv; double *dv;                              // [B][IV_COUNT], [B][DV_COUNT]
    double *init_state;                               // [B][NF*10 + 6]  P3 Q4 V3 per frame, Ba3 Bg3
    double *init_sfm;                                 // [B][FA][9] ImageFrame::R then [B][FA][3] ImageFrame::T then (int) [B] frame counts; lazily allocated
    // all_image_frame (VINS.hpp:141, filled while solver_flag == INITIAL, VINS.cpp:392-398): one record per camera frame since the stream
    // (re)started -- header, the IMU samples since the previous frame (tmp_pre_integration), the acc_0 / gyr_0 it started from, and the
    // gyroscope bias its pre-integration is currently linearised at.  Record IV_AF_N is the open one.
    int FA;                                           // capacity (3 NF)
    double *af_hdr, *af_imu0, *af_imu, *af_abg;       // [B][FA], [B][FA][6], [B][FA][MAXIMU][7], [B][FA][3]
    int *af_cnt;                                      // [B][FA]
    // feature table (FeatureManager::feature, in the list's insertion order; compacted order-preservingly)
    int *f_id, *f_start, *f_nobs, *f_flag; double *f_depth; double *f_obs;   // [B][FCAP], obs [B][FCAP][NF][2]
    // prior (MarginalizationInfo in information form, canonical layout [pose_i(6) sb_i(9)]_i ex(6))
    double *Hp, *bp; double *x0;                      // [B][NPX*NPX], [B][NPX], [B][NF*16+7]
    int *present;                                     // [B][2*NF+1]
    // solve workspace
    double *par, *cand;                               // [B][NF*16 + LCAP]   pose7 sb9 per frame, then inverse depths
    int *lm_slot, *fac_lm, *fac_j;                    // [B][LCAP], [B][PCAP], [B][PCAP]
    int *lm_fac0, *lm_anchor;                         // [B][LCAP+1] first factor (landmark order) of each landmark, [B][LCAP] anchor frame
    int *fac_sorted, *pair_off; double *fac_obs;      // factors in (anchor i, frame j) order: [B][PCAP] l | i<<16 | j<<24, [B][NF*NF+1], [B][PCAP][4]
    double *scratch; size_t scratch_stride;           // [B][scratch_stride] doubles
    double *post_solve;                               // [B][NF][16]
    double *state_out;                                // [B][NF][16] packed P,Q,V,Ba,Bg after the step
    long long *prof;                                  // [B][32] clock64 cycles per kernel phase (diagnostics)
    // loop closure (RetriveData, VINS.hpp:28-45): matched keyframe header, ids + measurements of the old keyframe (ids ascending), its pose
    int *loop_n; double *loop_hdr; int *loop_ids; double *loop_xy; double *loop_old;   // [B], [B], [B][MAXCNT], [B][MAXCNT][2], [B][7] P_old, Q_old (xyzw)
    int *lm_loop;                                     // [B][LCAP] index into loop_ids/xy of the landmark's loop factor, -1: none
    double *loop_out;                                 // [B][20] relative_t3 relative_q4(xyzw) relative_yaw1 drift_yaw1 t_drift3 (valid flag at [12])
};

__device__ __forceinline__ double *S_Ps(const BeState &s, int b, int i) { return s.Ps + ((size_t)b * s.NF + i) * 3; }
__device__ __forceinline__ double *S_Rs(const BeState &s, int b, int i) { return s.Rs + ((size_t)b * s.NF + i) * 9; }
__device__ __forceinline__ double *S_Vs(const BeState &s, int b, int i) { return s.Vs + ((size_t)b * s.NF + i) * 3; }
__device__ __forceinline__ double *S_Bas(const BeState &s, int b, int i) { return s.Bas + ((size_t)b * s.NF + i) * 3; }
__device__ __forceinline__ double *S_Bgs(const BeState &s, int b, int i) { return s.Bgs + ((size_t)b * s.NF + i) * 3; }
__device__ __forceinline__ double *S_pre(const BeState &s, int b, int i) { return s.pre + ((size_t)b * s.NF + i) * PR_STRIDE; }
__device__ __forceinline__ double *S_imu(const BeState &s, int b, int i) { return s.imu_buf + ((size_t)b * s.NF + i) * s.MAXIMU * 7; }
__device__ __forceinline__ int *S_iv(const BeState &s, int b) { return s.iv + (size_t)b * IV_COUNT; }
__device__ __forceinline__ double *S_dv(const BeState &s, int b) { return s.dv + (size_t)b * DV_COUNT; }
__device__ __forceinline__ double *S_obs(const BeState &s, int b, int slot) { return s.f_obs + ((size_t)b * s.FCAP + slot) * s.NF * 2; }
__device__ __forceinline__ double *S_par_pose(const BeState &s, double *par, int i) { return par + 16 * i; }
__device__ __forceinline__ double *S_par_sb(const BeState &s, double *par, int i) { return par + 16 * i + 7; }

// phase timer: thread 0 accumulates the cycles since the previous mark into prof[b][slot].  Compiled in only with -DVIO_BE_PROFILE
// (the debug library): every mark is a global read-modify-write on the CTA's critical path (~0.6 us each), dozens per solve.
#ifdef VIO_BE_PROFILE
#define BE_PROF_INIT long long _pt0 = clock64(); long long *_pp = s.prof + (size_t)blockIdx.x * 32
#define BE_PROF2_INIT long long _qt0 = clock64()
#define BE_PROF2(pp, slot) do { if (threadIdx.x == 0) { const long long _t = clock64(); (pp)[slot] += _t - _qt0; _qt0 = _t; } } while (0)
#define BE_PROF(slot) do { if (threadIdx.x == 0) { const long long _t = clock64(); _pp[slot] += _t - _pt0; _pt0 = _t; } } while (0)
#define BE_PROF_ONLY(...) __VA_ARGS__
#else
#define BE_PROF_INIT do { } while (0)
#define BE_PROF2_INIT do { } while (0)
#define BE_PROF2(pp, slot) do { (void)(pp); } while (0)
#define BE_PROF(slot) do { } while (0)
#define BE_PROF_ONLY(...)
#endif

// the predicate repeated throughout the reference: used_num >= 2 && start_frame < WINDOW_SIZE - 2
__device__ __forceinline__ bool in_solve(const BeState &s, int nobs, int start) { return nobs >= 2 && start < s.W - 2; }

}  // namespace be
