// backend.cu -- C-ABI of the batched sliding-window estimator (include/vio_b200.h, vio_backend_*), replacing
// VINS::processIMU / processImage / solve_ceres / slideWindow (/root/reference/VINS_ios/VINS.cpp:333-831,1149-1273) and
// FeatureManager (feature_manager.cpp).  Every stream of the batch is one VINS object; all state lives in HBM and every
// step is a fixed sequence of kernels with one CTA per stream -- no host round trip inside processImage.
// No CPU fallback.
#include "be_marg.cuh"
#include "be_align.cuh"

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>
#include <string.h>
#include <stdlib.h>

using namespace be;

struct vio_backend {
    vio_config cfg;
    BeState s;
    cudaStream_t stream;
    bool own_stream;
    KernelTimer timer;
    int64_t launches;
    std::vector<void *> allocs;
    // device staging for the host-pointer entry points
    int *d_counts, *d_ids; double *d_xyz, *d_headers; double *d_imu; size_t imu_cap;
    double *h_headers_pinned;
    size_t solve_smem, marg_smem;
    int use_smem_solve, solve_vec_off;
    int be_threads;
    cudaEvent_t evt_ready, evt_consumed;
    bool consumed_valid, record_consumed;
    // initialisation from SfM poses (vio_backend_set_init_sfm): alignment arguments over the back end's own arrays + scratch, lazily allocated
    AlignArgs align; bool sfm_armed;
};

template <typename T>
static int dalloc(vio_backend *be, T **p, size_t n) {
    VIO_CUDA_TRY(vio_dev_alloc((void **)p, n * sizeof(T), be->allocs));
    return VIO_OK;
}

__global__ void init_state_kernel(BeState s, const double *tic, const double *ric) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        double *dv = S_dv(s, b);
        for (int i = 0; i < 3; i++) dv[DV_TIC + i] = tic[i];
        for (int i = 0; i < 9; i++) dv[DV_RIC + i] = ric[i];
        stm(dv + DV_LAST_R, eye3()); stm(dv + DV_LAST_R_OLD, eye3()); stm(dv + DV_BACK_R0, eye3());
        int *iv = S_iv(s, b);
        for (int i = 0; i < IV_COUNT; i++) iv[i] = 0;
        iv[IV_ACTION] = ACT_NONE;
    }
    __syncthreads();
    clear_state_cta(s, b);
}

extern "C" int vio_backend_clear(vio_backend *be) {
    if (!be) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    double h[12];
    memcpy(h, be->cfg.tic, sizeof(double) * 3);
    memcpy(h + 3, be->cfg.ric, sizeof(double) * 9);
    double *ext = be->d_xyz + (size_t)be->s.B * be->cfg.max_cnt * 3;          // 16 spare doubles behind the image_msg copy
    VIO_CUDA_TRY(cudaMemcpyAsync(ext, h, sizeof(h), cudaMemcpyHostToDevice, be->stream));
    init_state_kernel<<<be->s.B, 128, 0, be->stream>>>(be->s, ext, ext + 3);
    be->launches++;
    VIO_CUDA_TRY(cudaGetLastError());
    return VIO_OK;
}

extern "C" int vio_backend_create(const vio_config *cfg, vio_backend **out) {
    if (!cfg || !out || cfg->batch < 1 || cfg->window_size < 4 || cfg->window_size > VIO_MAX_WIN || cfg->max_cnt < 1 || cfg->max_cnt > VIO_MAXP ||
        cfg->num_of_f < 1 || cfg->max_imu_per_frame < 1)
        return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    vio_poison_load_mask();
    vio_backend *be = new (std::nothrow) vio_backend();
    if (!be) return VIO_ERR_ARG;
    be->cfg = *cfg; be->launches = 0; be->own_stream = true; be->consumed_valid = false; be->record_consumed = false;
    VIO_CUDA_TRY_OR(cudaEventCreateWithFlags(&be->evt_ready, cudaEventDisableTiming), vio_backend_destroy(be));
    VIO_CUDA_TRY_OR(cudaEventCreateWithFlags(&be->evt_consumed, cudaEventDisableTiming), vio_backend_destroy(be));
    VIO_CUDA_TRY_OR(cudaStreamCreateWithFlags(&be->stream, cudaStreamNonBlocking), vio_backend_destroy(be));
    BeState &s = be->s;
    memset(&s, 0, sizeof(s));
    s.B = cfg->batch; s.W = cfg->window_size; s.NF = s.W + 1; s.NP = 15 * s.NF; s.NPX = s.NP + 6; s.NPW = 6 * s.NF;
    s.loop_on = cfg->loop_closure ? 1 : 0; s.NFS = s.NF + s.loop_on; s.NPS = 15 * s.NFS; s.NPWS = 6 * s.NFS;
    s.MAXCNT = cfg->max_cnt;
    s.FCAP = std::min(8192, (s.NF + 1) * cfg->max_cnt);
    s.LCAP = cfg->num_of_f;
    s.par_stride = (size_t)s.NFS * 16 + s.LCAP;
    s.PCAP = std::min(cfg->num_of_f * s.W, 16384);
    s.MAXIMU = cfg->max_imu_per_frame;
    s.gravity = cfg->gravity; s.min_parallax = cfg->min_parallax; s.init_depth = cfg->init_depth;
    s.sqrt_info = cfg->fx / 1.5;                                   // ProjectionFactor::sqrt_info, VINS.cpp:29-32
    s.noise[0] = s.noise[2] = cfg->acc_n * cfg->acc_n; s.noise[1] = s.noise[3] = cfg->gyr_n * cfg->gyr_n;
    s.noise[4] = cfg->acc_w * cfg->acc_w; s.noise[5] = cfg->gyr_w * cfg->gyr_w;
    s.max_iters = cfg->max_iters;
    s.eig_mode = cfg->marg_eig == 1 ? 0 : 1;
    be->be_threads = cfg->be_threads == 256 ? 256 : 512;
    s.force_slow_marg = cfg->marg_amm_eig ? 1 : 0;
    s.marg_direct = cfg->marg_mode == 1 ? 0 : 1;
    const size_t B = s.B, NF = s.NF;
    int rc = VIO_OK;
    if (!rc) rc = dalloc(be, &s.Ps, B * NF * 3);
    if (!rc) rc = dalloc(be, &s.Rs, B * NF * 9);
    if (!rc) rc = dalloc(be, &s.Vs, B * NF * 3);
    if (!rc) rc = dalloc(be, &s.Bas, B * NF * 3);
    if (!rc) rc = dalloc(be, &s.Bgs, B * NF * 3);
    if (!rc) rc = dalloc(be, &s.Headers, B * NF);
    if (!rc) rc = dalloc(be, &s.pre, B * NF * PR_STRIDE);
    if (!rc) rc = dalloc(be, &s.imu_buf, B * NF * s.MAXIMU * 7);
    if (!rc) rc = dalloc(be, &s.imu_cnt, B * NF);
    if (!rc) rc = dalloc(be, &s.iv, B * IV_COUNT);
    s.FA = 3 * s.NF;
    if (!rc) rc = dalloc(be, &s.af_hdr, B * s.FA);
    if (!rc) rc = dalloc(be, &s.af_imu0, B * s.FA * 6);
    if (!rc) rc = dalloc(be, &s.af_imu, B * s.FA * (size_t)s.MAXIMU * 7);
    if (!rc) rc = dalloc(be, &s.af_abg, B * s.FA * 3);
    if (!rc) rc = dalloc(be, &s.af_cnt, B * s.FA);
    if (!rc) rc = dalloc(be, &s.dv, B * DV_COUNT);
    if (!rc) rc = dalloc(be, &s.init_state, B * (NF * 10 + 6));
    if (!rc) rc = dalloc(be, &s.f_id, B * s.FCAP);
    if (!rc) rc = dalloc(be, &s.f_start, B * s.FCAP);
    if (!rc) rc = dalloc(be, &s.f_nobs, B * s.FCAP);
    if (!rc) rc = dalloc(be, &s.f_flag, B * s.FCAP);
    if (!rc) rc = dalloc(be, &s.f_depth, B * s.FCAP);
    if (!rc) rc = dalloc(be, &s.f_obs, B * s.FCAP * NF * 2);
    if (!rc) rc = dalloc(be, &s.Hp, B * s.NPX * s.NPX);
    if (!rc) rc = dalloc(be, &s.bp, B * s.NPX);
    if (!rc) rc = dalloc(be, &s.x0, B * (NF * 16 + 7));
    if (!rc) rc = dalloc(be, &s.present, B * (2 * NF + 1));
    if (!rc) rc = dalloc(be, &s.par, B * s.par_stride);
    if (!rc) rc = dalloc(be, &s.cand, B * s.par_stride);
    if (!rc) rc = dalloc(be, &s.loop_n, B);
    if (!rc) rc = dalloc(be, &s.loop_hdr, B);
    if (!rc) rc = dalloc(be, &s.loop_ids, B * (size_t)cfg->max_cnt);
    if (!rc) rc = dalloc(be, &s.loop_xy, B * (size_t)cfg->max_cnt * 2);
    if (!rc) rc = dalloc(be, &s.loop_old, B * 7);
    if (!rc) rc = dalloc(be, &s.lm_loop, B * (size_t)s.LCAP);
    if (!rc) rc = dalloc(be, &s.loop_out, B * 20);
    if (!rc) rc = dalloc(be, &s.lm_slot, B * s.LCAP);
    if (!rc) rc = dalloc(be, &s.fac_lm, B * s.PCAP);
    if (!rc) rc = dalloc(be, &s.fac_j, B * s.PCAP);
    if (!rc) rc = dalloc(be, &s.lm_fac0, B * (s.LCAP + 1));
    if (!rc) rc = dalloc(be, &s.lm_anchor, B * s.LCAP);
    if (!rc) rc = dalloc(be, &s.fac_sorted, B * s.PCAP);
    if (!rc) rc = dalloc(be, &s.pair_off, B * ((size_t)s.NFS * s.NFS + 1));
    if (!rc) rc = dalloc(be, &s.fac_obs, B * s.PCAP * 4);
    if (!rc) rc = dalloc(be, &s.post_solve, B * NF * 16);
    if (!rc) rc = dalloc(be, &s.state_out, B * NF * 16);
    if (!rc) rc = dalloc(be, &s.prof, B * 32);
    size_t sc = solve_scratch_doubles(s.NPS, s.NPX, s.NPWS, s.LCAP, s.PCAP);
    sc = std::max(sc, marg_scratch_doubles(s.NPX, s.LCAP, cfg->max_cnt));
    sc = std::max(sc, (size_t)s.FCAP * (5 + 2 * NF));
    s.scratch_stride = (sc + 15) & ~(size_t)15;
    if (!rc) rc = dalloc(be, &s.scratch, B * s.scratch_stride);
    if (!rc) rc = dalloc(be, &be->d_counts, B);
    if (!rc) rc = dalloc(be, &be->d_ids, B * cfg->max_cnt);
    if (!rc) rc = dalloc(be, &be->d_xyz, B * cfg->max_cnt * 3 + 16);
    if (!rc) rc = dalloc(be, &be->d_headers, B);
    be->imu_cap = 64;
    if (!rc) rc = dalloc(be, &be->d_imu, be->imu_cap * B * 7);
    if (!rc && cudaMallocHost((void **)&be->h_headers_pinned, B * sizeof(double)) != cudaSuccess) rc = VIO_ERR_CUDA;
    if (rc) { vio_backend_destroy(be); return rc; }
    size_t dyn_max = 0;
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(finish_kernel, cfg->device, &dyn_max), vio_backend_destroy(be));
    if ((size_t)s.FCAP + 64 > dyn_max) { vio_backend_destroy(be); return VIO_ERR_CAPACITY; }
    // reduced system resident in shared memory when it fits one SM (W = 10: 110 KB packed); otherwise the global-memory path
    be->solve_smem = solve_smem_bytes(s.NPS, s.NPWS);
    if (be->solve_smem > 200 * 1024 || schur_ntile(s.NPWS) > be->be_threads) be->solve_smem = 0;
    be->use_smem_solve = be->solve_smem > 0;
    // reduced system in registers as DMMA tiles (be_tilechol.cuh) when it fits 16 warps x 16 tiles; vio_config::solve_path = 1 keeps the packed path
    if (be->use_smem_solve && cfg->solve_path != 1 && tile_path_fits(s.NFS, be->be_threads)) {
        be->use_smem_solve = 2;
        be->solve_smem = std::max(be->solve_smem, tile_smem_doubles(s.NFS) * sizeof(double));
    }
    be->solve_smem = std::max(be->solve_smem, eval_smem_bytes(s.NFS - 1));
    if (!be->use_smem_solve)                       // global-memory path: Schur chunk staging and the Cholesky panel (used one after the other)
        be->solve_smem = std::max(be->solve_smem, std::max((size_t)SCHUR_CHUNK * schur_ld(s.NPWS), chol_global_smem_doubles(s.NPS)) * sizeof(double));
    be->solve_vec_off = (int)((be->solve_smem / sizeof(double) + 3) & ~(size_t)3);
    be->solve_smem = (be->solve_vec_off + solve_vec_doubles(s.NPS, s.NPX)) * sizeof(double);
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(solve_kernel, cfg->device, &dyn_max), vio_backend_destroy(be));
    if (be->solve_smem > dyn_max) { vio_backend_destroy(be); return VIO_ERR_CAPACITY; }
    be->marg_smem = sizeof(MargSmem) + 16 + (size_t)2 * MARG_NCAP * MARG_NCAP * sizeof(double);
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(marg_kernel, cfg->device, &dyn_max), vio_backend_destroy(be));
    if (be->marg_smem > dyn_max) { vio_backend_destroy(be); return VIO_ERR_CAPACITY; }
    // dalloc() zero-fills with cudaMemset on the legacy default stream, which does not order against this handle's non-blocking
    // stream (nor against a caller's non-blocking stream given to vio_backend_use_stream): finish the fills before any kernel
    VIO_CUDA_TRY_OR(cudaDeviceSynchronize(), vio_backend_destroy(be));
    rc = vio_backend_clear(be);
    if (rc) { vio_backend_destroy(be); return rc; }
    VIO_CUDA_TRY_OR(cudaStreamSynchronize(be->stream), vio_backend_destroy(be));
    *out = be;
    return VIO_OK;
}

extern "C" void vio_backend_destroy(vio_backend *be) {
    if (!be) return;
    cudaSetDevice(be->cfg.device);
    if (be->stream) cudaStreamSynchronize(be->stream);
    for (void *p : be->allocs) cudaFree(p);
    if (be->evt_ready) cudaEventDestroy(be->evt_ready);
    if (be->evt_consumed) cudaEventDestroy(be->evt_consumed);
    if (be->h_headers_pinned) cudaFreeHost(be->h_headers_pinned);
    if (be->own_stream && be->stream) cudaStreamDestroy(be->stream);
    delete be;
}

extern "C" int vio_backend_process_imu_dev(vio_backend *be, int n, const double *dt, const double *acc, const double *gyr) {
    if (!be || n < 0) return VIO_ERR_ARG;
    if (n == 0) return VIO_OK;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    VIO_LAUNCH(be->timer, be->stream, "imu_kernel", (imu_kernel<<<(be->s.B + 3) / 4, 128, 0, be->stream>>>(be->s, n, dt, acc, gyr)));
    be->launches++;
    VIO_CUDA_TRY(cudaGetLastError());
    return VIO_OK;
}

extern "C" int vio_backend_process_imu(vio_backend *be, int n, const double *dt, const double *acc, const double *gyr) {
    if (!be || n < 0 || (n > 0 && (!dt || !acc || !gyr))) return VIO_ERR_ARG;
    if (n == 0) return VIO_OK;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t B = be->s.B;
    if ((size_t)n > be->imu_cap) {
        VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
        double *p;
        VIO_CUDA_TRY(cudaMalloc((void **)&p, (size_t)n * B * 7 * sizeof(double)));
        be->allocs.push_back(p);
        be->d_imu = p; be->imu_cap = n;
    }
    double *d_dt = be->d_imu, *d_acc = d_dt + (size_t)n * B, *d_gyr = d_acc + (size_t)n * B * 3;
    VIO_CUDA_TRY(cudaMemcpyAsync(d_dt, dt, (size_t)n * B * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(d_acc, acc, (size_t)n * B * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(d_gyr, gyr, (size_t)n * B * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    return vio_backend_process_imu_dev(be, n, d_dt, d_acc, d_gyr);
}

__global__ void set_init_kernel(BeState s, const double *P, const double *Q, const double *V, const double *Ba, const double *Bg) {
    const int b = blockIdx.x;
    double *o = s.init_state + (size_t)b * (s.NF * 10 + 6);
    for (int i = threadIdx.x; i < s.NF; i += blockDim.x) {
        const size_t k = (size_t)b * s.NF + i;
        for (int c = 0; c < 3; c++) { o[10 * i + c] = P[3 * k + c]; o[10 * i + 7 + c] = V[3 * k + c]; }
        for (int c = 0; c < 4; c++) o[10 * i + 3 + c] = Q[4 * k + c];
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; c++) { o[10 * s.NF + c] = Ba[3 * b + c]; o[10 * s.NF + 3 + c] = Bg[3 * b + c]; }
        S_iv(s, b)[IV_INIT_PENDING] = 1;
    }
}

extern "C" int vio_backend_set_init_window(vio_backend *be, const double *P, const double *Q, const double *V, const double *Ba, const double *Bg) {
    if (!be || !P || !Q || !V || !Ba || !Bg) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t B = be->s.B, NF = be->s.NF;
    double *d;
    const size_t tot = B * NF * 10 + B * 6;
    VIO_CUDA_TRY(cudaMalloc((void **)&d, tot * sizeof(double)));
    double *dP = d, *dQ = dP + B * NF * 3, *dV = dQ + B * NF * 4, *dBa = dV + B * NF * 3, *dBg = dBa + B * 3;
    cudaMemcpyAsync(dP, P, B * NF * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream);
    cudaMemcpyAsync(dQ, Q, B * NF * 4 * sizeof(double), cudaMemcpyHostToDevice, be->stream);
    cudaMemcpyAsync(dV, V, B * NF * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream);
    cudaMemcpyAsync(dBa, Ba, B * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream);
    cudaMemcpyAsync(dBg, Bg, B * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream);
    set_init_kernel<<<be->s.B, 32, 0, be->stream>>>(be->s, dP, dQ, dV, dBa, dBg);
    be->launches++;
    cudaError_t e = cudaStreamSynchronize(be->stream);
    cudaFree(d);
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

__global__ void set_sfm_pending_kernel(BeState s) {
    if (threadIdx.x == 0) S_iv(s, blockIdx.x)[IV_INIT_PENDING] = 2;
}

// ImageFrame::R / T of every frame of all_image_frame as VINS::solveInitial leaves them after the global SfM and the PnP of the
// non-keyframes (VINS.cpp:889-958: R = body attitude in the SfM frame, T = camera position in it, unknown scale), in map (= time) order;
// n_frames[b] must equal the number of frames the back end holds for the stream (vio_backend_get_init_frames).  Consumed by the
// vio_backend_process_image call that fills the window: VisualIMUAlignment + the rest of visualInitialAlign (VINS.cpp:1022-1102) run on
// the device, then the first solve (VINS.cpp:415-447).
extern "C" int vio_backend_set_init_sfm_frames(vio_backend *be, const int32_t *n_frames, int max_frames, const double *R, const double *T) {
    if (!be || !n_frames || !R || !T || max_frames < 2) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    BeState &s = be->s;
    const size_t B = s.B, NF = s.NF, FA = s.FA, NS = 3 * FA + 4;
    for (size_t b = 0; b < B; b++) if (n_frames[b] < 0 || n_frames[b] > max_frames || n_frames[b] > (int)FA) return VIO_ERR_CAPACITY;
    if (!s.init_sfm) {
        int rc = dalloc(be, &s.init_sfm, B * FA * 12 + (B + 1) / 2 + 1);       // R, T, then B ints
        AlignArgs &a = be->align;
        memset(&a, 0, sizeof(a));
        a.B = s.B; a.F = (int)FA; a.MAXIMU = s.MAXIMU; a.NS = (int)NS;
        double *scr = nullptr; int *iscr = nullptr;
        const size_t nd = B * FA * PR_STRIDE + 2 * B * NS * NS + B * NS + B * FA * 110 + 2 * B * NS + B * 3 + B * 3 + B * NS;
        if (!rc) rc = dalloc(be, &scr, nd);
        if (!rc) rc = dalloc(be, &iscr, B * NS + B);
        if (rc) { s.init_sfm = nullptr; return rc; }
        double *p = scr;
        a.pre = p; p += B * FA * PR_STRIDE;
        a.A = p; p += B * NS * NS;
        a.Aw = p; p += B * NS * NS;
        a.rhs = p; p += B * NS;
        a.pairs = p; p += B * FA * 110;
        a.xs = p; p += 2 * B * NS;
        a.bgs_out = p; p += B * 3;
        a.g_out = p; p += B * 3;
        a.x_out = p; p += B * NS;
        a.perm = iscr; a.ok = iscr + B * NS;
        a.R = s.init_sfm; a.T = s.init_sfm + B * FA * 9;
        a.counts = s.af_cnt; a.imu = s.af_imu;
        a.imu0 = s.af_imu0; a.imu0_stride = 6;
        a.bg0 = s.Bgs; a.bg0_stride = (int)(3 * NF);
        a.abg = s.af_abg; a.abg_stride = 3;
        memcpy(a.tic, be->cfg.tic, 24);
        a.g_norm = be->cfg.gravity; a.g_thr = 3.0;          // G_NORM, G_THRESHOLD (global_param.hpp:49-50)
        memcpy(a.noise, s.noise, sizeof(a.noise));
    }
    std::vector<double> h(B * FA * 12, 0.0);
    for (size_t b = 0; b < B; b++)
        for (int k = 0; k < n_frames[b]; k++) {
            memcpy(&h[(b * FA + k) * 9], R + (b * max_frames + k) * 9, 72);
            memcpy(&h[B * FA * 9 + (b * FA + k) * 3], T + (b * max_frames + k) * 3, 24);
        }
    VIO_CUDA_TRY(cudaMemcpyAsync(s.init_sfm, h.data(), B * FA * 12 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(s.init_sfm + B * FA * 12, n_frames, B * sizeof(int), cudaMemcpyHostToDevice, be->stream));
    set_sfm_pending_kernel<<<s.B, 32, 0, be->stream>>>(s);
    be->launches++;
    be->sfm_armed = true;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));        // h and n_frames are host memory of this call
    return VIO_OK;
}

// The common case: every frame of the map is a keyframe (no MARGIN_SECOND_NEW slide since the stream started), so the map's frames are the
// window's W + 1 frames.  R [batch][W+1][9], T [batch][W+1][3].
extern "C" int vio_backend_set_init_sfm(vio_backend *be, const double *R, const double *T) {
    if (!be) return VIO_ERR_ARG;
    std::vector<int32_t> n(be->s.B, be->s.NF);
    return vio_backend_set_init_sfm_frames(be, n.data(), be->s.NF, R, T);
}

// Headers of the frames the back end holds in all_image_frame for stream s (what the SfM must deliver poses for), oldest first.
extern "C" int vio_backend_get_init_frames(vio_backend *be, int sidx, int cap, int32_t *n, double *headers) {
    if (!be || !n || !headers || sidx < 0 || sidx >= be->s.B || cap < 0) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    int an = 0;
    VIO_CUDA_TRY(cudaMemcpy(&an, be->s.iv + (size_t)sidx * IV_COUNT + IV_AF_N, sizeof(int), cudaMemcpyDeviceToHost));
    if (an > be->s.FA) return VIO_ERR_CAPACITY;             // the list overflowed (more than 3 (W + 1) - 1 frames without an initialisation)
    *n = an;
    if (an > cap) return VIO_ERR_CAPACITY;
    if (an > 0) VIO_CUDA_TRY(cudaMemcpy(headers, be->s.af_hdr + (size_t)sidx * be->s.FA, an * sizeof(double), cudaMemcpyDeviceToHost));
    return VIO_OK;
}

// Outcome of the last VisualIMUAlignment of stream s: ok = 1 / 0 (-1: none yet since the stream started), g = gravity in the window's frame
// after the alignment (vins.g, VINS.cpp:1086-1091; ~ (0, 0, G_NORM)), scale = the metric scale applied to the SfM.
extern "C" int vio_backend_get_init_result(vio_backend *be, int sidx, int32_t *ok, double g[3], double *scale) {
    if (!be || !ok || !g || !scale || sidx < 0 || sidx >= be->s.B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    int iv[IV_COUNT];
    VIO_CUDA_TRY(cudaMemcpy(iv, be->s.iv + (size_t)sidx * IV_COUNT, sizeof(iv), cudaMemcpyDeviceToHost));
    *ok = iv[IV_ALIGN_OK]; g[0] = g[1] = g[2] = 0; *scale = 0;
    if (iv[IV_ALIGN_OK] >= 0) {
        double dvh[4];
        VIO_CUDA_TRY(cudaMemcpy(dvh, be->s.dv + (size_t)sidx * DV_COUNT + DV_INIT_SCALE, sizeof(dvh), cudaMemcpyDeviceToHost));
        *scale = dvh[0]; g[0] = dvh[1]; g[1] = dvh[2]; g[2] = dvh[3];
    }
    return VIO_OK;
}

__global__ void clear_init_pending_kernel(BeState s) {
    int *iv = S_iv(s, blockIdx.x);
    if (threadIdx.x == 0 && iv[IV_ACTION] == ACT_INIT_SOLVE) iv[IV_INIT_PENDING] = 0;
}

static int run_process_image(vio_backend *be, const int32_t *counts, const int32_t *ids, const double *xyz, const double *headers_dev) {
    BeState &s = be->s;
    cudaStream_t st = be->stream;
    VIO_LAUNCH(be->timer, st, "addfeat_kernel", (addfeat_kernel<<<s.B, 256, 0, st>>>(s, counts, ids, xyz, headers_dev)));
    if (be->record_consumed) { cudaEventRecord(be->evt_consumed, st); be->consumed_valid = true; be->record_consumed = false; }   // image_msg fully read
    if (be->sfm_armed) {        // VINS::visualInitialAlign for the streams whose window fills with SfM poses pending; a no-op for the others
        VIO_LAUNCH(be->timer, st, "init_align_kernel", (init_align_kernel<<<s.B, ALIGN_THREADS, 0, st>>>(s, be->align)));
        be->launches += 1;
    }
    VIO_LAUNCH(be->timer, st, "triangulate_kernel", (triangulate_kernel<<<s.B, 128, 0, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "prepare_kernel", (prepare_kernel<<<s.B, 256, 0, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "solve_kernel", (solve_kernel<<<s.B, be->be_threads, be->solve_smem, st>>>(s, be->use_smem_solve, be->solve_vec_off)));
    VIO_LAUNCH(be->timer, st, "post_solve_kernel", (post_solve_kernel<<<s.B, 256, 0, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "marg_kernel", (marg_kernel<<<s.B, be->be_threads, be->marg_smem, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "finish_kernel", (finish_kernel<<<s.B, 256, s.FCAP + 64, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "clear_init_pending_kernel", (clear_init_pending_kernel<<<s.B, 32, 0, st>>>(s)));
    be->launches += 8;
    VIO_CUDA_TRY(cudaGetLastError());
    return VIO_OK;
}

// retrive_pose_data (VINS.hpp:28-45, written by the loop-closure thread at ViewController.mm:964): per stream the matched old keyframe --
// header of the window frame it was matched against, ids (ascending) and normalised measurements of the shared features in the OLD
// keyframe, its pose (P_old, Q_old xyzw).  counts[b] = 0 clears the stream's match.  Stays in force until replaced, like front_pose.
extern "C" int vio_backend_set_loop_match(vio_backend *be, const int32_t *counts, const double *headers, const int32_t *ids, const double *xy,
                                          const double *pose_old) {
    if (!be || !counts || !headers || !ids || !xy || !pose_old) return VIO_ERR_ARG;
    if (!be->s.loop_on) return VIO_ERR_STATE;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t B = be->s.B, P = be->cfg.max_cnt;
    for (size_t b = 0; b < B; b++) if (counts[b] < 0 || counts[b] > (int)P) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaMemcpyAsync(be->s.loop_n, counts, B * sizeof(int), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->s.loop_hdr, headers, B * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->s.loop_ids, ids, B * P * sizeof(int), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->s.loop_xy, xy, B * P * 2 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->s.loop_old, pose_old, B * 7 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));           // the caller's buffers may be pageable and short-lived
    return VIO_OK;
}

// what the last solve made of the match (VINS.cpp:664-680, :174-195): out[0..2] relative_t, [3..6] relative_q (xyzw), [7] relative_yaw (deg),
// [8] drift yaw (deg; r_drift = ypr2R(drift_yaw, 0, 0)), [9..11] t_drift.  *n_factors = loop factors in that solve (0: no loop constraint).
extern "C" int vio_backend_get_loop_result(vio_backend *be, int s, double out[12], int32_t *n_factors) {
    if (!be || s < 0 || s >= be->s.B || !out || !n_factors) return VIO_ERR_ARG;
    if (!be->s.loop_on) return VIO_ERR_STATE;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    double lo[20]; int iv[IV_COUNT];
    VIO_CUDA_TRY(cudaMemcpyAsync(lo, be->s.loop_out + (size_t)s * 20, sizeof(lo), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(iv, be->s.iv + (size_t)s * IV_COUNT, sizeof(iv), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    for (int i = 0; i < 12; i++) out[i] = lo[12] != 0.0 ? lo[i] : 0.0;
    *n_factors = lo[12] != 0.0 ? iv[IV_LOOP_NFAC] : 0;
    return VIO_OK;
}

// VINS::solve_ceres() on its own (VINS.hpp:153, VINS.cpp:480-831): problem build (old2new), <= max_iters dogleg iterations, new2old and
// the marginalisation selected by the current marginalization_flag -- on the window as it stands, without the processImage
// bookkeeping around it (no addFeatureCheckParallax, triangulate, failureDetection or slideWindow).  Streams that are not in the
// NON_LINEAR state with a full window are left untouched.
__global__ void solve_only_action_kernel(BeState s, int begin) {
    int *iv = S_iv(s, blockIdx.x);
    if (threadIdx.x != 0) return;
    if (begin) {
        const bool ready = iv[IV_SOLVER_FLAG] == 1 && iv[IV_FRAME_COUNT] == s.W;
        iv[IV_ACTION] = ready ? ACT_NL_SOLVE : ACT_NONE;
        iv[IV_N_LM] = 0; iv[IV_N_FAC] = 0; iv[IV_ITERS] = 0; iv[IV_CHOL_RETRY] = 0; iv[IV_MARG_SWEEPS] = 0; iv[IV_MARG_FAST] = 0;
    } else iv[IV_ACTION] = ACT_NONE;
}

extern "C" int vio_backend_solve(vio_backend *be) {
    if (!be) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    BeState &s = be->s;
    cudaStream_t st = be->stream;
    solve_only_action_kernel<<<s.B, 32, 0, st>>>(s, 1);
    VIO_LAUNCH(be->timer, st, "prepare_kernel", (prepare_kernel<<<s.B, 256, 0, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "solve_kernel", (solve_kernel<<<s.B, be->be_threads, be->solve_smem, st>>>(s, be->use_smem_solve, be->solve_vec_off)));
    VIO_LAUNCH(be->timer, st, "post_solve_kernel", (post_solve_kernel<<<s.B, 256, 0, st>>>(s)));
    VIO_LAUNCH(be->timer, st, "marg_kernel", (marg_kernel<<<s.B, be->be_threads, be->marg_smem, st>>>(s)));
    solve_only_action_kernel<<<s.B, 32, 0, st>>>(s, 0);
    VIO_LAUNCH(be->timer, st, "finish_kernel", (finish_kernel<<<s.B, 256, s.FCAP + 64, st>>>(s)));      // action NONE: refreshes the packed state only
    be->launches += 7;
    VIO_CUDA_TRY(cudaGetLastError());
    return VIO_OK;
}

extern "C" int vio_backend_process_image_dev(vio_backend *be, const int32_t *counts, const int32_t *ids, const double *xyz, const double *headers_host) {
    if (!be || !counts || !ids || !xyz || !headers_host) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    // pageable source: the runtime stages the bytes before returning, so the caller's buffer may be reused immediately
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_headers, headers_host, be->s.B * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    return run_process_image(be, counts, ids, xyz, be->d_headers);
}

extern "C" int vio_backend_process_image(vio_backend *be, const int32_t *counts, const int32_t *ids, const double *xyz, const double *headers) {
    if (!be || !counts || !ids || !xyz || !headers) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t B = be->s.B, P = be->cfg.max_cnt;
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_counts, counts, B * sizeof(int), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_ids, ids, B * P * sizeof(int), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_xyz, xyz, B * P * 3 * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_headers, headers, B * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    return run_process_image(be, be->d_counts, be->d_ids, be->d_xyz, be->d_headers);
}

extern "C" void *vio_frontend_stream(vio_frontend *fe);
// VINS::processImage fed straight from a front end's device-resident image_msg.  If the two handles run on different CUDA streams
// the hand-over is event-ordered (no host synchronisation): the front end may already track the next frames while the solve runs.
extern "C" int vio_backend_process_image_from_frontend(vio_backend *be, vio_frontend *fe, const double *headers_host) {
    if (!be || !fe || !headers_host) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const int32_t *cnt, *ids; const double *xyz;
    int rc = vio_frontend_image_msg_dev(fe, &cnt, &ids, &xyz);
    if (rc) return rc;
    cudaStream_t fs = (cudaStream_t)vio_frontend_stream(fe);
    if (fs == be->stream) return vio_backend_process_image_dev(be, cnt, ids, xyz, headers_host);
    const size_t B = be->s.B, P = be->cfg.max_cnt;
    if (be->consumed_valid) VIO_CUDA_TRY(cudaStreamWaitEvent(fs, be->evt_consumed, 0));     // previous image_msg copy has been read
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_counts, cnt, B * sizeof(int), cudaMemcpyDeviceToDevice, fs));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_ids, ids, B * P * sizeof(int), cudaMemcpyDeviceToDevice, fs));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_xyz, xyz, B * P * 3 * sizeof(double), cudaMemcpyDeviceToDevice, fs));
    VIO_CUDA_TRY(cudaEventRecord(be->evt_ready, fs));
    VIO_CUDA_TRY(cudaStreamWaitEvent(be->stream, be->evt_ready, 0));
    VIO_CUDA_TRY(cudaMemcpyAsync(be->d_headers, headers_host, B * sizeof(double), cudaMemcpyHostToDevice, be->stream));
    be->record_consumed = true;
    return run_process_image(be, be->d_counts, be->d_ids, be->d_xyz, be->d_headers);
}

static int stream_err(vio_backend *be, int s) {
    int e = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&e, be->s.iv + (size_t)s * IV_COUNT + IV_ERR, sizeof(int), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    return e;
}

extern "C" int vio_backend_get_error(vio_backend *be, int s, int clear, int32_t *code) {
    if (!be || s < 0 || s >= be->s.B || !code) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    int e = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&e, be->s.iv + (size_t)s * IV_COUNT + IV_ERR, sizeof(int), cudaMemcpyDeviceToHost, be->stream));
    if (clear) VIO_CUDA_TRY(cudaMemsetAsync(be->s.iv + (size_t)s * IV_COUNT + IV_ERR, 0, sizeof(int), be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    *code = e;
    return VIO_OK;
}

template <typename T>
static int bd2h(vio_backend *be, T *dst, const T *src, size_t n) {
    if (!dst || n == 0) return VIO_OK;
    VIO_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, be->stream));
    return VIO_OK;
}

extern "C" int vio_backend_get_state(vio_backend *be, int s, double *P, double *Q, double *V, double *Ba, double *Bg, double *headers) {
    if (!be || s < 0 || s >= be->s.B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const int NF = be->s.NF;
    std::vector<double> h((size_t)NF * 16), R((size_t)NF * 9);
    int rc = bd2h(be, h.data(), be->s.state_out + (size_t)s * NF * 16, (size_t)NF * 16);
    if (!rc) rc = bd2h(be, headers, be->s.Headers + (size_t)s * NF, NF);
    if (rc) return rc;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    for (int i = 0; i < NF; i++) {
        const double *o = &h[16 * i];
        for (int c = 0; c < 3; c++) { if (P) P[3 * i + c] = o[c]; if (V) V[3 * i + c] = o[7 + c]; if (Ba) Ba[3 * i + c] = o[10 + c]; if (Bg) Bg[3 * i + c] = o[13 + c]; }
        if (Q) for (int c = 0; c < 4; c++) Q[4 * i + c] = o[3 + c];
    }
    return stream_err(be, s);
}

extern "C" int vio_backend_get_post_solve(vio_backend *be, int s, double *out) {
    if (!be || s < 0 || s >= be->s.B || !out) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    int rc = bd2h(be, out, be->s.post_solve + (size_t)s * be->s.NF * 16, (size_t)be->s.NF * 16);
    if (rc) return rc;
    return stream_err(be, s);
}

extern "C" int vio_backend_state_dev(vio_backend *be, const double **state, int64_t *n_doubles) {
    if (!be) return VIO_ERR_ARG;
    if (state) *state = be->s.state_out;
    if (n_doubles) *n_doubles = (int64_t)be->s.B * be->s.NF * 16;
    return VIO_OK;
}

extern "C" int vio_backend_get_info(vio_backend *be, int s, int32_t info[8], double dinfo[4]) {
    if (!be || s < 0 || s >= be->s.B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    int iv[IV_COUNT]; double dv[DV_COUNT];
    int rc = bd2h(be, iv, be->s.iv + (size_t)s * IV_COUNT, IV_COUNT);
    if (!rc) rc = bd2h(be, dv, be->s.dv + (size_t)s * DV_COUNT, DV_COUNT);
    if (rc) return rc;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    if (info) { info[0] = iv[IV_SOLVER_FLAG]; info[1] = iv[IV_MARG_FLAG]; info[2] = iv[IV_FRAME_COUNT]; info[3] = iv[IV_FAILURE];
                info[4] = iv[IV_N_LM]; info[5] = iv[IV_N_FAC]; info[6] = iv[IV_ITERS]; info[7] = iv[IV_LAST_TRACK]; }
    if (dinfo) { dinfo[0] = dv[DV_COST0]; dinfo[1] = dv[DV_COST1]; dinfo[2] = iv[IV_PRIOR_VALID] ? iv[IV_PRIOR_N] : 0;
                 dinfo[3] = iv[IV_ERR] + 16.0 * iv[IV_MARG_FAST] + 32.0 * iv[IV_MARG_SWEEPS] + 4096.0 * iv[IV_MARG_M] + 4194304.0 * iv[IV_CHOL_RETRY]; }
    return VIO_OK;
}

extern "C" int vio_backend_get_features(vio_backend *be, int s, int cap, int *n_out, int32_t *ids, int32_t *start, int32_t *nobs, double *depth,
                                        int32_t *flag) {
    if (!be || s < 0 || s >= be->s.B || !n_out) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    int nf = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&nf, be->s.iv + (size_t)s * IV_COUNT + IV_NFEAT, sizeof(int), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    const int n = std::min(nf, cap);
    const size_t fo = (size_t)s * be->s.FCAP;
    int rc = bd2h(be, ids, be->s.f_id + fo, n);
    if (!rc) rc = bd2h(be, start, be->s.f_start + fo, n);
    if (!rc) rc = bd2h(be, nobs, be->s.f_nobs + fo, n);
    if (!rc) rc = bd2h(be, depth, be->s.f_depth + fo, n);
    if (!rc) rc = bd2h(be, flag, be->s.f_flag + fo, n);
    if (rc) return rc;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    *n_out = nf;
    return nf <= cap ? VIO_OK : VIO_ERR_CAPACITY;
}

// FeaturePerId::feature_per_frame[k].point (x, y; z = 1) of the first min(n_features, cap) features of stream s, in the order of
// vio_backend_get_features: obs [cap][W + 1][2], entry k < n_obs valid.  What a host-side SfM (relativePose / GlobalSFM, VINS.cpp:857-886)
// and FeatureManager::getCorresponding (feature_manager.cpp:157-176) read from f_manager.
extern "C" int vio_backend_get_observations(vio_backend *be, int s, int cap, int *n_out, double *obs) {
    if (!be || s < 0 || s >= be->s.B || !n_out || !obs || cap < 0) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    int nf = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&nf, be->s.iv + (size_t)s * IV_COUNT + IV_NFEAT, sizeof(int), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    const int n = std::min(nf, cap);
    int rc = bd2h(be, obs, be->s.f_obs + (size_t)s * be->s.FCAP * be->s.NF * 2, (size_t)n * be->s.NF * 2);
    if (rc) return rc;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    *n_out = nf;
    return nf <= cap ? VIO_OK : VIO_ERR_CAPACITY;
}

extern "C" int vio_backend_get_prior(vio_backend *be, int s, double *H, double *b, int32_t *present, double *c0) {
    if (!be || s < 0 || s >= be->s.B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t N = be->s.NPX;
    int valid = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&valid, be->s.iv + (size_t)s * IV_COUNT + IV_PRIOR_VALID, sizeof(int), cudaMemcpyDeviceToHost, be->stream));
    int rc = bd2h(be, H, be->s.Hp + (size_t)s * N * N, N * N);
    if (!rc) rc = bd2h(be, b, be->s.bp + (size_t)s * N, N);
    if (!rc) rc = bd2h(be, present, be->s.present + (size_t)s * (2 * be->s.NF + 1), 2 * be->s.NF + 1);
    if (!rc && c0) rc = bd2h(be, c0, be->s.dv + (size_t)s * DV_COUNT + DV_PRIOR_C0, 1);
    if (rc) return rc;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    return valid ? VIO_OK : VIO_ERR_STATE;
}

// whole-batch packed state [batch][W+1][16] -> caller memory (host or device), stream-ordered; host copies synchronise
extern "C" int vio_backend_copy_state(vio_backend *be, double *dst, int dst_is_device) {
    if (!be || !dst) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    const size_t n = (size_t)be->s.B * be->s.NF * 16 * sizeof(double);
    VIO_CUDA_TRY(cudaMemcpyAsync(dst, be->s.state_out, n, dst_is_device == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, be->stream));
    if (dst_is_device == 0) VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));      // 2: pinned host destination, caller synchronises
    return VIO_OK;
}
// diagnostics: per-stream clock64 cycle counters of the solve / marginalisation phases, [batch][32]; reset = 1 zeroes them
extern "C" int vio_backend_phase_cycles(vio_backend *be, long long *out, int reset) {
    if (!be) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    if (out) VIO_CUDA_TRY(cudaMemcpyAsync(out, be->s.prof, (size_t)be->s.B * 32 * sizeof(long long), cudaMemcpyDeviceToHost, be->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    if (reset) VIO_CUDA_TRY(cudaMemsetAsync(be->s.prof, 0, (size_t)be->s.B * 32 * sizeof(long long), be->stream));
    return VIO_OK;
}
extern "C" int64_t vio_backend_launch_count(const vio_backend *be) { return be ? be->launches : 0; }
extern "C" int vio_backend_sync(vio_backend *be) {
    if (!be) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(be->cfg.device));
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    return VIO_OK;
}
extern "C" int vio_backend_profile(vio_backend *be, int enable, char *out, int cap) {
    if (!be) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    const std::string r = be->timer.drain();
    if (out && cap > 0) { strncpy(out, r.c_str(), cap - 1); out[cap - 1] = 0; }
    be->timer.on = enable != 0;
    return VIO_OK;
}
// run the back end on another stream (e.g. the front end's) so that device-to-device hand-over needs no host sync
extern "C" int vio_backend_use_stream(vio_backend *be, void *cuda_stream) {
    if (!be || !cuda_stream) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaStreamSynchronize(be->stream));
    if (be->own_stream) cudaStreamDestroy(be->stream);
    be->stream = (cudaStream_t)cuda_stream; be->own_stream = false;
    return VIO_OK;
}

// ------------------------------------------------------------------ factor-level primitives for the parity tests
__global__ void prim_preint_kernel(int n, const double *dt, const double *acc, const double *gyr, const double *init, const double *noise, double *pr) {
    __shared__ PreScratch scr;
    if (threadIdx.x == 0) pre_init(pr, ld3(init), ld3(init + 3), ld3(init + 6), ld3(init + 9));
    __syncwarp();
    for (int i = 0; i < n; i++) pre_propagate_warp(pr, dt[i], ld3(acc + 3 * i), ld3(gyr + 3 * i), noise, scr);
}

extern "C" int vio_prim_preintegrate(const vio_config *cfg, int n, const double *dt, const double *acc, const double *gyr, const double acc0[3],
                                     const double gyr0[3], const double ba[3], const double bg[3], double *pqv, double *jac, double *cov, double *sum_dt) {
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    double *d;
    const size_t tot = (size_t)n * 7 + 12 + 6 + PR_STRIDE;
    VIO_CUDA_TRY(cudaMalloc((void **)&d, tot * sizeof(double)));
    double *d_dt = d, *d_acc = d + n, *d_gyr = d_acc + 3 * n, *d_init = d_gyr + 3 * n, *d_noise = d_init + 12, *d_pr = d_noise + 6;
    double init[12], noise[6];
    memcpy(init, acc0, 24); memcpy(init + 3, gyr0, 24); memcpy(init + 6, ba, 24); memcpy(init + 9, bg, 24);
    noise[0] = noise[2] = cfg->acc_n * cfg->acc_n; noise[1] = noise[3] = cfg->gyr_n * cfg->gyr_n; noise[4] = cfg->acc_w * cfg->acc_w; noise[5] = cfg->gyr_w * cfg->gyr_w;
    cudaMemcpy(d_dt, dt, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_acc, acc, n * 24, cudaMemcpyHostToDevice); cudaMemcpy(d_gyr, gyr, n * 24, cudaMemcpyHostToDevice);
    cudaMemcpy(d_init, init, sizeof(init), cudaMemcpyHostToDevice); cudaMemcpy(d_noise, noise, sizeof(noise), cudaMemcpyHostToDevice);
    cudaMemset(d_pr, 0, PR_STRIDE * 8);
    prim_preint_kernel<<<1, 32>>>(n, d_dt, d_acc, d_gyr, d_init, d_noise, d_pr);
    std::vector<double> h(PR_STRIDE);
    cudaError_t e = cudaMemcpy(h.data(), d_pr, PR_STRIDE * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return VIO_ERR_CUDA;
    memcpy(pqv, &h[PR_DP], 10 * 8); memcpy(jac, &h[PR_JAC], 225 * 8); memcpy(cov, &h[PR_COV], 225 * 8); *sum_dt = h[PR_SUMDT];
    return VIO_OK;
}

__global__ void prim_imu_factor_kernel(double *pr, double g, const double *x, double *res, double *J, int sqi_given) {
    if (threadIdx.x != 0) return;
    double rr[15], Jr[450];
    if (!sqi_given) imu_sqrt_info(pr + PR_COV, pr + PR_SQI);
    imu_residual(pr, g, x, x + 7, x + 16, x + 23, rr, Jr);
    const double *U = pr + PR_SQI;
    for (int r = 0; r < 15; r++) {
        double t = 0;
        for (int k = r; k < 15; k++) t += U[r * 15 + k] * rr[k];
        res[r] = t;
        for (int c = 0; c < 30; c++) { double u = 0; for (int k = r; k < 15; k++) u += U[r * 15 + k] * Jr[k * 30 + c]; J[r * 30 + c] = u; }
    }
}

extern "C" int vio_prim_imu_factor_sqi(const vio_config *cfg, const double *pqv, const double *jac, const double *cov, double sum_dt,
                                       const double lba[3], const double lbg[3], const double pi[7], const double sbi[9], const double pj[7],
                                       const double sbj[9], const double *sqrt_info_in, double *sqrt_info_out, double *res, double *J) {
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    std::vector<double> h(PR_STRIDE, 0.0);
    memcpy(&h[PR_DP], pqv, 80); memcpy(&h[PR_LBA], lba, 24); memcpy(&h[PR_LBG], lbg, 24); h[PR_SUMDT] = sum_dt; h[PR_VALID] = 1;
    memcpy(&h[PR_JAC], jac, 225 * 8); memcpy(&h[PR_COV], cov, 225 * 8);
    if (sqrt_info_in) memcpy(&h[PR_SQI], sqrt_info_in, 225 * 8);
    double x[32];
    memcpy(x, pi, 56); memcpy(x + 7, sbi, 72); memcpy(x + 16, pj, 56); memcpy(x + 23, sbj, 72);
    double *d;
    VIO_CUDA_TRY(cudaMalloc((void **)&d, (PR_STRIDE + 32 + 15 + 450) * 8));
    cudaMemcpy(d, h.data(), PR_STRIDE * 8, cudaMemcpyHostToDevice); cudaMemcpy(d + PR_STRIDE, x, sizeof(x), cudaMemcpyHostToDevice);
    prim_imu_factor_kernel<<<1, 32>>>(d, cfg->gravity, d + PR_STRIDE, d + PR_STRIDE + 32, d + PR_STRIDE + 47, sqrt_info_in ? 1 : 0);
    cudaMemcpy(res, d + PR_STRIDE + 32, 15 * 8, cudaMemcpyDeviceToHost);
    if (sqrt_info_out) cudaMemcpy(sqrt_info_out, d + PR_SQI, 225 * 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(J, d + PR_STRIDE + 47, 450 * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

extern "C" int vio_prim_imu_factor(const vio_config *cfg, const double *pqv, const double *jac, const double *cov, double sum_dt, const double lba[3],
                                   const double lbg[3], const double pi[7], const double sbi[9], const double pj[7], const double sbj[9], double *res,
                                   double *J) {
    return vio_prim_imu_factor_sqi(cfg, pqv, jac, cov, sum_dt, lba, lbg, pi, sbi, pj, sbj, nullptr, nullptr, res, J);
}

__global__ void prim_proj_kernel(ProjConst K, const double *in, double *out) {
    if (threadIdx.x != 0) return;
    double r2[2], Ji[12], Jj[12], Jl[2], sq;
    proj_eval(K, ld3(in), ld3(in + 3), in + 6, in + 13, in[20], r2, Ji, Jj, Jl, &sq);
    // undo the robust correction so that the raw ProjectionFactor::Evaluate output can be compared: r = r2 / sqrt(rho1)
    const double sr = sqrt(fmax(2.2250738585072014e-308, 1.0 / (1.0 + sq)));
    out[0] = r2[0] / sr; out[1] = r2[1] / sr;
    for (int r = 0; r < 2; r++) {
        for (int c = 0; c < 6; c++) { out[2 + r * 13 + c] = Ji[6 * r + c] / sr; out[2 + r * 13 + 6 + c] = Jj[6 * r + c] / sr; }
        out[2 + r * 13 + 12] = Jl[r] / sr;
    }
    out[28] = 0.5 * log(1.0 + sq);
}

extern "C" int vio_prim_projection_factor(const vio_config *cfg, const double pts_i[3], const double pts_j[3], const double pose_i[7],
                                          const double pose_j[7], double inv_dep, double *res, double *J) {
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    ProjConst K;
    memcpy(K.ric.m, cfg->ric, 72);
    K.tic.x = cfg->tic[0]; K.tic.y = cfg->tic[1]; K.tic.z = cfg->tic[2];
    K.sqrt_info = cfg->fx / 1.5;
    double in[21], out[29];
    memcpy(in, pts_i, 24); memcpy(in + 3, pts_j, 24); memcpy(in + 6, pose_i, 56); memcpy(in + 13, pose_j, 56); in[20] = inv_dep;
    double *d;
    VIO_CUDA_TRY(cudaMalloc((void **)&d, 50 * 8));
    cudaMemcpy(d, in, sizeof(in), cudaMemcpyHostToDevice);
    prim_proj_kernel<<<1, 32>>>(K, d, d + 21);
    cudaError_t e = cudaMemcpy(out, d + 21, sizeof(out), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return VIO_ERR_CUDA;
    memcpy(res, out, 16); memcpy(J, out + 2, 26 * 8);
    return VIO_OK;
}

// ================================================================== visual-inertial alignment (linear half)

extern "C" int vio_visual_imu_align(const vio_config *cfg, int batch, int max_frames, int max_imu, const int32_t *n_frames, const double *R,
                                    const double *T, const int32_t *imu_counts, const double *imu0, const double *imu, const double *bg0,
                                    double *bgs, double *g, double *x, int32_t *ok) {
    if (!cfg || !n_frames || !R || !T || !imu_counts || !imu0 || !imu || !bg0 || !bgs || !g || !x || !ok) return VIO_ERR_ARG;
    if (batch <= 0 || max_frames < 2 || max_frames > 256 || max_imu <= 0) return VIO_ERR_ARG;
    for (int b = 0; b < batch; b++) {
        if (n_frames[b] < 2 || n_frames[b] > max_frames) return VIO_ERR_CAPACITY;
        for (int k = 0; k < n_frames[b]; k++)
            if (imu_counts[(size_t)b * max_frames + k] < 0 || imu_counts[(size_t)b * max_frames + k] > max_imu) return VIO_ERR_CAPACITY;
    }
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    AlignArgs a;
    memset(&a, 0, sizeof(a));
    a.B = batch; a.F = max_frames; a.MAXIMU = max_imu; a.NS = 3 * max_frames + 4;
    const size_t B = batch, F = max_frames, NS = a.NS;
    // one allocation: inputs, scratch, outputs (doubles first, then ints)
    const size_t nd_in = B * F * 9 + B * F * 3 + B * F * 6 + B * F * max_imu * 7 + B * 3;
    const size_t nd_scr = B * F * PR_STRIDE + 2 * B * NS * NS + B * NS + B * F * 110 + 2 * B * NS;
    const size_t nd_out = B * 3 + B * 3 + B * NS;
    const size_t ni = B + B * F + B * NS + B;
    // device scratch: one grow-only buffer per device, kept between calls (cudaMalloc / cudaFree synchronise the whole device and cost
    // milliseconds next to a running pipeline); calls are serialised by the lock
    static std::mutex mu;
    static double *cache[64];
    static size_t cache_bytes[64];
    std::lock_guard<std::mutex> lock(mu);
    const int dev = cfg->device;
    if (dev < 0 || dev >= 64) return VIO_ERR_ARG;
    const size_t need = (nd_in + nd_scr + nd_out) * sizeof(double) + ni * sizeof(int);
    if (cache_bytes[dev] < need) {
        if (cache[dev]) cudaFree(cache[dev]);
        cache[dev] = nullptr; cache_bytes[dev] = 0;
        VIO_CUDA_TRY(cudaMalloc((void **)&cache[dev], need));
        cache_bytes[dev] = need;
    }
    double *d = cache[dev];
    double *p = d;
    double *dR = p; p += B * F * 9;
    double *dT = p; p += B * F * 3;
    double *dI0 = p; p += B * F * 6;
    double *dI = p; p += B * F * max_imu * 7;
    double *dBg = p; p += B * 3;
    a.pre = p; p += B * F * PR_STRIDE;
    a.A = p; p += B * NS * NS;
    a.Aw = p; p += B * NS * NS;
    a.rhs = p; p += B * NS;
    a.pairs = p; p += B * F * 110;
    a.xs = p; p += 2 * B * NS;
    a.bgs_out = p; p += B * 3;
    a.g_out = p; p += B * 3;
    a.x_out = p; p += B * NS;
    int *q = reinterpret_cast<int *>(p);
    int *dN = q; q += B;
    int *dC = q; q += B * F;
    a.perm = q; q += B * NS;
    a.ok = q;
    a.n_frames = dN; a.counts = dC; a.R = dR; a.T = dT; a.imu0 = dI0; a.imu = dI; a.bg0 = dBg;
    a.imu0_stride = 6; a.bg0_stride = 3; a.abg = nullptr; a.abg_stride = 0;
    memcpy(a.tic, cfg->tic, 24);
    a.g_norm = cfg->gravity;             // G_NORM, global_param.hpp:50
    a.g_thr = 3.0;                       // G_THRESHOLD, global_param.hpp:49
    a.noise[0] = a.noise[2] = cfg->acc_n * cfg->acc_n; a.noise[1] = a.noise[3] = cfg->gyr_n * cfg->gyr_n;
    a.noise[4] = cfg->acc_w * cfg->acc_w; a.noise[5] = cfg->gyr_w * cfg->gyr_w;
    cudaMemcpy(dR, R, B * F * 9 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dT, T, B * F * 3 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dI0, imu0, B * F * 6 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dI, imu, B * F * max_imu * 7 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dBg, bg0, B * 3 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dN, n_frames, B * 4, cudaMemcpyHostToDevice); cudaMemcpy(dC, imu_counts, B * F * 4, cudaMemcpyHostToDevice);
    cudaMemset(a.x_out, 0, B * NS * 8); cudaMemset(a.g_out, 0, B * 3 * 8); cudaMemset(a.bgs_out, 0, B * 3 * 8);
    visual_imu_align_kernel<<<batch, ALIGN_THREADS>>>(a);
    cudaMemcpy(bgs, a.bgs_out, B * 3 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(g, a.g_out, B * 3 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(x, a.x_out, B * NS * 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(ok, a.ok, B * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

// ================================================================== motion-only PnP tracker
#include "be_pnp.cuh"

struct vio_pnp {
    vio_config cfg;
    PnpState s;
    cudaStream_t stream;
    std::vector<void *> allocs;
    int64_t launches;
    int *d_counts, *d_ids, *d_track;
    double *d_obs, *d_pos, *d_hdr, *d_imu, *d_init;
    size_t imu_cap;
    size_t smem;
};

template <typename T>
static int palloc(vio_pnp *p, T **ptr, size_t n) {
    VIO_CUDA_TRY(vio_dev_alloc((void **)ptr, n * sizeof(T), p->allocs));
    return VIO_OK;
}
__global__ void pnp_fill_kernel(double *p, size_t n, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void pnp_identity_kernel(double *R, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) R[i] = (i % 9) % 4 == 0 ? 1.0 : 0.0;
}

extern "C" void vio_pnp_destroy(vio_pnp *p) {
    if (!p) return;
    cudaSetDevice(p->cfg.device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void *q : p->allocs) cudaFree(q);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" int vio_pnp_create(const vio_config *cfg, vio_pnp **out) {
    if (!cfg || !out || cfg->batch < 1 || cfg->max_cnt < 1 || cfg->max_cnt > VIO_MAXP) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    vio_pnp *p = new (std::nothrow) vio_pnp();
    if (!p) return VIO_ERR_ARG;
    p->cfg = *cfg; p->launches = 0;
    VIO_CUDA_TRY_OR(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking), vio_pnp_destroy(p));
    PnpState &s = p->s;
    memset(&s, 0, sizeof(s));
    s.B = cfg->batch; s.MAXF = cfg->max_cnt; s.max_iters = 5;                 // options.max_num_iterations = 5 (vins_pnp.cpp:320)
    s.gravity = cfg->gravity; s.sqrt_info = cfg->fx / 1.5;                    // PerspectiveFactor::sqrt_info (vins_pnp.cpp:17-20)
    s.noise[0] = s.noise[2] = cfg->acc_n * cfg->acc_n; s.noise[1] = s.noise[3] = cfg->gyr_n * cfg->gyr_n;
    s.noise[4] = cfg->acc_w * cfg->acc_w; s.noise[5] = cfg->gyr_w * cfg->gyr_w;
    memcpy(s.tic, cfg->tic, sizeof(s.tic)); memcpy(s.ric, cfg->ric, sizeof(s.ric));
    const size_t B = s.B, N = PNP_N, F = s.MAXF;
    int rc = VIO_OK;
    if (!rc) rc = palloc(p, &s.Ps, B * N * 3);
    if (!rc) rc = palloc(p, &s.Rs, B * N * 9);
    if (!rc) rc = palloc(p, &s.Vs, B * N * 3);
    if (!rc) rc = palloc(p, &s.Bas, B * N * 3);
    if (!rc) rc = palloc(p, &s.Bgs, B * N * 3);
    if (!rc) rc = palloc(p, &s.Headers, B * N);
    if (!rc) rc = palloc(p, &s.iv, B * 8);
    if (!rc) rc = palloc(p, &s.solved, B * N);
    if (!rc) rc = palloc(p, &s.dv, B * 8);
    if (!rc) rc = palloc(p, &s.pre, B * N * PR_STRIDE);
    if (!rc) rc = palloc(p, &s.f_n, B * N);
    if (!rc) rc = palloc(p, &s.f_id, B * N * F);
    if (!rc) rc = palloc(p, &s.f_track, B * N * F);
    if (!rc) rc = palloc(p, &s.f_obs, B * N * F * 2);
    if (!rc) rc = palloc(p, &s.f_pos, B * N * F * 3);
    if (!rc) rc = palloc(p, &p->d_counts, B);
    if (!rc) rc = palloc(p, &p->d_ids, B * F);
    if (!rc) rc = palloc(p, &p->d_track, B * F);
    if (!rc) rc = palloc(p, &p->d_obs, B * F * 2);
    if (!rc) rc = palloc(p, &p->d_pos, B * F * 3);
    if (!rc) rc = palloc(p, &p->d_hdr, B);
    if (!rc) rc = palloc(p, &p->d_init, B * 22);
    p->imu_cap = 64;
    if (!rc) rc = palloc(p, &p->d_imu, p->imu_cap * B * 7);
    if (rc) { vio_pnp_destroy(p); return rc; }
    VIO_CUDA_TRY_OR(cudaDeviceSynchronize(), vio_pnp_destroy(p));                                     // zero-fills run on the legacy default stream
    // clearState (vins_pnp.cpp:22-52): identity rotations; Headers start at -1 (the reference leaves them uninitialised)
    pnp_identity_kernel<<<64, 256, 0, p->stream>>>(s.Rs, B * N * 9);
    pnp_fill_kernel<<<64, 256, 0, p->stream>>>(s.Headers, B * N, -1.0);
    p->smem = pnp_smem_doubles() * sizeof(double);
    size_t dyn_max = 0;
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(pnp_image_kernel, cfg->device, &dyn_max), vio_pnp_destroy(p));
    if (p->smem > dyn_max) { vio_pnp_destroy(p); return VIO_ERR_CAPACITY; }
    VIO_CUDA_TRY_OR(cudaStreamSynchronize(p->stream), vio_pnp_destroy(p));
    p->launches += 2;
    *out = p;
    return VIO_OK;
}

// solved_vins of every stream (ViewController.mm:734-739): header[B], P[B][3], R[B][9] row-major, V[B][3], Ba[B][3], Bg[B][3]
extern "C" int vio_pnp_set_init(vio_pnp *p, const double *header, const double *P, const double *R, const double *V, const double *Ba, const double *Bg) {
    if (!p || !header || !P || !R || !V || !Ba || !Bg) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(p->cfg.device));
    const size_t B = p->s.B;
    double *d = p->d_init;
    double *dh = d, *dP = dh + B, *dR = dP + 3 * B, *dV = dR + 9 * B, *dBa = dV + 3 * B, *dBg = dBa + 3 * B;
    VIO_CUDA_TRY(cudaMemcpyAsync(dh, header, B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dP, P, 3 * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dR, R, 9 * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dV, V, 3 * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dBa, Ba, 3 * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dBg, Bg, 3 * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    pnp_set_init_kernel<<<(p->s.B + 127) / 128, 128, 0, p->stream>>>(p->s, dh, dP, dR, dV, dBa, dBg);
    p->launches++;
    VIO_CUDA_TRY(cudaGetLastError());
    VIO_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return VIO_OK;
}

// vinsPnP::processIMU for n consecutive samples: dt[n][B], acc[n][B][3], gyr[n][B][3]
extern "C" int vio_pnp_process_imu(vio_pnp *p, int n, const double *dt, const double *acc, const double *gyr) {
    if (!p || n < 0 || (n > 0 && (!dt || !acc || !gyr))) return VIO_ERR_ARG;
    if (n == 0) return VIO_OK;
    VIO_CUDA_TRY(cudaSetDevice(p->cfg.device));
    const size_t B = p->s.B;
    if ((size_t)n > p->imu_cap) {
        VIO_CUDA_TRY(cudaStreamSynchronize(p->stream));
        double *q;
        VIO_CUDA_TRY(cudaMalloc((void **)&q, (size_t)n * B * 7 * sizeof(double)));
        p->allocs.push_back(q);
        p->d_imu = q; p->imu_cap = n;
    }
    double *d_dt = p->d_imu, *d_acc = d_dt + (size_t)n * B, *d_gyr = d_acc + (size_t)n * B * 3;
    VIO_CUDA_TRY(cudaMemcpyAsync(d_dt, dt, (size_t)n * B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(d_acc, acc, (size_t)n * B * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(d_gyr, gyr, (size_t)n * B * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    pnp_imu_kernel<<<(p->s.B + 3) / 4, 128, 0, p->stream>>>(p->s, n, d_dt, d_acc, d_gyr);
    p->launches++;
    VIO_CUDA_TRY(cudaGetLastError());
    VIO_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return VIO_OK;
}

// vinsPnP::processImage: the landmarks solveVinsPnP matched in this frame (feature_tracker.cpp:119-132), ids ascending per stream:
// counts[B], ids[B][max_cnt], obs_xy[B][max_cnt][2] (normalised image coordinates), pos_xyz[B][max_cnt][3] (world), track_num[B][max_cnt]
extern "C" int vio_pnp_process_image(vio_pnp *p, const int32_t *counts, const int32_t *ids, const double *obs_xy, const double *pos_xyz,
                                     const int32_t *track_num, const double *headers, int use_pnp) {
    if (!p || !counts || !ids || !obs_xy || !pos_xyz || !track_num || !headers) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(p->cfg.device));
    const size_t B = p->s.B, F = p->s.MAXF;
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_counts, counts, B * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_ids, ids, B * F * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_track, track_num, B * F * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_obs, obs_xy, B * F * 2 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_pos, pos_xyz, B * F * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(p->d_hdr, headers, B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    pnp_image_kernel<<<p->s.B, PNP_T, p->smem, p->stream>>>(p->s, p->d_counts, p->d_ids, p->d_obs, p->d_pos, p->d_track, p->d_hdr, use_pnp ? 1 : 0);
    p->launches++;
    VIO_CUDA_TRY(cudaGetLastError());
    VIO_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return VIO_OK;
}

// window of stream s after the last call: P[7][3], R[7][9] row-major, V[7][3], headers[7], find_solved[7]; info = {frame_count, err,
// iterations of the last solve}; cost = {initial, final} of the last solve.  FeatureTracker reads index PNP_SIZE - 1 = 5 (feature_tracker.cpp:156).
extern "C" int vio_pnp_get_state(vio_pnp *p, int s, double *P, double *R, double *V, double *headers, int32_t *find_solved, int32_t info[3], double cost[2]) {
    if (!p || s < 0 || s >= p->s.B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(p->cfg.device));
    const size_t k = (size_t)s * PNP_N;
    int iv[8]; double dv[8];
    if (P) VIO_CUDA_TRY(cudaMemcpyAsync(P, p->s.Ps + 3 * k, PNP_N * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (R) VIO_CUDA_TRY(cudaMemcpyAsync(R, p->s.Rs + 9 * k, PNP_N * 9 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (V) VIO_CUDA_TRY(cudaMemcpyAsync(V, p->s.Vs + 3 * k, PNP_N * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (headers) VIO_CUDA_TRY(cudaMemcpyAsync(headers, p->s.Headers + k, PNP_N * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (find_solved) VIO_CUDA_TRY(cudaMemcpyAsync(find_solved, p->s.solved + k, PNP_N * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(iv, p->s.iv + (size_t)s * 8, sizeof(iv), cudaMemcpyDeviceToHost, p->stream));
    VIO_CUDA_TRY(cudaMemcpyAsync(dv, p->s.dv + (size_t)s * 8, sizeof(dv), cudaMemcpyDeviceToHost, p->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (info) { info[0] = iv[0]; info[1] = iv[2]; info[2] = iv[3]; }
    if (cost) { cost[0] = dv[6]; cost[1] = dv[7]; }
    return VIO_OK;
}
extern "C" int64_t vio_pnp_launch_count(const vio_pnp *p) { return p ? p->launches : 0; }
