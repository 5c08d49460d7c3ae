// frontend.cu -- C-ABI of the batched front end (include/vio_b200.h, vio_frontend_*), replacing
// FeatureTracker::readImage and its helpers (/root/reference/VINS_ios/feature_tracker.cpp:18-321).
// No CPU fallback: every compute entry point launches CUDA kernels on cfg.device.
#include "fe_track.cuh"

#include <new>
#include <vector>
#include <string.h>

using namespace fe;

struct vio_frontend {
    vio_config cfg;
    int B, maxp;
    int lr[4], lc[4];
    size_t lsz[4];
    uint8_t *pyr[2][4];
    CUtensorMap *tmap_dev;    // [2][4] TMA descriptors of the pyramid levels in device memory (lk_kernel), valid where tmap_ok[level]
    int tmap_ok[4];
    int cur;             // index of the cur pyramid; forw = cur ^ 1
    bool has_cur;        // forw_img.empty() == false
    TrackArrays A;
    int img_cnt;
    cudaStream_t stream;
    bool own_stream;
    KernelTimer timer;
    int64_t launches;
    int *err_flag_dev;
    std::vector<void *> allocs;
    // host-image path: uploads go through a double-buffered staging area on their own stream so that the copy of frame k+1 overlaps
    // the kernels of frame k (allocated on first use)
    cudaStream_t up_stream;
    uint8_t *stage[2];
    cudaEvent_t ev_up[2], ev_free[2];
    int up_idx;
    bool stage_used[2];
    // optional CLAHE pre-processing (vio_frontend_set_clahe)
    int clahe_on, clahe_tx, clahe_ty, clahe_clip;
    uint8_t *clahe_lut;
};

template <typename T>
static int dev_alloc(vio_frontend *fe, T **p, size_t n) {
    VIO_CUDA_TRY(vio_dev_alloc((void **)p, n * sizeof(T), fe->allocs));
    return VIO_OK;
}

#ifdef VIO_DEBUG_POISON
// debug library only: scan the guard bands of every live device array; prints and returns the number of corrupted bands
extern "C" int vio_debug_check_guards() {
    cudaDeviceSynchronize();
    int bad = 0;
    std::vector<unsigned char> h(VIO_GUARD);
    for (const VioAlloc &a : vio_guard_registry())
        for (int side = 0; side < 2; side++) {
            const unsigned char *g = side ? (unsigned char *)a.user + a.bytes : (unsigned char *)a.raw;
            if (cudaMemcpy(h.data(), g, VIO_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); continue; }   // freed
            for (size_t i = 0; i < VIO_GUARD; i++)
                if (h[i] != 0xA5) {
                    bad++;
                    fprintf(stderr, "[guard] alloc #%d (%zu bytes) %s band: first bad byte at offset %zd, bytes:", a.idx, a.bytes, side ? "upper" : "lower",
                            side ? (ssize_t)i : (ssize_t)i - (ssize_t)VIO_GUARD);
                    for (size_t j = i; j < i + 16 && j < VIO_GUARD; j++) fprintf(stderr, " %02x", h[j]);
                    fprintf(stderr, "\n");
                    break;
                }
        }
    return bad;
}
#endif

extern "C" void vio_config_default(vio_config *c) {
    memset(c, 0, sizeof(*c));
    c->rows = 640; c->cols = 480;
    c->fx = 526.600; c->fy = 526.678; c->cx = 243.481; c->cy = 315.280;       // global_param.cpp:29-32 (iPhone7P)
    c->tic[0] = 0.0; c->tic[1] = 0.092; c->tic[2] = 0.01;                      // global_param.cpp:37-39
    const double r[9] = {1, 0, 0, 0, -1, 0, 0, 0, -1};                          // ypr2R(0,0,180): global_param.hpp:23-25
    memcpy(c->ric, r, sizeof(r));
    c->max_cnt = 150; c->min_dist = 30; c->f_threshold = 1.0;
    c->freq = 3; c->window_size = 10; c->num_of_f = 1000;
    c->acc_n = 0.5; c->acc_w = 0.002; c->gyr_n = 0.2; c->gyr_w = 4.0e-5; c->gravity = 9.805;
    c->max_iters = 10;
    c->min_parallax = 10.0 / 549.0; c->init_depth = 5.0;
    c->max_imu_per_frame = 256; c->batch = 1; c->device = 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static int encode_level_map(CUtensorMap *out, const uint8_t *base, int rows, int cols, int batch) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return VIO_ERR_CUDA;
        fn = (encode_fn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)cols, (cuuint64_t)rows * cols};          // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {LK_BOX, 32, 1}, estr[3] = {1, 1, 1};     // see LkMaps
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? VIO_OK : VIO_ERR_CUDA;
}

extern "C" int vio_frontend_create(const vio_config *cfg, vio_frontend **out) {
    if (!cfg || !out || cfg->batch < 1 || cfg->max_cnt < 1 || cfg->max_cnt > VIO_MAXP || cfg->rows < 64 || cfg->cols < 64 ||
        cfg->min_dist < 1 || cfg->min_dist > 127 || cfg->freq < 1)
        return VIO_ERR_ARG;
    const int gw = (cfg->cols + cfg->min_dist - 1) / cfg->min_dist, gh = (cfg->rows + cfg->min_dist - 1) / cfg->min_dist;
    if (gw * gh > 2048) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    vio_poison_load_mask();
    vio_frontend *fe = new (std::nothrow) vio_frontend();
    if (!fe) return VIO_ERR_ARG;
    fe->cfg = *cfg; fe->B = cfg->batch; fe->maxp = cfg->max_cnt;
    fe->cur = 0; fe->has_cur = false; fe->img_cnt = 0; fe->launches = 0; fe->own_stream = true;
    VIO_CUDA_TRY_OR(cudaStreamCreateWithFlags(&fe->stream, cudaStreamNonBlocking), vio_frontend_destroy(fe));
    int r = cfg->rows, c = cfg->cols;
    for (int l = 0; l < 4; l++) {
        fe->lr[l] = r; fe->lc[l] = c; fe->lsz[l] = (size_t)r * c;
        r = (r + 1) / 2; c = (c + 1) / 2;
    }
    const size_t B = fe->B, P = fe->maxp;
    int rc = VIO_OK;
    for (int k = 0; k < 2 && !rc; k++)
        for (int l = 0; l < 4 && !rc; l++) rc = dev_alloc(fe, &fe->pyr[k][l], B * fe->lsz[l]);
    if (!rc) rc = dev_alloc(fe, &fe->tmap_dev, 8);
    if (!rc) {
        alignas(64) CUtensorMap host_maps[8];
        memset(host_maps, 0, sizeof(host_maps));
        for (int l = 0; l < 4 && !rc; l++) {
            fe->tmap_ok[l] = (fe->lc[l] % 16 == 0 && fe->lc[l] >= LK_BOX && fe->lr[l] >= 32) ? 1 : 0;
            for (int k = 0; k < 2 && !rc && fe->tmap_ok[l]; k++) rc = encode_level_map(&host_maps[4 * k + l], fe->pyr[k][l], fe->lr[l], fe->lc[l], fe->B);
        }
        if (!rc && cudaMemcpy(fe->tmap_dev, host_maps, sizeof(host_maps), cudaMemcpyHostToDevice) != cudaSuccess) rc = VIO_ERR_CUDA;
    }
    TrackArrays &A = fe->A;
    if (!rc) rc = dev_alloc(fe, &A.cur_pts, B * P);
    if (!rc) rc = dev_alloc(fe, &A.pre_pts, B * P);
    if (!rc) rc = dev_alloc(fe, &A.forw_pts, B * P);
    if (!rc) rc = dev_alloc(fe, &A.pmin, B * P);
    if (!rc) rc = dev_alloc(fe, &A.pmax, B * P);
    if (!rc) rc = dev_alloc(fe, &A.ids, B * P);
    if (!rc) rc = dev_alloc(fe, &A.track_cnt, B * P);
    if (!rc) rc = dev_alloc(fe, &A.n, B);
    if (!rc) rc = dev_alloc(fe, &A.status, B * P);
    if (!rc) rc = dev_alloc(fe, &A.kept, B * P);
    if (!rc) rc = dev_alloc(fe, &A.n_kept, B);
    if (!rc) rc = dev_alloc(fe, &A.good_pts, 2 * B * P);
    if (!rc) rc = dev_alloc(fe, &A.track_len, 2 * B * P);
    if (!rc) rc = dev_alloc(fe, &A.n_good, B);
    if (!rc) rc = dev_alloc(fe, &A.stats, B * 8);
    if (!rc) rc = dev_alloc(fe, &A.n_id, B);
    if (!rc) rc = dev_alloc(fe, &A.msg_cnt, B);
    if (!rc) rc = dev_alloc(fe, &A.msg_ids, B * P);
    if (!rc) rc = dev_alloc(fe, &A.msg_xyz, B * P * 3);
    if (!rc) rc = dev_alloc(fe, &A.max_bits, B);
    if (!rc) rc = dev_alloc(fe, &A.cand, B * (size_t)CAND_CAP);
    if (!rc) rc = dev_alloc(fe, &A.cand_cnt, B);
    if (!rc) rc = dev_alloc(fe, &fe->err_flag_dev, 1);
    A.err_flag = fe->err_flag_dev;
    if (rc) { vio_frontend_destroy(fe); return rc; }
    VIO_CUDA_TRY_OR(cudaDeviceSynchronize(), vio_frontend_destroy(fe));       // dev_alloc() zero-fills on the legacy default stream; fe->stream is non-blocking
    size_t dyn_max = 0;
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(post_track_kernel, cfg->device, &dyn_max), vio_frontend_destroy(fe));
    if (sizeof(TrackSmem) > dyn_max) { vio_frontend_destroy(fe); return VIO_ERR_CAPACITY; }
    VIO_CUDA_TRY_OR(vio_allow_max_dynamic_smem(select_kernel, cfg->device, &dyn_max), vio_frontend_destroy(fe));
    if (sizeof(SelectSmem) > dyn_max) { vio_frontend_destroy(fe); return VIO_ERR_CAPACITY; }
    *out = fe;
    return VIO_OK;
}

extern "C" void vio_frontend_destroy(vio_frontend *fe) {
    if (!fe) return;
    cudaSetDevice(fe->cfg.device);
    if (fe->stream) cudaStreamSynchronize(fe->stream);
    if (fe->up_stream) {
        cudaStreamSynchronize(fe->up_stream);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(fe->ev_up[i]); cudaEventDestroy(fe->ev_free[i]); }
        cudaStreamDestroy(fe->up_stream);
    }
    for (void *p : fe->allocs) cudaFree(p);
    if (fe->own_stream && fe->stream) cudaStreamDestroy(fe->stream);
    delete fe;
}

// TMA descriptors of pyramid k (template side) and pyramid kj (search side) for lk_kernel
static LkMaps maps_of(const vio_frontend *fe, int k, int kj) {
    LkMaps m;
    m.I = fe->tmap_dev + 4 * k; m.J = fe->tmap_dev + 4 * kj;
    for (int l = 0; l < 4; l++) m.ok[l] = fe->tmap_ok[l];
    return m;
}

static PyrLevels levels_of(const vio_frontend *fe, int k) {
    PyrLevels L;
    for (int l = 0; l < 4; l++) { L.p[l] = fe->pyr[k][l]; L.rows[l] = fe->lr[l]; L.cols[l] = fe->lc[l]; L.stride[l] = fe->lsz[l]; }
    return L;
}

// CLAHE of the B images at `img` (in place): two launches
static void run_clahe(vio_frontend *fe, uint8_t *img, cudaStream_t s) {
    const int rows = fe->cfg.rows, cols = fe->cfg.cols, tw = cols / fe->clahe_tx, th = rows / fe->clahe_ty;
    const float lut_scale = 255.f / (float)(tw * th);
    VIO_LAUNCH(fe->timer, s, "clahe_lut_kernel", (clahe_lut_kernel<<<dim3(fe->clahe_tx * fe->clahe_ty, fe->B), 256, 0, s>>>(img, fe->lsz[0], cols, tw, th, fe->clahe_tx,
               fe->clahe_clip, lut_scale, fe->clahe_lut)));
    VIO_LAUNCH(fe->timer, s, "clahe_apply_kernel", (clahe_apply_kernel<<<dim3((cols + 1023) / 1024, rows, fe->B), 256, 0, s>>>(img, fe->lsz[0], rows, cols, tw, th,
               fe->clahe_tx, fe->clahe_ty, fe->clahe_lut)));
    fe->launches += 2;
}

extern "C" int vio_frontend_set_clahe(vio_frontend *fe, int enable, double clip_limit, int tiles_x, int tiles_y) {
    if (!fe) return VIO_ERR_ARG;
    if (!enable) { fe->clahe_on = 0; return VIO_OK; }
    if (tiles_x < 1 || tiles_y < 1 || tiles_x * tiles_y > 4096 || fe->cfg.cols % tiles_x || fe->cfg.rows % tiles_y) return VIO_ERR_ARG;   // OpenCV pads otherwise
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    if (!fe->clahe_lut || tiles_x * tiles_y > fe->clahe_tx * fe->clahe_ty) {
        void *p = nullptr;
        VIO_CUDA_TRY(vio_dev_alloc(&p, (size_t)fe->B * tiles_x * tiles_y * 256, fe->allocs));
        VIO_CUDA_TRY(cudaDeviceSynchronize());
        fe->clahe_lut = (uint8_t *)p;
    }
    const int area = (fe->cfg.cols / tiles_x) * (fe->cfg.rows / tiles_y);
    fe->clahe_tx = tiles_x; fe->clahe_ty = tiles_y;
    fe->clahe_clip = clip_limit > 0.0 ? std::max((int)(clip_limit * area / 256), 1) : 0;      // clahe.cpp: clipLimit = max(int(clip * tileArea / histSize), 1)
    fe->clahe_on = 1;
    return VIO_OK;
}

// The body of readImage once the new frame sits in pyr[forw][0].
static int run_frame(vio_frontend *fe, int *published) {
    const int forw = fe->cur ^ 1, B = fe->B;
    cudaStream_t s = fe->stream;
    if (fe->clahe_on) run_clahe(fe, fe->pyr[forw][0], s);
    for (int l = 0; l < 3; l++) {                                           // K1
        dim3 blk(32, 8), grd((fe->lc[l + 1] + 127) / 128, (fe->lr[l + 1] + 7) / 8, B);
        VIO_LAUNCH(fe->timer, s, "pyr_down_kernel", (pyr_down_kernel<<<grd, blk, 0, s>>>(fe->pyr[forw][l], fe->pyr[forw][l + 1], fe->lr[l], fe->lc[l],
                   fe->lr[l + 1], fe->lc[l + 1], fe->lsz[l], fe->lsz[l + 1])));
        fe->launches++;
    }
    const int detect = fe->img_cnt == 0;
    if (fe->has_cur) {                                                      // K4
        dim3 grd((fe->maxp + LK_WARPS - 1) / LK_WARPS, B);
        VIO_LAUNCH(fe->timer, s, "lk_kernel", (lk_kernel<<<grd, LK_WARPS * 32, 0, s>>>(levels_of(fe, fe->cur), levels_of(fe, forw), maps_of(fe, fe->cur, forw), fe->A.cur_pts,
                   fe->A.forw_pts, fe->A.status, fe->A.n, fe->maxp)));
        fe->launches++;
    }
    VIO_LAUNCH(fe->timer, s, "post_track_kernel", (post_track_kernel<<<B, 256, sizeof(TrackSmem), s>>>(fe->A, fe->maxp, fe->cfg.rows, fe->cfg.cols,
               fe->cfg.min_dist, fe->cfg.f_threshold, detect, fe->has_cur ? 1 : 0)));
    fe->launches++;
    if (detect) {
        dim3 grd(((fe->cfg.cols + EG_W - 1) / EG_W + EG_WARPS - 1) / EG_WARPS, (fe->cfg.rows + EG_ROWS - 1) / EG_ROWS, B);
        VIO_LAUNCH(fe->timer, s, "eig_candidates_kernel", (eig_candidates_kernel<<<grd, EG_WARPS * 32, 0, s>>>(fe->pyr[forw][0], fe->lsz[0], fe->cfg.rows,
                   fe->cfg.cols, fe->A.kept, fe->A.n_kept, fe->maxp, fe->cfg.min_dist, fe->A.max_bits, fe->A.cand, fe->A.cand_cnt)));
        VIO_LAUNCH(fe->timer, s, "select_kernel", (select_kernel<<<B, 256, sizeof(SelectSmem), s>>>(fe->A, fe->maxp, fe->cfg.rows, fe->cfg.cols,
                   fe->cfg.max_cnt, fe->cfg.min_dist, fe->cfg.fx, fe->cfg.fy, fe->cfg.cx, fe->cfg.cy)));
        fe->launches += 2;
    }
    VIO_CUDA_TRY(cudaGetLastError());
    fe->cur = forw;                 // cur_img = forw_img
    fe->has_cur = true;
    if (published) *published = detect;
    fe->img_cnt = (fe->img_cnt + 1) % fe->cfg.freq;                         // ViewController.mm:494
    return VIO_OK;
}

extern "C" uint8_t *vio_frontend_next_image_buffer(vio_frontend *fe) { return fe ? fe->pyr[fe->cur ^ 1][0] : nullptr; }

extern "C" int vio_frontend_read_images(vio_frontend *fe, const uint8_t *images_host, int *published) {
    if (!fe || !images_host) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    const size_t bytes = (size_t)fe->B * fe->lsz[0];
    if (!fe->up_stream) {
        VIO_CUDA_TRY(cudaStreamCreateWithFlags(&fe->up_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            VIO_CUDA_TRY(cudaMalloc((void **)&fe->stage[i], bytes));
            fe->allocs.push_back(fe->stage[i]);
            VIO_CUDA_TRY(cudaEventCreateWithFlags(&fe->ev_up[i], cudaEventDisableTiming));
            VIO_CUDA_TRY(cudaEventCreateWithFlags(&fe->ev_free[i], cudaEventDisableTiming));
        }
    }
    const int i = fe->up_idx;
    fe->up_idx ^= 1;
    if (fe->stage_used[i]) VIO_CUDA_TRY(cudaStreamWaitEvent(fe->up_stream, fe->ev_free[i], 0));
    VIO_CUDA_TRY(cudaMemcpyAsync(fe->stage[i], images_host, bytes, cudaMemcpyHostToDevice, fe->up_stream));
    VIO_CUDA_TRY(cudaEventRecord(fe->ev_up[i], fe->up_stream));
    VIO_CUDA_TRY(cudaStreamWaitEvent(fe->stream, fe->ev_up[i], 0));
    VIO_CUDA_TRY(cudaMemcpyAsync(fe->pyr[fe->cur ^ 1][0], fe->stage[i], bytes, cudaMemcpyDeviceToDevice, fe->stream));
    VIO_CUDA_TRY(cudaEventRecord(fe->ev_free[i], fe->stream));
    fe->stage_used[i] = true;
    return run_frame(fe, published);
}

extern "C" int vio_frontend_read_images_dev(vio_frontend *fe, const uint8_t *images_dev, int *published) {
    if (!fe || !images_dev) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    uint8_t *dst = fe->pyr[fe->cur ^ 1][0];
    if (images_dev != dst)
        VIO_CUDA_TRY(cudaMemcpyAsync(dst, images_dev, (size_t)fe->B * fe->lsz[0], cudaMemcpyDeviceToDevice, fe->stream));
    return run_frame(fe, published);
}

static int check_err(vio_frontend *fe) {
    int e = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&e, fe->err_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, fe->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    return e;
}

template <typename T>
static int d2h(vio_frontend *fe, T *dst, const T *src, size_t n) {
    if (!dst || n == 0) return VIO_OK;
    VIO_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, fe->stream));
    return VIO_OK;
}

extern "C" int vio_frontend_get_stream(vio_frontend *fe, int s, int *n_out, int32_t *ids, float *pts_xy, int32_t *track_cnt, double *norm_xyz) {
    if (!fe || s < 0 || s >= fe->B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    int n = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&n, fe->A.n + s, sizeof(int), cudaMemcpyDeviceToHost, fe->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    const size_t o = (size_t)s * fe->maxp;
    int rc = d2h(fe, ids, fe->A.ids + o, n);
    if (!rc) rc = d2h(fe, (float2 *)pts_xy, fe->A.cur_pts + o, n);
    if (!rc) rc = d2h(fe, track_cnt, fe->A.track_cnt + o, n);
    if (!rc) rc = d2h(fe, norm_xyz, fe->A.msg_xyz + 3 * o, (size_t)3 * n);
    if (rc) return rc;
    if (n_out) *n_out = n;
    return check_err(fe);
}

extern "C" int vio_frontend_get_ui(vio_frontend *fe, int s, int *n_out, float *good_pts_xy, double *track_len) {
    if (!fe || s < 0 || s >= fe->B) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    int n = 0;
    VIO_CUDA_TRY(cudaMemcpyAsync(&n, fe->A.n_good + s, sizeof(int), cudaMemcpyDeviceToHost, fe->stream));
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    const size_t o = (size_t)s * 2 * fe->maxp;
    int rc = d2h(fe, (float2 *)good_pts_xy, fe->A.good_pts + o, n);
    if (!rc) rc = d2h(fe, track_len, fe->A.track_len + o, n);
    if (rc) return rc;
    if (n_out) *n_out = n;
    return check_err(fe);
}

extern "C" int vio_frontend_get_stats(vio_frontend *fe, int s, int32_t stats[8]) {
    if (!fe || s < 0 || s >= fe->B || !stats) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    int rc = d2h(fe, stats, fe->A.stats + s * 8, 8);
    if (rc) return rc;
    return check_err(fe);
}

extern "C" int vio_frontend_image_msg_dev(vio_frontend *fe, const int32_t **counts, const int32_t **ids, const double **norm_xyz) {
    if (!fe) return VIO_ERR_ARG;
    if (counts) *counts = fe->A.msg_cnt;
    if (ids) *ids = fe->A.msg_ids;
    if (norm_xyz) *norm_xyz = fe->A.msg_xyz;
    return VIO_OK;
}

extern "C" int64_t vio_frontend_launch_count(const vio_frontend *fe) { return fe ? fe->launches : 0; }

// cudaStream handle for callers that chain the back end on the same stream (C-ABI keeps it opaque)
extern "C" void *vio_frontend_stream(vio_frontend *fe) { return fe ? (void *)fe->stream : nullptr; }
extern "C" int vio_frontend_use_stream(vio_frontend *fe, void *cuda_stream) {
    if (!fe) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    if (fe->own_stream) cudaStreamDestroy(fe->stream);
    fe->stream = (cudaStream_t)cuda_stream; fe->own_stream = false;
    return VIO_OK;
}
extern "C" int vio_frontend_profile(vio_frontend *fe, int enable, char *out, int cap) {
    if (!fe) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    const std::string r = fe->timer.drain();
    if (out && cap > 0) { strncpy(out, r.c_str(), cap - 1); out[cap - 1] = 0; }
    fe->timer.on = enable != 0;
    return VIO_OK;
}
extern "C" int vio_frontend_sync(vio_frontend *fe) {
    if (!fe) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(fe->cfg.device));
    VIO_CUDA_TRY(cudaStreamSynchronize(fe->stream));
    return VIO_OK;
}

// ------------------------------------------------------------------ primitives for the parity tests (batch 1)
static int make_single(const vio_config *cfg, vio_frontend **fe) {
    vio_config c = *cfg;
    c.batch = 1;
    return vio_frontend_create(&c, fe);
}

extern "C" int vio_prim_pyramid(const vio_config *cfg, const uint8_t *img, uint8_t *l1, uint8_t *l2, uint8_t *l3) {
    vio_frontend *fe = nullptr;
    int rc = make_single(cfg, &fe);
    if (rc) return rc;
    cudaMemcpyAsync(fe->pyr[0][0], img, fe->lsz[0], cudaMemcpyHostToDevice, fe->stream);
    for (int l = 0; l < 3; l++) {
        dim3 blk(32, 8), grd((fe->lc[l + 1] + 127) / 128, (fe->lr[l + 1] + 7) / 8, 1);
        pyr_down_kernel<<<grd, blk, 0, fe->stream>>>(fe->pyr[0][l], fe->pyr[0][l + 1], fe->lr[l], fe->lc[l], fe->lr[l + 1], fe->lc[l + 1],
                                                     fe->lsz[l], fe->lsz[l + 1]);
    }
    uint8_t *outs[3] = {l1, l2, l3};
    for (int l = 0; l < 3; l++) cudaMemcpyAsync(outs[l], fe->pyr[0][l + 1], fe->lsz[l + 1], cudaMemcpyDeviceToHost, fe->stream);
    cudaError_t e = cudaStreamSynchronize(fe->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    vio_frontend_destroy(fe);
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

extern "C" int vio_prim_clahe(const vio_config *cfg, const uint8_t *img, double clip_limit, int tiles_x, int tiles_y, uint8_t *out) {
    vio_frontend *fe = nullptr;
    int rc = make_single(cfg, &fe);
    if (rc) return rc;
    rc = vio_frontend_set_clahe(fe, 1, clip_limit, tiles_x, tiles_y);
    if (rc) { vio_frontend_destroy(fe); return rc; }
    cudaMemcpyAsync(fe->pyr[0][0], img, fe->lsz[0], cudaMemcpyHostToDevice, fe->stream);
    run_clahe(fe, fe->pyr[0][0], fe->stream);
    cudaMemcpyAsync(out, fe->pyr[0][0], fe->lsz[0], cudaMemcpyDeviceToHost, fe->stream);
    cudaError_t e = cudaStreamSynchronize(fe->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    vio_frontend_destroy(fe);
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

extern "C" int vio_prim_lk(const vio_config *cfg, const uint8_t *prev, const uint8_t *next, const float *pts_xy, int n, float *next_xy,
                           uint8_t *status) {
    if (n > cfg->max_cnt) return VIO_ERR_ARG;
    vio_frontend *fe = nullptr;
    int rc = make_single(cfg, &fe);
    if (rc) return rc;
    cudaStream_t s = fe->stream;
    cudaMemcpyAsync(fe->pyr[0][0], prev, fe->lsz[0], cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->pyr[1][0], next, fe->lsz[0], cudaMemcpyHostToDevice, s);
    for (int k = 0; k < 2; k++)
        for (int l = 0; l < 3; l++) {
            dim3 blk(32, 8), grd((fe->lc[l + 1] + 127) / 128, (fe->lr[l + 1] + 7) / 8, 1);
            pyr_down_kernel<<<grd, blk, 0, s>>>(fe->pyr[k][l], fe->pyr[k][l + 1], fe->lr[l], fe->lc[l], fe->lr[l + 1], fe->lc[l + 1], fe->lsz[l],
                                                fe->lsz[l + 1]);
        }
    cudaMemcpyAsync(fe->A.cur_pts, pts_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->A.n, &n, sizeof(int), cudaMemcpyHostToDevice, s);
    dim3 grd((fe->maxp + LK_WARPS - 1) / LK_WARPS, 1);
    lk_kernel<<<grd, LK_WARPS * 32, 0, s>>>(levels_of(fe, 0), levels_of(fe, 1), maps_of(fe, 0, 1), fe->A.cur_pts, fe->A.forw_pts, fe->A.status, fe->A.n, fe->maxp);
    cudaMemcpyAsync(next_xy, fe->A.forw_pts, sizeof(float2) * n, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(status, fe->A.status, n, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    vio_frontend_destroy(fe);
    return e == cudaSuccess ? VIO_OK : VIO_ERR_CUDA;
}

// goodFeaturesToTrack(img, max_corners, 0.01, min_dist, mask) with mask = complement of discs around kept_xy (rounded)
extern "C" int vio_prim_min_eig_candidates(const vio_config *cfg, const uint8_t *img, const float *kept_xy, int n_kept, int max_corners,
                                           float *corners_xy, int *n_corners, float *max_val) {
    if (n_kept + max_corners > cfg->max_cnt) return VIO_ERR_ARG;
    vio_frontend *fe = nullptr;
    vio_config c = *cfg;
    c.max_cnt = n_kept + max_corners;
    int rc = make_single(&c, &fe);
    if (rc) return rc;
    cudaStream_t s = fe->stream;
    std::vector<int2> k(n_kept > 0 ? n_kept : 1);
    std::vector<float2> kp(n_kept > 0 ? n_kept : 1);
    for (int i = 0; i < n_kept; i++) {
        k[i] = make_int2((int)rintf(kept_xy[2 * i]), (int)rintf(kept_xy[2 * i + 1]));
        kp[i] = make_float2(kept_xy[2 * i], kept_xy[2 * i + 1]);
    }
    cudaMemcpyAsync(fe->pyr[0][0], img, fe->lsz[0], cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->A.kept, k.data(), sizeof(int2) * n_kept, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->A.forw_pts, kp.data(), sizeof(float2) * n_kept, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->A.n_kept, &n_kept, sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(fe->A.n, &n_kept, sizeof(int), cudaMemcpyHostToDevice, s);
    dim3 grd(((c.cols + EG_W - 1) / EG_W + EG_WARPS - 1) / EG_WARPS, (c.rows + EG_ROWS - 1) / EG_ROWS, 1);
    eig_candidates_kernel<<<grd, EG_WARPS * 32, 0, s>>>(fe->pyr[0][0], fe->lsz[0], c.rows, c.cols, fe->A.kept, fe->A.n_kept, fe->maxp, c.min_dist,
                                              fe->A.max_bits, fe->A.cand, fe->A.cand_cnt);
    select_kernel<<<1, 256, sizeof(SelectSmem), s>>>(fe->A, fe->maxp, c.rows, c.cols, c.max_cnt, c.min_dist, c.fx, c.fy, c.cx, c.cy);
    int n = 0;
    unsigned mb = 0;
    cudaMemcpyAsync(&n, fe->A.n, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&mb, fe->A.max_bits, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && n > n_kept) e = cudaMemcpy(corners_xy, fe->A.forw_pts + n_kept, sizeof(float2) * (n - n_kept), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaGetLastError();
    *n_corners = n - n_kept;
    if (max_val) memcpy(max_val, &mb, 4);
    int err = check_err(fe);
    vio_frontend_destroy(fe);
    if (e != cudaSuccess) return VIO_ERR_CUDA;
    return err;
}

__global__ void __launch_bounds__(256) ransac_test_kernel(const float2 *p1, const float2 *p2, int n, double thresh, uint8_t *mask, int *ok_out,
                                                          int *iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrackSmem &s = *reinterpret_cast<TrackSmem *>(smem_raw);
    for (int i = threadIdx.x; i < n; i += 256) { s.cur[i] = p1[i]; s.forw[i] = p2[i]; s.flag[i] = 1; }
    __syncthreads();
    const bool ok = find_fundamental_cta(s.cur, s.forw, n, thresh, 0.99, s.flag, s.rs, iters);
    for (int i = threadIdx.x; i < n; i += 256) mask[i] = s.flag[i];
    if (threadIdx.x == 0) *ok_out = ok ? 1 : 0;
}

extern "C" int vio_prim_ransac_f(const vio_config *cfg, const float *p1_xy, const float *p2_xy, int n, uint8_t *mask, int *iters) {
    if (n < 8 || n > VIO_MAXP) return VIO_ERR_ARG;
    VIO_CUDA_TRY(cudaSetDevice(cfg->device));
    float2 *d1, *d2; uint8_t *dm; int *dok;
    VIO_CUDA_TRY(cudaMalloc(&d1, sizeof(float2) * n));
    VIO_CUDA_TRY(cudaMalloc(&d2, sizeof(float2) * n));
    VIO_CUDA_TRY(cudaMalloc(&dm, n));
    VIO_CUDA_TRY(cudaMalloc(&dok, 2 * sizeof(int)));
    VIO_CUDA_TRY(cudaMemcpy(d1, p1_xy, sizeof(float2) * n, cudaMemcpyHostToDevice));
    VIO_CUDA_TRY(cudaMemcpy(d2, p2_xy, sizeof(float2) * n, cudaMemcpyHostToDevice));
    VIO_CUDA_TRY(vio_allow_max_dynamic_smem(ransac_test_kernel, cfg->device, nullptr));
    ransac_test_kernel<<<1, 256, sizeof(TrackSmem)>>>(d1, d2, n, cfg->f_threshold, dm, dok, dok + 1);
    int h[2] = {0, 0};
    VIO_CUDA_TRY(cudaMemcpy(h, dok, sizeof(h), cudaMemcpyDeviceToHost));
    VIO_CUDA_TRY(cudaMemcpy(mask, dm, n, cudaMemcpyDeviceToHost));
    if (iters) *iters = h[1];
    cudaFree(d1); cudaFree(d2); cudaFree(dm); cudaFree(dok);
    return h[0] ? VIO_OK : VIO_ERR_STATE;     // VIO_ERR_STATE: no model (OpenCV returns an empty matrix / mask)
}
