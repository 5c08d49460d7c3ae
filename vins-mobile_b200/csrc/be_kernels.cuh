// be_kernels.cuh -- estimator-loop kernels around the solve (one CTA per stream unless noted):
//   imu_kernel        VINS::processIMU                     VINS.cpp:333-375  (+ IntegrationBase::push_back)
//   addfeat_kernel    FeatureManager::addFeatureCheckParallax feature_manager.cpp:103-155, control flow of VINS::processImage VINS.cpp:377-478
//   triangulate_kernel FeatureManager::triangulate         feature_manager.cpp:190-257 (+ caller-supplied initial window)
//   prepare_kernel    VINS::old2new + problem enumeration  VINS.cpp:89-129, 516-567
//   post_solve_kernel VINS::new2old + setDepth             VINS.cpp:131-212, feature_manager.cpp:331-349
//   finish_kernel     failureDetection, slideWindow*, removeBack*/removeFront/removeFailures
//                                                           VINS.cpp:214-265,1149-1273; feature_manager.cpp:259-298,356-406
#pragma once
#include "be_state.cuh"

namespace be {

// block-wide exclusive scan of v (one value per thread), returns the exclusive prefix; *total gets the block sum (valid in all threads)
__device__ inline int block_excl_scan(int v, int *smem_warp /*>=33 ints*/, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) smem_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? smem_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < nw) smem_warp[lane] = w;
        if (lane == nw - 1) smem_warp[32] = w;
    }
    __syncthreads();
    const int base = warp == 0 ? 0 : smem_warp[warp - 1];
    *total = smem_warp[32];
    return base + x - v;
}

__device__ inline double block_sum_d(double v, double *smem /*>=32*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0;
    for (int i = 0; i < nw; i++) r += smem[i];      // fixed order: deterministic
    return r;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) imu_kernel(BeState s, int n_samples, const double *__restrict__ dts, const double *__restrict__ accs,
                                                  const double *__restrict__ gyrs) {
    VIO_POISON(1u);
    __shared__ PreScratch scr[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= s.B) return;
    int *iv = S_iv(s, b);
    double *dv = S_dv(s, b);
    for (int n = 0; n < n_samples; n++) {
        const double dt = dts[(size_t)n * s.B + b];
        const V3 acc = ld3(accs + ((size_t)n * s.B + b) * 3), gyr = ld3(gyrs + ((size_t)n * s.B + b) * 3);
        const int fc = iv[IV_FRAME_COUNT];
        if (lane == 0 && !iv[IV_FIRST_IMU]) { iv[IV_FIRST_IMU] = 1; st3(dv + DV_ACC0, acc); st3(dv + DV_GYR0, gyr); }
        __syncwarp();
        double *pr = S_pre(s, b, fc);
        if (lane == 0 && pr[PR_VALID] == 0.0) pre_init(pr, ld3(dv + DV_ACC0), ld3(dv + DV_GYR0), ld3(S_Bas(s, b, fc)), ld3(S_Bgs(s, b, fc)));
        __syncwarp();
        if (fc != 0) {
            pre_propagate_warp(pr, dt, acc, gyr, s.noise, scr[warp]);
            if (lane == 0) {
                int &cnt = s.imu_cnt[(size_t)b * s.NF + fc];
                if (cnt < s.MAXIMU) {
                    double *e = S_imu(s, b, fc) + 7 * cnt;
                    e[0] = dt; st3(e + 1, acc); st3(e + 4, gyr);
                    cnt++;
                } else iv[IV_ERR] = VIO_ERR_CAPACITY;
                if (iv[IV_SOLVER_FLAG] == 0) {               // tmp_pre_integration->push_back (VINS.cpp:352-353): the open all_image_frame record
                    const int k = iv[IV_AF_N];
                    if (k < s.FA) {
                        int &ac = s.af_cnt[(size_t)b * s.FA + k];
                        if (ac < s.MAXIMU) {
                            double *e = s.af_imu + (((size_t)b * s.FA + k) * s.MAXIMU + ac) * 7;
                            e[0] = dt; st3(e + 1, acc); st3(e + 4, gyr);
                            ac++;
                        }
                    }
                }
                // mid-point propagation of the newest state (VINS.cpp:359-370)
                const V3 g = v3(0, 0, s.gravity);
                const V3 a0 = ld3(dv + DV_ACC0), g0 = ld3(dv + DV_GYR0), ba = ld3(S_Bas(s, b, fc)), bg = ld3(S_Bgs(s, b, fc));
                M3 R = ldm(S_Rs(s, b, fc));
                V3 P = ld3(S_Ps(s, b, fc)), V = ld3(S_Vs(s, b, fc));
                const V3 un_acc_0 = R * (a0 - ba) - g;
                const V3 un_gyr = 0.5 * (g0 + gyr) - bg;
                R = R * q2R(deltaQ(un_gyr * dt));
                const V3 un_acc_1 = R * (acc - ba) - g;
                const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
                P = P + dt * V + (0.5 * dt * dt) * un_acc;
                V = V + dt * un_acc;
                stm(S_Rs(s, b, fc), R); st3(S_Ps(s, b, fc), P); st3(S_Vs(s, b, fc), V);
            }
        }
        __syncwarp();
        if (lane == 0) { st3(dv + DV_ACC0, acc); st3(dv + DV_GYR0, gyr); }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) addfeat_kernel(BeState s, const int *__restrict__ counts, const int *__restrict__ ids,
                                                      const double *__restrict__ xyz, const double *__restrict__ headers) {
    VIO_POISON(2u);
    __shared__ int sh_new[VIO_MAXP];
    __shared__ int sh_scan[33];
    __shared__ double sh_d[32];
    __shared__ int sh_found, sh_nnew;
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int fc = iv[IV_FRAME_COUNT], nf = iv[IV_NFEAT];
    const int n = min(counts[b], s.MAXCNT);
    const size_t fo = (size_t)b * s.FCAP;
    if (tid == 0) { sh_found = 0; sh_nnew = 0; }
    __syncthreads();
    const int *mid = ids + (size_t)b * s.MAXCNT;
    const double *mx = xyz + (size_t)b * s.MAXCNT * 3;
    for (int i = tid; i < n; i += 256) {
        // find_if over the list in list order (feature_manager.cpp:113-116).  The table is in INSERTION order like the reference's
        // std::list, which is not sorted by id once removeFailures() has erased a landmark the front end keeps tracking (its id comes
        // back and is appended behind larger ids), so the lookup is a scan, not a bisection.  All lanes read the same address.
        const int id = mid[i];
        int slot = -1;
#pragma unroll 8
        for (int m = 0; m < nf; m++)
            if (slot < 0 && s.f_id[fo + m] == id) slot = m;
        sh_new[i] = slot < 0;
        if (slot >= 0) {
            const int k = s.f_nobs[fo + slot];
            if (k < s.NF) {
                double *o = S_obs(s, b, slot) + 2 * k;
                const double z = mx[3 * i + 2];
                o[0] = mx[3 * i] / z; o[1] = mx[3 * i + 1] / z;
                s.f_nobs[fo + slot] = k + 1;
            }
            atomicAdd(&sh_found, 1);
        } else atomicAdd(&sh_nnew, 1);
    }
    __syncthreads();
    if (nf + sh_nnew > s.FCAP) { if (tid == 0) iv[IV_ERR] = VIO_ERR_CAPACITY; }
    else
        for (int i = tid; i < n; i += 256)
            if (sh_new[i]) {
                const int id = mid[i];
                int rank = 0;
                for (int j = 0; j < n; j++) rank += sh_new[j] && (mid[j] < id);
                const int slot = nf + rank;
                s.f_id[fo + slot] = id; s.f_start[fo + slot] = fc; s.f_nobs[fo + slot] = 1; s.f_depth[fo + slot] = -1.0; s.f_flag[fo + slot] = 0;
                double *o = S_obs(s, b, slot);
                const double z = mx[3 * i + 2];
                o[0] = mx[3 * i] / z; o[1] = mx[3 * i + 1] / z;
            }
    __syncthreads();
    const int nf2 = (nf + sh_nnew > s.FCAP) ? nf : nf + sh_nnew;
    const int last_track = sh_found;
    // parallax over features seen in both fc-2 and fc-1 (compensatedParallax2 with COMPENSATE_ROTATION false)
    double psum = 0; int pnum = 0;
    if (fc >= 2 && last_track >= 20)
        for (int k = tid; k < nf2; k += 256) {
            const int st = s.f_start[fo + k], no = s.f_nobs[fo + k];
            if (st <= fc - 2 && st + no - 1 >= fc - 1) {
                const double *o = S_obs(s, b, k);
                const double *pi = o + 2 * (fc - 2 - st), *pj = o + 2 * (fc - 1 - st);
                const double du = pi[0] - pj[0], dvv = pi[1] - pj[1];
                psum += sqrt(du * du + dvv * dvv);
                pnum++;
            }
        }
    psum = block_sum_d(psum, sh_d);
    int tot;
    block_excl_scan(pnum, sh_scan, &tot);
    if (tid == 0) {
        int marg = 0;
        if (!(fc < 2 || last_track < 20) && tot > 0) marg = (psum / tot >= s.min_parallax) ? 0 : 1;
        iv[IV_MARG_FLAG] = marg;
        iv[IV_NFEAT] = nf2;
        iv[IV_LAST_TRACK] = last_track;
        s.Headers[(size_t)b * s.NF + fc] = headers[b];
        if (iv[IV_SOLVER_FLAG] == 0) {      // all_image_frame.insert(header -> tmp_pre_integration); tmp = new IntegrationBase{acc_0, gyr_0, 0, 0}  (VINS.cpp:392-398)
            const int k = iv[IV_AF_N];
            if (k < s.FA) {
                s.af_hdr[(size_t)b * s.FA + k] = headers[b];
                iv[IV_AF_N] = k + 1;
                if (k + 1 < s.FA) {
                    const double *dvv = S_dv(s, b);
                    double *i0 = s.af_imu0 + ((size_t)b * s.FA + k + 1) * 6;
                    for (int c = 0; c < 3; c++) { i0[c] = dvv[DV_ACC0 + c]; i0[3 + c] = dvv[DV_GYR0 + c]; }
                    s.af_cnt[(size_t)b * s.FA + k + 1] = 0;
                    st3(s.af_abg + ((size_t)b * s.FA + k + 1) * 3, v3(0, 0, 0));
                } else iv[IV_AF_N] = s.FA + 1;  // list full (no room for the open record): poisoned until clearState, initialisation attempts are refused
            }
        }
        int act;
        if (iv[IV_SOLVER_FLAG] == 0) {
            if (fc != s.W) act = ACT_ACCUMULATE;
            else if (last_track < 20) act = ACT_CLEAR;                   // VINS.cpp:401-405: too few tracked points to initialise -> clearState()
            else act = iv[IV_INIT_PENDING] ? ACT_INIT_SOLVE : ACT_SLIDE_ONLY;
        } else act = ACT_NL_SOLVE;
        iv[IV_ACTION] = act;
        iv[IV_N_LM] = 0; iv[IV_N_FAC] = 0; iv[IV_ITERS] = 0; iv[IV_CHOL_RETRY] = 0; iv[IV_MARG_SWEEPS] = 0; iv[IV_MARG_FAST] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// smallest right singular vector of A (rows x 4) by one-sided (Hestenes) Jacobi; returns V[:,argmin sigma]
__device__ inline void smallest_right_sv4(double *A, int rows, double out[4]) {
    double V[16];
    for (int i = 0; i < 16; i++) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double a = 0, bb = 0, c = 0;
                for (int r = 0; r < rows; r++) { const double x = A[4 * r + p], y = A[4 * r + q]; a += x * x; bb += y * y; c += x * y; }
                if (fabs(c) <= 1e-300 || fabs(c) <= 2.3e-16 * sqrt(a * bb)) continue;
                off = fmax(off, fabs(c) / sqrt(a * bb));
                const double zeta = (bb - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < rows; r++) { const double x = A[4 * r + p], y = A[4 * r + q]; A[4 * r + p] = cs * x - sn * y; A[4 * r + q] = sn * x + cs * y; }
                for (int r = 0; r < 4; r++) { const double x = V[4 * r + p], y = V[4 * r + q]; V[4 * r + p] = cs * x - sn * y; V[4 * r + q] = sn * x + cs * y; }
            }
        if (off < 1e-15) break;
    }
    int best = 0; double bn = 1e300;
    for (int c = 0; c < 4; c++) { double nn = 0; for (int r = 0; r < rows; r++) nn += A[4 * r + c] * A[4 * r + c]; if (nn < bn) { bn = nn; best = c; } }
    for (int r = 0; r < 4; r++) out[r] = V[4 * r + best];
}

// FeatureManager::triangulate (feature_manager.cpp:190-257) for one stream: every in-solve landmark without a depth; `nt` threads of the CTA
__device__ inline void triangulate_stream(const BeState &s, int b, V3 tic, int tid, int nt) {
    const size_t fo = (size_t)b * s.FCAP;
    const int nf = S_iv(s, b)[IV_NFEAT];
    const double *dv = S_dv(s, b);
    const M3 ric = ldm(dv + DV_RIC);
    double A[4 * 2 * (VIO_MAX_WIN + 1)];
    for (int k = tid; k < nf; k += nt) {
        const int st = s.f_start[fo + k], no = s.f_nobs[fo + k];
        if (!in_solve(s, no, st) || s.f_depth[fo + k] > 0) continue;
        const M3 Rs0 = ldm(S_Rs(s, b, st));
        const V3 t0 = ld3(S_Ps(s, b, st)) + Rs0 * tic;
        const M3 R0 = Rs0 * ric;
        const M3 R0T = tr(R0);
        const double *o = S_obs(s, b, k);
        for (int j = 0; j < no; j++) {
            const M3 Rsj = ldm(S_Rs(s, b, st + j));
            const V3 t1 = ld3(S_Ps(s, b, st + j)) + Rsj * tic;
            const M3 R1 = Rsj * ric;
            const V3 t = R0T * (t1 - t0);
            const M3 R = R0T * R1;
            const M3 RT = tr(R);
            const V3 mt = -(RT * t);
            double P[12];
            for (int r = 0; r < 3; r++) { P[4 * r] = RT.m[3 * r]; P[4 * r + 1] = RT.m[3 * r + 1]; P[4 * r + 2] = RT.m[3 * r + 2]; }
            P[3] = mt.x; P[7] = mt.y; P[11] = mt.z;
            V3 f = v3(o[2 * j], o[2 * j + 1], 1.0);
            const double fn = norm(f);
            f = f * (1.0 / fn);
            for (int c = 0; c < 4; c++) {
                A[4 * (2 * j) + c] = f.x * P[8 + c] - f.z * P[c];
                A[4 * (2 * j + 1) + c] = f.y * P[8 + c] - f.z * P[4 + c];
            }
        }
        double v[4];
        smallest_right_sv4(A, 2 * no, v);
        double dep = v[2] / v[3];
        if (dep < 0.1) dep = s.init_depth;
        s.f_depth[fo + k] = dep;
    }
}

__global__ void __launch_bounds__(128) triangulate_kernel(BeState s) {
    VIO_POISON(4u);
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    const size_t fo = (size_t)b * s.FCAP;
    const int nf = iv[IV_NFEAT];
    if (act == ACT_INIT_SOLVE && iv[IV_INIT_PENDING] != 3) {      // 3: init_align_kernel (visualInitialAlign) has already prepared the window
        // caller-supplied window replaces solveInitial()/visualInitialAlign() (VINS.cpp:833-1102, out of scope)
        const double *in = s.init_state + (size_t)b * (s.NF * 10 + 6);
        for (int i = tid; i < s.NF; i += 128) {
            st3(S_Ps(s, b, i), ld3(in + 10 * i));
            stm(S_Rs(s, b, i), q2R(qnormalized(ldq(in + 10 * i + 3))));
            st3(S_Vs(s, b, i), ld3(in + 10 * i + 7));
            st3(S_Bas(s, b, i), ld3(in + 10 * s.NF)); st3(S_Bgs(s, b, i), ld3(in + 10 * s.NF + 3));
        }
        for (int k = tid; k < nf; k += 128) s.f_depth[fo + k] = -1.0;      // clearDepth(-1), VINS.cpp:1047-1050
        __syncthreads();
    }
    triangulate_stream(s, b, ld3(S_dv(s, b) + DV_TIC), tid, 128);
}

// ---------------------------------------------------------------------------------------------------------
// old2new() + enumeration of landmarks / projection factors in f_manager order + per-interval sqrt_info
__global__ void __launch_bounds__(256) prepare_kernel(BeState s) {
    VIO_POISON(8u);
    __shared__ int sh_scan[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    const size_t fo = (size_t)b * s.FCAP;
    const int nf = iv[IV_NFEAT];
    double *par = s.par + (size_t)b * s.par_stride;
    for (int i = tid; i < s.NF; i += 256) {
        double *p = par + 16 * i;
        st3(p, ld3(S_Ps(s, b, i)));
        stq(p + 3, R2q(ldm(S_Rs(s, b, i))));
        st3(p + 7, ld3(S_Vs(s, b, i))); st3(p + 10, ld3(S_Bas(s, b, i))); st3(p + 13, ld3(S_Bgs(s, b, i)));
    }
    // Loop closure (VINS.cpp:571-600): the retrieved keyframe must still be in the window (header >= Headers[0]) and equal the header of a
    // frame i < WINDOW_SIZE; the loop pose ("12th pose", solve frame NF) starts from para_Pose[i].  Without a match the extra frame is inert.
    __shared__ int sh_loop_i;
    if (s.loop_on) {
        if (tid == 0) {
            int li = -1;
            const double h = s.loop_hdr[b];
            if (act == ACT_NL_SOLVE && s.loop_n[b] > 0 && h >= s.Headers[(size_t)b * s.NF])
                for (int i = 0; i < s.W; i++) if (s.Headers[(size_t)b * s.NF + i] == h) li = i;
            sh_loop_i = li;
            iv[IV_LOOP_FRAME] = li; iv[IV_LOOP_NFAC] = 0;
            double *p = par + 16 * s.NF;
            for (int k = 0; k < 16; k++) p[k] = 0.0;
            p[6] = 1.0;
            if (li >= 0) { st3(p, ld3(S_Ps(s, b, li))); stq(p + 3, R2q(ldm(S_Rs(s, b, li)))); }
        }
        __syncthreads();
    }
    const int loop_i = s.loop_on ? sh_loop_i : -1;
    // IMUFactor(pre_integrations[i+1]): sqrt_info = LLT(cov^-1).matrixL()^T.  Cached while the covariance is unchanged (only the
    // newest frames propagate between two solves); warps 0 and 1 factor the ones that need it.
    __shared__ double sh_sqi[2][3][225];
    if (tid < 64) {
        const int w = tid >> 5, lane = tid & 31;
        for (int i = w; i < s.W; i += 2) {
            double *pr = S_pre(s, b, i + 1);
            if (pr[PR_SQI_OK] != 0.0) continue;
            if (!imu_sqrt_info_warp(pr + PR_COV, pr + PR_SQI, sh_sqi[w][0], sh_sqi[w][1], sh_sqi[w][2], lane) && lane == 0) iv[IV_ERR] = VIO_ERR_STATE;
            if (lane == 0) pr[PR_SQI_OK] = 1.0;
        }
    }
    // landmarks (in list order) and their factors
    int lm_base = 0, fac_base = 0;
    for (int c0 = 0; c0 < nf; c0 += 256) {
        const int k = c0 + tid;
        int is = 0, nfac = 0;
        if (k < nf) { const int st = s.f_start[fo + k], no = s.f_nobs[fo + k]; is = in_solve(s, no, st); nfac = is ? no - 1 : 0; }
        int tl, tf;
        const int li = lm_base + block_excl_scan(is, sh_scan, &tl);
        __syncthreads();
        const int fi = fac_base + block_excl_scan(nfac, sh_scan, &tf);
        __syncthreads();
        if (is) {
            if (li < s.LCAP && fi + nfac <= s.PCAP) {
                s.lm_slot[(size_t)b * s.LCAP + li] = k;
                s.lm_fac0[(size_t)b * (s.LCAP + 1) + li] = fi;
                s.lm_anchor[(size_t)b * s.LCAP + li] = s.f_start[fo + k];
                par[16 * s.NFS + li] = 1.0 / s.f_depth[fo + k];
                const int st = s.f_start[fo + k];
                for (int j = 1; j <= nfac; j++) { s.fac_lm[(size_t)b * s.PCAP + fi + j - 1] = li; s.fac_j[(size_t)b * s.PCAP + fi + j - 1] = st + j; }
            } else iv[IV_ERR] = VIO_ERR_CAPACITY;
        }
        lm_base += tl; fac_base += tf;
    }
    const int nl = min(lm_base, s.LCAP), nfac_all = min(fac_base, s.PCAP);
    if (tid == 0) { iv[IV_N_LM] = nl; iv[IV_N_FAC] = nfac_all; s.lm_fac0[(size_t)b * (s.LCAP + 1) + nl] = nfac_all; }
    // loop-closure factors: the reference walks the landmarks in solve order and the retrieved ids with one forward cursor (both ascending
    // by id, VINS.cpp:603-631); a landmark observed in frame loop_i whose id is found gets ProjectionFactor(first observation, old measurement)
    // between its anchor pose and the loop pose.  Sequential by construction (one thread).
    int n_loop_fac = 0;
    if (s.loop_on) {
        int *lml = s.lm_loop + (size_t)b * s.LCAP;
        __syncthreads();
        for (int l = tid; l < nl; l += 256) lml[l] = -1;
        __syncthreads();
        if (tid == 0 && loop_i >= 0) {
            const int n = min(s.loop_n[b], s.MAXCNT);
            const int *lid = s.loop_ids + (size_t)b * s.MAXCNT;
            int ri = 0, cnt = 0;
            for (int l = 0; l < nl && ri < n; l++) {
                const int k = s.lm_slot[(size_t)b * s.LCAP + l];
                const int st = s.f_start[fo + k], no = s.f_nobs[fo + k], id = s.f_id[fo + k];
                if (st <= loop_i && st + no - loop_i - 1 >= 0) {
                    while (ri < n && lid[ri] < id) ri++;
                    if (ri < n && lid[ri] == id && nfac_all + cnt < s.PCAP) { lml[l] = ri; ri++; cnt++; }
                }
            }
            iv[IV_LOOP_NFAC] = cnt;
        }
        __syncthreads();
        n_loop_fac = iv[IV_LOOP_NFAC];
    }
    if (tid == 0) iv[IV_N_FAC_ALL] = nfac_all + n_loop_fac;
    // the same factors ordered by (anchor frame i, observing frame j): a landmark contributes at most one factor to a pair, so
    // "landmark order within the pair" is a deterministic order.  One thread per pair counts, then places.
    __shared__ int sh_cnt[(VIO_MAX_WIN + 2) * (VIO_MAX_WIN + 2) + 1];
    const int NF = s.NFS, nkey = NF * NF, LF = s.NF;                 // LF: index of the loop pose among the solve frames
    const int nfac_tot = nfac_all + n_loop_fac;
    const int *slot = s.lm_slot + (size_t)b * s.LCAP;
    const int *lml = s.lm_loop + (size_t)b * s.LCAP;
    __shared__ unsigned short sh_lm[2048][2];                        // (anchor frame, observations) per landmark
    const int nls = min(nl, 2048);
    __syncthreads();
    for (int l = tid; l < nls; l += 256) { const int k = slot[l]; sh_lm[l][0] = (unsigned short)s.f_start[fo + k]; sh_lm[l][1] = (unsigned short)s.f_nobs[fo + k]; }
    __syncthreads();
    auto lm_start = [&](int l) { return l < 2048 ? (int)sh_lm[l][0] : s.f_start[fo + slot[l]]; };
    auto lm_nobs = [&](int l) { return l < 2048 ? (int)sh_lm[l][1] : s.f_nobs[fo + slot[l]]; };
    auto has_fac = [&](int l, int i, int j) {
        if (lm_start(l) != i) return false;
        return (s.loop_on && j == LF) ? (n_loop_fac > 0 && lml[l] >= 0) : (j < i + lm_nobs(l));
    };
    for (int key = tid; key < nkey; key += 256) {
        const int i = key / NF, j = key - i * NF;
        int c = 0;
        if (j > i)
            for (int l = 0; l < nl; l++) c += has_fac(l, i, j);
        sh_cnt[key] = c;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int key = 0; key < nkey; key++) { const int c = sh_cnt[key]; sh_cnt[key] = acc; acc += c; }
        sh_cnt[nkey] = acc;
    }
    __syncthreads();
    int *po = s.pair_off + (size_t)b * ((size_t)s.NFS * s.NFS + 1);
    for (int key = tid; key <= nkey; key += 256) po[key] = min(sh_cnt[key], nfac_tot);
    int *fs = s.fac_sorted + (size_t)b * s.PCAP;
    double *fobs = s.fac_obs + (size_t)b * s.PCAP * 4;
    const double *lxy = s.loop_xy + (size_t)b * s.MAXCNT * 2;
    for (int key = tid; key < nkey; key += 256) {
        const int i = key / NF, j = key - i * NF;
        if (j <= i) continue;
        int pos = sh_cnt[key];
        for (int l = 0; l < nl; l++) {
            if (has_fac(l, i, j) && pos < s.PCAP) {
                const double *o = S_obs(s, b, slot[l]);
                fs[pos] = l | (i << 16) | (j << 24);
                fobs[4 * (size_t)pos] = o[0]; fobs[4 * (size_t)pos + 1] = o[1];
                if (s.loop_on && j == LF) { fobs[4 * (size_t)pos + 2] = lxy[2 * lml[l]]; fobs[4 * (size_t)pos + 3] = lxy[2 * lml[l] + 1]; }
                else { fobs[4 * (size_t)pos + 2] = o[2 * (j - i)]; fobs[4 * (size_t)pos + 3] = o[2 * (j - i) + 1]; }
                pos++;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// new2old(): unpack + yaw / P0 re-anchoring (VINS.cpp:131-171), setDepth() flags (feature_manager.cpp:331-349)
__global__ void __launch_bounds__(256) post_solve_kernel(BeState s) {
    VIO_POISON(16u);
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    const int act = iv[IV_ACTION];
    if (act != ACT_INIT_SOLVE && act != ACT_NL_SOLVE) return;
    double *dv = S_dv(s, b);
    const double *par = s.par + (size_t)b * s.par_stride;
    V3 oR0 = R2ypr(ldm(S_Rs(s, b, 0)));
    V3 oP0 = ld3(S_Ps(s, b, 0));
    if (iv[IV_FAILURE]) { oR0 = R2ypr(ldm(dv + DV_LAST_R_OLD)); oP0 = ld3(dv + DV_LAST_P_OLD); }
    const V3 oR00 = R2ypr(q2R(ldq(par + 3)));
    const M3 rot = ypr2R(v3(oR0.x - oR00.x, 0, 0));
    const V3 p0 = ld3(par);
    if (s.loop_on && tid == 0) {
        // loop-closure results (VINS.cpp:664-680 from the raw solved parameters, :174-195 after the yaw / P0 re-anchoring)
        double *lo = s.loop_out + (size_t)b * 20;
        const int li = iv[IV_LOOP_FRAME];
        lo[12] = 0.0;
        if (li >= 0 && iv[IV_LOOP_NFAC] > 0) {
            const double *pl = par + 16 * s.NF, *pi = par + 16 * li;
            const M3 Ri = q2R(qnormalized(ldq(pi + 3))), Rl = q2R(qnormalized(ldq(pl + 3)));
            const V3 Pi = ld3(pi), Pl = ld3(pl);
            st3(lo, tr(Rl) * (Pi - Pl));
            stq(lo + 3, R2q(tr(Rl) * Ri));
            double ry = R2ypr(Ri).x - R2ypr(Rl).x;                   // Utility::normalizeAngle (degrees, utility.hpp:171-179)
            ry = ry > 0 ? ry - 360.0 * floor((ry + 180.0) / 360.0) : ry + 360.0 * floor((-ry + 180.0) / 360.0);
            lo[7] = ry;
            const M3 Rl2 = rot * Rl;
            const V3 Pl2 = rot * (Pl - p0) + oP0;
            const double *old = s.loop_old + (size_t)b * 7;
            const double dyaw = R2ypr(q2R(ldq(old + 3))).x - R2ypr(Rl2).x;
            lo[8] = dyaw;
            st3(lo + 9, ld3(old) - ypr2R(v3(dyaw, 0, 0)) * Pl2);
            lo[12] = 1.0;
        }
    }
    __syncthreads();                                   // everyone has read Rs[0]/Ps[0] before they are overwritten
    for (int i = tid; i < s.NF; i += 256) {
        const double *p = par + 16 * i;
        const M3 R = rot * q2R(qnormalized(ldq(p + 3)));
        const V3 P = rot * (ld3(p) - p0) + oP0;
        const V3 V = rot * ld3(p + 7);
        stm(S_Rs(s, b, i), R); st3(S_Ps(s, b, i), P); st3(S_Vs(s, b, i), V);
        st3(S_Bas(s, b, i), ld3(p + 10)); st3(S_Bgs(s, b, i), ld3(p + 13));
        double *ps = s.post_solve + ((size_t)b * s.NF + i) * 16;
        st3(ps, P); stq(ps + 3, R2q(R)); st3(ps + 7, V); st3(ps + 10, ld3(p + 10)); st3(ps + 13, ld3(p + 13));
    }
    const int nl = iv[IV_N_LM];
    const size_t fo = (size_t)b * s.FCAP;
    for (int l = tid; l < nl; l += 256) {
        const int k = s.lm_slot[(size_t)b * s.LCAP + l];
        const double d = 1.0 / par[16 * s.NFS + l];
        s.f_depth[fo + k] = d;
        s.f_flag[fo + k] = d < 0 ? 2 : 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// order-preserving compaction of the feature table; keep[k] != 0 survives.  tmp = scratch (>= FCAP*(5+2NF) doubles)
__device__ inline void compact_features(const BeState &s, int b, int nf, const unsigned char *keep, double *tmp, int *sh_scan, int *sh_total) {
    const int tid = threadIdx.x;
    const size_t fo = (size_t)b * s.FCAP;
    const int rowd = 5 + 2 * s.NF;
    int base = 0;
    for (int c0 = 0; c0 < nf; c0 += blockDim.x) {
        const int k = c0 + tid;
        const int kp = (k < nf) ? (keep[k] != 0) : 0;
        int tot;
        const int d = base + block_excl_scan(kp, sh_scan, &tot);
        __syncthreads();
        if (kp) {
            double *r = tmp + (size_t)d * rowd;
            r[0] = s.f_id[fo + k]; r[1] = s.f_start[fo + k]; r[2] = s.f_nobs[fo + k]; r[3] = s.f_flag[fo + k]; r[4] = s.f_depth[fo + k];
            const double *o = S_obs(s, b, k);
            for (int j = 0; j < 2 * s.NF; j++) r[5 + j] = o[j];
        }
        base += tot;
    }
    __syncthreads();
    for (int k = tid; k < base; k += blockDim.x) {
        const double *r = tmp + (size_t)k * rowd;
        s.f_id[fo + k] = (int)r[0]; s.f_start[fo + k] = (int)r[1]; s.f_nobs[fo + k] = (int)r[2]; s.f_flag[fo + k] = (int)r[3]; s.f_depth[fo + k] = r[4];
        double *o = S_obs(s, b, k);
        for (int j = 0; j < 2 * s.NF; j++) o[j] = r[5 + j];
    }
    __syncthreads();
    if (tid == 0) *sh_total = base;
    __syncthreads();
}

__device__ inline void clear_state_cta(const BeState &s, int b) {      // VINS::clearState, VINS.cpp:35-80
    const int tid = threadIdx.x;
    int *iv = S_iv(s, b);
    for (int i = tid; i < s.NF; i += blockDim.x) {
        stm(S_Rs(s, b, i), eye3()); st3(S_Ps(s, b, i), v3(0, 0, 0)); st3(S_Vs(s, b, i), v3(0, 0, 0));
        st3(S_Bas(s, b, i), v3(0, 0, 0)); st3(S_Bgs(s, b, i), v3(0, 0, 0));
        S_pre(s, b, i)[PR_VALID] = 0.0;
        s.imu_cnt[(size_t)b * s.NF + i] = 0;
        double *o = s.state_out + ((size_t)b * s.NF + i) * 16;           // what vio_backend_get_state / copy_state read
        st3(o, v3(0, 0, 0)); o[3] = 0; o[4] = 0; o[5] = 0; o[6] = 1; st3(o + 7, v3(0, 0, 0)); st3(o + 10, v3(0, 0, 0)); st3(o + 13, v3(0, 0, 0));
        s.Headers[(size_t)b * s.NF + i] = 0.0;
    }
    if (tid == 0) {
        iv[IV_FRAME_COUNT] = 0; iv[IV_FIRST_IMU] = 0; iv[IV_SOLVER_FLAG] = 0; iv[IV_NFEAT] = 0; iv[IV_PRIOR_VALID] = 0; iv[IV_PRIOR_N] = 0;
        iv[IV_ALLKEY] = 1; iv[IV_ALIGN_OK] = -1;                         // all_image_frame.clear(), VINS.cpp:62-68
        iv[IV_AF_N] = 0; s.af_cnt[(size_t)b * s.FA] = 0;
        for (int c = 0; c < 6; c++) s.af_imu0[(size_t)b * s.FA * 6 + c] = 0.0;
        st3(s.af_abg + (size_t)b * s.FA * 3, v3(0, 0, 0));
    }
}

__global__ void __launch_bounds__(256) finish_kernel(BeState s) {
    VIO_POISON(32u);
    __shared__ int sh_scan[33];
    __shared__ int sh_total, sh_fail;
    __shared__ PreScratch scr;
    extern __shared__ unsigned char keep[];            // FCAP bytes
    const int b = blockIdx.x, tid = threadIdx.x;
    int *iv = S_iv(s, b);
    double *dv = S_dv(s, b);
    const int act = iv[IV_ACTION];
    const int W = s.W;
    const size_t fo = (size_t)b * s.FCAP;
    double *tmp = s.scratch + (size_t)b * s.scratch_stride;
    if (act == ACT_ACCUMULATE) { if (tid == 0) iv[IV_FRAME_COUNT] += 1; }
    else if (act == ACT_CLEAR) clear_state_cta(s, b);
    else if (act != ACT_NONE) {
        // initialisation check (VINS.cpp:415-425): a first solve that ends above cost 200 is discarded -- the prior it built is deleted,
        // solver_flag stays INITIAL and the window only slides (no removeFailures, last_R / last_P untouched, failure_occur untouched)
        const bool init_rejected = act == ACT_INIT_SOLVE && dv[DV_COST1] > 200.0;
        const bool solved = (act == ACT_INIT_SOLVE || act == ACT_NL_SOLVE) && !init_rejected;
        if (tid == 0) {
            sh_fail = 0;
            if (act == ACT_NL_SOLVE) {                       // failureDetection(), VINS.cpp:214-265
                bool f = iv[IV_LAST_TRACK] < 4;
                f |= norm(ld3(S_Bgs(s, b, W))) > 1;
                const V3 tp = ld3(S_Ps(s, b, W)), lp = ld3(dv + DV_LAST_P);
                f |= norm(tp - lp) > 1;
                f |= fabs(tp.z - lp.z) > 0.5;
                const M3 dR = tr(ldm(S_Rs(s, b, W))) * ldm(dv + DV_LAST_R);
                const double ang = acos(R2q(dR).w) * 2.0 / 3.14 * 180.0;
                f |= ang > 40;
                sh_fail = f ? 1 : 0;
            }
            if (solved) iv[IV_FAILURE] = 0;
        }
        __syncthreads();
        if (sh_fail) {
            if (tid == 0) iv[IV_FAILURE] = 1;
            __syncthreads();
            clear_state_cta(s, b);
        } else {
            if (act == ACT_INIT_SOLVE && tid == 0) {
                if (init_rejected) { iv[IV_PRIOR_VALID] = 0; iv[IV_PRIOR_N] = 0; }
                else iv[IV_SOLVER_FLAG] = 1;
            }
            __syncthreads();
            const int marg = iv[IV_MARG_FLAG];
            const int nonlinear = iv[IV_SOLVER_FLAG] == 1;
            int nf = iv[IV_NFEAT];
            if (marg == 0) {                                 // ---- MARGIN_OLD: slideWindow() + slideWindowOld()
                const M3 backR0 = ldm(S_Rs(s, b, 0));
                const V3 backP0 = ld3(S_Ps(s, b, 0));
                __syncthreads();
                if (tid == 0) {
                    for (int i = 0; i < W; i++) {            // swaps = rotate-left of Rs, Ps, Vs, Headers (Bas/Bgs NOT shifted: quirk Q9)
                        const M3 R = ldm(S_Rs(s, b, i)); stm(S_Rs(s, b, i), ldm(S_Rs(s, b, i + 1))); stm(S_Rs(s, b, i + 1), R);
                        const V3 P = ld3(S_Ps(s, b, i)); st3(S_Ps(s, b, i), ld3(S_Ps(s, b, i + 1))); st3(S_Ps(s, b, i + 1), P);
                        const V3 V = ld3(S_Vs(s, b, i)); st3(S_Vs(s, b, i), ld3(S_Vs(s, b, i + 1))); st3(S_Vs(s, b, i + 1), V);
                        s.Headers[(size_t)b * s.NF + i] = s.Headers[(size_t)b * s.NF + i + 1];
                    }
                    s.Headers[(size_t)b * s.NF + W] = s.Headers[(size_t)b * s.NF + W - 1];
                    st3(S_Ps(s, b, W), ld3(S_Ps(s, b, W - 1))); st3(S_Vs(s, b, W), ld3(S_Vs(s, b, W - 1))); stm(S_Rs(s, b, W), ldm(S_Rs(s, b, W - 1)));
                    st3(S_Bas(s, b, W), ld3(S_Bas(s, b, W - 1))); st3(S_Bgs(s, b, W), ld3(S_Bgs(s, b, W - 1)));
                }
                // pre_integrations / buffers: shift left by one, slot W becomes a fresh IntegrationBase
                for (int i = 0; i < W; i++) {
                    double *d = S_pre(s, b, i); const double *sr = S_pre(s, b, i + 1);
                    for (int k = tid; k < PR_STRIDE; k += 256) d[k] = sr[k];
                    double *bd = S_imu(s, b, i); const double *bs = S_imu(s, b, i + 1);
                    const int c = s.imu_cnt[(size_t)b * s.NF + i + 1];
                    for (int k = tid; k < 7 * c; k += 256) bd[k] = bs[k];
                    __syncthreads();
                    if (tid == 0) s.imu_cnt[(size_t)b * s.NF + i] = c;
                    __syncthreads();
                }
                if (tid == 0) {
                    pre_init(S_pre(s, b, W), ld3(dv + DV_ACC0), ld3(dv + DV_GYR0), ld3(S_Bas(s, b, W)), ld3(S_Bgs(s, b, W)));
                    s.imu_cnt[(size_t)b * s.NF + W] = 0;
                    stm(dv + DV_BACK_R0, backR0); st3(dv + DV_BACK_P0, backP0);
                }
                __syncthreads();
                if (!nonlinear) {                            // all_image_frame.erase(begin, find(Headers[0])), VINS.cpp:1186-1193
                    const double t0 = s.Headers[(size_t)b * s.NF];
                    const int an = iv[IV_AF_N];
                    int m = 0;
                    while (m < an && s.af_hdr[(size_t)b * s.FA + m] != t0) m++;
                    if (m > 0 && m < an && an < s.FA) {
                        const int last = an;                 // the open record moves along
                        for (int r = m; r <= last; r++) {
                            const size_t src = (size_t)b * s.FA + r, dst = src - m;
                            const int c = s.af_cnt[src];
                            const double *bs = s.af_imu + src * s.MAXIMU * 7; double *bd = s.af_imu + dst * s.MAXIMU * 7;
                            for (int k = tid; k < 7 * c; k += 256) bd[k] = bs[k];
                            if (tid < 6) s.af_imu0[dst * 6 + tid] = s.af_imu0[src * 6 + tid];
                            if (tid < 3) s.af_abg[dst * 3 + tid] = s.af_abg[src * 3 + tid];
                            if (tid == 0) { s.af_cnt[dst] = c; if (r < an) s.af_hdr[dst] = s.af_hdr[src]; }
                            __syncthreads();
                        }
                        if (tid == 0) iv[IV_AF_N] = an - m;
                        __syncthreads();
                    }
                }
                // removeBackShiftDepth (NON_LINEAR) / removeBack (INITIAL)
                const M3 ric = ldm(dv + DV_RIC);
                const V3 tic = ld3(dv + DV_TIC);
                const M3 R0 = backR0 * ric, R1 = ldm(S_Rs(s, b, 0)) * ric;
                const V3 P0 = backP0 + backR0 * tic, P1 = ld3(S_Ps(s, b, 0)) + ldm(S_Rs(s, b, 0)) * tic;
                for (int k = tid; k < nf; k += 256) {
                    unsigned char kp = 1;
                    if (s.f_start[fo + k] != 0) s.f_start[fo + k] -= 1;
                    else {
                        double *o = S_obs(s, b, k);
                        const V3 uv = v3(o[0], o[1], 1.0);
                        const int no = s.f_nobs[fo + k] - 1;
                        for (int j = 0; j < 2 * no; j++) o[j] = o[j + 2];
                        s.f_nobs[fo + k] = no;
                        if (nonlinear) {
                            if (no < 2) kp = 0;
                            else {
                                const V3 w = R0 * (uv * s.f_depth[fo + k]) + P0;
                                const double dep = (tr(R1) * (w - P1)).z;
                                s.f_depth[fo + k] = dep > 0 ? dep : s.init_depth;
                            }
                        } else if (no == 0) kp = 0;
                    }
                    keep[k] = kp;
                }
                __syncthreads();
                compact_features(s, b, nf, keep, tmp, sh_scan, &sh_total);
                nf = sh_total;
            } else {                                         // ---- MARGIN_SECOND_NEW: slideWindow() + slideWindowNew()
                const int c = s.imu_cnt[(size_t)b * s.NF + W];
                if (!nonlinear && tid == 0) iv[IV_ALLKEY] = 0;  // the dropped frame stays in all_image_frame as a non-keyframe
                if (tid < 32) {                              // pre_integrations[W-1]->push_back(every buffered sample of frame W)
                    double *pr = S_pre(s, b, W - 1);
                    const double *buf = S_imu(s, b, W);
                    for (int i = 0; i < c; i++) {
                        pre_propagate_warp(pr, buf[7 * i], ld3(buf + 7 * i + 1), ld3(buf + 7 * i + 4), s.noise, scr);
                        if (tid == 0) {
                            int &cn = s.imu_cnt[(size_t)b * s.NF + W - 1];
                            if (cn < s.MAXIMU) { double *e = S_imu(s, b, W - 1) + 7 * cn; for (int k = 0; k < 7; k++) e[k] = buf[7 * i + k]; cn++; }
                            else iv[IV_ERR] = VIO_ERR_CAPACITY;
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    s.Headers[(size_t)b * s.NF + W - 1] = s.Headers[(size_t)b * s.NF + W];
                    st3(S_Ps(s, b, W - 1), ld3(S_Ps(s, b, W))); st3(S_Vs(s, b, W - 1), ld3(S_Vs(s, b, W))); stm(S_Rs(s, b, W - 1), ldm(S_Rs(s, b, W)));
                    st3(S_Bas(s, b, W - 1), ld3(S_Bas(s, b, W))); st3(S_Bgs(s, b, W - 1), ld3(S_Bgs(s, b, W)));
                    pre_init(S_pre(s, b, W), ld3(dv + DV_ACC0), ld3(dv + DV_GYR0), ld3(S_Bas(s, b, W)), ld3(S_Bgs(s, b, W)));
                    s.imu_cnt[(size_t)b * s.NF + W] = 0;
                }
                __syncthreads();
                const int fc = iv[IV_FRAME_COUNT];           // removeFront(frame_count)
                for (int k = tid; k < nf; k += 256) {
                    unsigned char kp = 1;
                    const int st = s.f_start[fo + k], no = s.f_nobs[fo + k];
                    if (st == fc) s.f_start[fo + k] = st - 1;
                    else if (st + no - 1 >= fc - 1) {
                        const int j = W - 1 - st;
                        double *o = S_obs(s, b, k);
                        for (int q = 2 * j; q < 2 * (no - 1); q++) o[q] = o[q + 2];
                        s.f_nobs[fo + k] = no - 1;
                        if (no - 1 == 0) kp = 0;
                    }
                    keep[k] = kp;
                }
                __syncthreads();
                compact_features(s, b, nf, keep, tmp, sh_scan, &sh_total);
                nf = sh_total;
            }
            if (solved) {                                    // removeFailures(): solve_flag == 2
                for (int k = tid; k < nf; k += 256) keep[k] = s.f_flag[fo + k] != 2;
                __syncthreads();
                compact_features(s, b, nf, keep, tmp, sh_scan, &sh_total);
                nf = sh_total;
                if (tid == 0) {
                    stm(dv + DV_LAST_R, ldm(S_Rs(s, b, W))); st3(dv + DV_LAST_P, ld3(S_Ps(s, b, W)));
                    stm(dv + DV_LAST_R_OLD, ldm(S_Rs(s, b, 0))); st3(dv + DV_LAST_P_OLD, ld3(S_Ps(s, b, 0)));
                }
            }
            if (tid == 0) iv[IV_NFEAT] = nf;
        }
    }
    __syncthreads();
    for (int i = tid; i < s.NF; i += 256) {                  // packed state for gathers / get_state
        double *o = s.state_out + ((size_t)b * s.NF + i) * 16;
        st3(o, ld3(S_Ps(s, b, i))); stq(o + 3, R2q(ldm(S_Rs(s, b, i)))); st3(o + 7, ld3(S_Vs(s, b, i)));
        st3(o + 10, ld3(S_Bas(s, b, i))); st3(o + 13, ld3(S_Bgs(s, b, i)));
    }
}

}  // namespace be
