// fe_track.cuh -- per-stream bookkeeping kernels of FeatureTracker::readImage (feature_tracker.cpp:162-321):
//   post_track_kernel : inBorder + reduceVector (K5), RANSAC-F x2 (K6), parallax bookkeeping, track_cnt++, setMask (K3a)
//   select_kernel     : goodFeaturesToTrack tail (threshold, value sort, 30-px greedy; K3b), addPoints, updateID, image_msg
// One CTA (256 threads) per stream; every array of the stream lives in shared memory for the duration.
#pragma once
#include "fe_kernels.cuh"
#include "fe_ransac.cuh"

namespace fe {

struct TrackArrays {
    float2 *cur_pts, *pre_pts, *forw_pts, *pmin, *pmax;
    int *ids, *track_cnt, *n;
    uint8_t *status;
    int2 *kept; int *n_kept;
    float2 *good_pts; double *track_len; int *n_good;      // UI outputs, stride 2 * max_cnt per stream (tracked before setMask + new corners)
    int *stats;
    int *n_id;
    int *msg_cnt, *msg_ids; double *msg_xyz;
    unsigned *max_bits; unsigned long long *cand; int *cand_cnt;
    int *err_flag;
};

struct TrackSmem {
    float2 cur[VIO_MAXP], pre[VIO_MAXP], forw[VIO_MAXP], pmin[VIO_MAXP], pmax[VIO_MAXP];
    int ids[VIO_MAXP], cnt[VIO_MAXP];
    uint8_t flag[VIO_MAXP];
    short pos[VIO_MAXP];
    RansacScratch rs;
    int n;
};

// order-preserving compaction of all seven parallel arrays by flag[] (reduceVector x7, feature_tracker.cpp:26-34,186-191)
__device__ inline void compact(TrackSmem &s) {
    const int tid = threadIdx.x, n = s.n;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        int p = 0;
        for (int j = 0; j < i; j++) p += s.flag[j];
        s.pos[i] = (short)p;
    }
    __syncthreads();
    float2 c[2], pr[2], fw[2], mn[2], mx[2]; int id[2], ct[2]; int dst[2];
    int k = 0;
    for (int i = tid; i < n; i += blockDim.x, k++) {
        dst[k] = s.flag[i] ? s.pos[i] : -1;
        c[k] = s.cur[i]; pr[k] = s.pre[i]; fw[k] = s.forw[i]; mn[k] = s.pmin[i]; mx[k] = s.pmax[i]; id[k] = s.ids[i]; ct[k] = s.cnt[i];
    }
    int total = 0;
    if (n > 0) total = s.pos[n - 1] + s.flag[n - 1];
    __syncthreads();
    k = 0;
    for (int i = tid; i < n; i += blockDim.x, k++) {
        if (dst[k] >= 0) {
            const int d = dst[k];
            s.cur[d] = c[k]; s.pre[d] = pr[k]; s.forw[d] = fw[k]; s.pmin[d] = mn[k]; s.pmax[d] = mx[k]; s.ids[d] = id[k]; s.cnt[d] = ct[k];
        }
    }
    if (tid == 0) s.n = total;
    __syncthreads();
}

// parallax_cnt min/max update + UI outputs (feature_tracker.cpp:209-226 / 237-250)
__device__ inline void parallax_update(TrackSmem &s, float2 *good, double *tl, double div) {
    for (int i = threadIdx.x; i < s.n; i += blockDim.x) {
        const float2 p = s.forw[i];
        if (p.x < s.pmin[i].x || p.y < s.pmin[i].y) s.pmin[i] = p;
        else if (p.x > s.pmax[i].x || p.y > s.pmax[i].y) s.pmax[i] = p;
        const double dx = (double)s.pmax[i].x - (double)s.pmin[i].x, dy = (double)s.pmax[i].y - (double)s.pmin[i].y;
        const double nrm = sqrt(dx * dx + dy * dy);
        const double par = nrm < 2.0 ? 0.0 : nrm;
        good[i] = p;
        tl[i] = fmin(1.0, 1.0 * par / div);
    }
}

__global__ void __launch_bounds__(256) post_track_kernel(TrackArrays A, int maxp, int rows, int cols, int min_dist, double f_thresh,
                                                         int detect, int have_tracks) {
    VIO_POISON(1024u);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrackSmem &s = *reinterpret_cast<TrackSmem *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t o = (size_t)b * maxp;
    int *st = A.stats + b * 8;
    if (tid == 0) { s.n = have_tracks ? A.n[b] : 0; for (int i = 0; i < 8; i++) st[i] = 0; }
    __syncthreads();
    const int n0 = s.n;
    for (int i = tid; i < n0; i += 256) {
        s.cur[i] = A.cur_pts[o + i]; s.pre[i] = A.pre_pts[o + i]; s.forw[i] = A.forw_pts[o + i];
        s.pmin[i] = A.pmin[o + i]; s.pmax[i] = A.pmax[o + i]; s.ids[i] = A.ids[o + i]; s.cnt[i] = A.track_cnt[o + i];
        // status && inBorder(forw): cvRound == rint (half to even), 1-px border (feature_tracker.cpp:18-24,183-185)
        const int x = __float2int_rn(s.forw[i].x), y = __float2int_rn(s.forw[i].y);
        s.flag[i] = A.status[o + i] && (1 <= x && x < cols - 1 && 1 <= y && y < rows - 1);
    }
    int n_good = 0;
    if (n0 > 0) {
        if (tid == 0) st[0] = n0;
        compact(s);
        if (tid == 0) st[1] = s.n;
        if (s.n >= 8) {                                       // feature_tracker.cpp:194-205
            int it = 0;
            const bool ok = find_fundamental_cta(s.cur, s.forw, s.n, f_thresh, 0.99, s.flag, s.rs, &it);
            if (tid == 0) st[7] = it;
            if (ok) compact(s);
        }
        if (tid == 0) st[2] = s.n;
        if (!detect) {
            parallax_update(s, A.good_pts + 2 * o, A.track_len + 2 * o, 30.0);
            n_good = s.n;
        }
    }
    if (detect) {
        if (s.n >= 8) {                                       // rejectWithF(), feature_tracker.cpp:89-103
            const bool ok = find_fundamental_cta(s.pre, s.forw, s.n, f_thresh, 0.99, s.flag, s.rs, nullptr);
            if (ok) compact(s);
        }
        if (tid == 0) st[3] = s.n;
        __syncthreads();
        parallax_update(s, A.good_pts + 2 * o, A.track_len + 2 * o, 50.0);
        n_good = s.n;
        for (int i = tid; i < s.n; i += 256) s.cnt[i] += 1;   // for (auto &n : track_cnt) n++
        __syncthreads();
        // setMask(): stable order by track_cnt desc, greedy keep when the rounded centre is outside every kept disc
        const int n = s.n;
        for (int i = tid; i < n; i += 256) {
            int r = 0;
            const int ci = s.cnt[i];
            for (int j = 0; j < n; j++) r += (s.cnt[j] > ci) || (s.cnt[j] == ci && j < i);
            s.pos[r] = (short)i;                             // pos[rank] = source index
        }
        __syncthreads();
        if (tid < 32) {                                       // one warp walks the sorted list
            int2 *kept = A.kept + o;
            int nk = 0;
            const int md2 = min_dist * min_dist;
            for (int r = 0; r < n; r++) {
                const int i = s.pos[r];
                const int cx = __float2int_rn(s.forw[i].x), cy = __float2int_rn(s.forw[i].y);
                bool hit = false;
                for (int k = tid; k < nk; k += 32) {
                    const int dx = cx - kept[k].x, dy = cy - kept[k].y;
                    hit |= (dx * dx + dy * dy <= md2);
                }
                hit = __any_sync(0xffffffffu, hit);
                if (tid == 0) s.flag[i] = hit ? 0 : 1;
                if (!hit) { if (tid == 0) kept[nk] = make_int2(cx, cy); nk++; }
                __syncwarp();
            }
            if (tid == 0) { A.n_kept[b] = nk; st[4] = nk; }
        }
        __syncthreads();
        // emit kept points in sorted order (forw_pts/ids/track_cnt/parallax_cnt are rebuilt in that order)
        {
            float2 fw[2], mn[2], mx[2], pr[2], cu[2]; int id[2], ct[2], dst[2];
            int k = 0;
            for (int r = tid; r < n; r += 256, k++) {
                const int i = s.pos[r];
                int d = -1;
                if (s.flag[i]) { d = 0; for (int q = 0; q < r; q++) d += s.flag[s.pos[q]]; }
                dst[k] = d; fw[k] = s.forw[i]; mn[k] = s.pmin[i]; mx[k] = s.pmax[i]; id[k] = s.ids[i]; ct[k] = s.cnt[i]; pr[k] = s.pre[i]; cu[k] = s.cur[i];
            }
            __syncthreads();
            k = 0;
            for (int r = tid; r < n; r += 256, k++)
                if (dst[k] >= 0) { const int d = dst[k]; s.forw[d] = fw[k]; s.pmin[d] = mn[k]; s.pmax[d] = mx[k]; s.ids[d] = id[k]; s.cnt[d] = ct[k]; s.pre[d] = pr[k]; s.cur[d] = cu[k]; }
            __syncthreads();
            if (tid == 0) s.n = A.n_kept[b];
            __syncthreads();
        }
        if (tid == 0) { A.max_bits[b] = 0u; A.cand_cnt[b] = 0; }
    }
    __syncthreads();
    // write back; cur_pts = forw_pts (feature_tracker.cpp:284-285).  On detect frames select_kernel appends the new corners.
    const int n1 = s.n;
    for (int i = tid; i < n1; i += 256) {
        A.cur_pts[o + i] = s.forw[i]; A.forw_pts[o + i] = s.forw[i]; A.pre_pts[o + i] = s.pre[i];
        A.pmin[o + i] = s.pmin[i]; A.pmax[o + i] = s.pmax[i]; A.ids[o + i] = s.ids[i]; A.track_cnt[o + i] = s.cnt[i];
    }
    if (tid == 0) { A.n[b] = n1; A.n_good[b] = n_good; }
}

// goodFeaturesToTrack tail + addPoints + updateID + image_msg (feature_tracker.cpp:257-307,311-321)
constexpr int SEL_BINS = 8192;          // histogram of (f32 bits >> 13) above the quality threshold: a x100 value range spans < 7 * 1024 bins
struct SelectSmem {
    unsigned long long key[SORT_CAP];
    int cell_cnt[2048];
    short cell_pt[2048][4][2];
    int hist[SEL_BINS];
    int csum[SEL_BINS / 32];               // sums of 32 consecutive histogram bins (band search skips whole chunks)
    int n_sorted, n_new, taken, band_lo, band_hi, band_err;
    short newxy[VIO_MAXP][2];
};

__global__ void __launch_bounds__(256) select_kernel(TrackArrays A, int maxp, int rows, int cols, int max_cnt, int min_dist,
                                                     double fx, double fy, double cx, double cy) {
    VIO_POISON(2048u);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SelectSmem &s = *reinterpret_cast<SelectSmem *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t o = (size_t)b * maxp;
    const int n_kept = A.n[b];
    const int n_max = max_cnt - n_kept;
    int *st = A.stats + b * 8;
    if (tid == 0) { s.n_sorted = 0; s.n_new = 0; s.taken = 0; s.band_err = 0; }
    const int gw = (cols + min_dist - 1) / min_dist, gh = (rows + min_dist - 1) / min_dist;
    for (int i = tid; i < gw * gh; i += 256) s.cell_cnt[i] = 0;
    __syncthreads();
    if (n_max > 0) {
        const int nc = min(A.cand_cnt[b], CAND_CAP);
        if (tid == 0) { st[6] = A.cand_cnt[b]; if (A.cand_cnt[b] > CAND_CAP) atomicExch(A.err_flag, VIO_ERR_CAPACITY); }
        // threshold(eig, maxVal*qualityLevel, TOZERO): thr is the f32 cast of the f64 product
        const float maxv = __uint_as_float(A.max_bits[b]);
        const float thr = (float)((double)maxv * 0.01);
        const unsigned thr_bin = __float_as_uint(thr) >> 13;
        const unsigned long long *cand = A.cand + (size_t)b * CAND_CAP;
        // Candidates are consumed in DESCENDING value order, in bands of at most SORT_CAP keys (one band in the common case; several when
        // a high-resolution frame yields tens of thousands of corners): histogram of the value bits -> band boundaries.
        for (int i = tid; i < SEL_BINS; i += 256) s.hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < nc; i += 256) {
            const unsigned vb = (unsigned)(cand[i] >> 32);
            if (__uint_as_float(vb) > thr) atomicAdd(&s.hist[min((vb >> 13) - thr_bin, (unsigned)SEL_BINS - 1)], 1);
        }
        __syncthreads();
        if (tid == 0) s.band_hi = SEL_BINS;                        // exclusive upper bin of the next band
        { int c = 0; for (int q = 0; q < 32; q++) c += s.hist[tid * 32 + ((q + tid) & 31)]; s.csum[tid] = c; }       // 256 threads x 32 bins = SEL_BINS
        __syncthreads();
        const int md2 = min_dist * min_dist;
        while (true) {
            if (tid == 0) {
                int lo = s.band_hi, cnt = 0;
                while (lo > 0) {
                    if ((lo & 31) == 0 && cnt + s.csum[(lo >> 5) - 1] <= SORT_CAP) { cnt += s.csum[(lo >> 5) - 1]; lo -= 32; continue; }
                    if (cnt + s.hist[lo - 1] <= SORT_CAP || lo == s.band_hi) { cnt += s.hist[lo - 1]; lo--; } else break;
                }
                if (cnt > SORT_CAP) { s.band_err = 1; cnt = 0; lo = 0; }      // > SORT_CAP keys inside one 2^-10-relative value bin
                s.band_lo = lo; s.n_sorted = 0;
            }
            __syncthreads();
            if (s.band_err) { if (tid == 0) atomicExch(A.err_flag, VIO_ERR_CAPACITY); break; }
            const int blo = s.band_lo, bhi = s.band_hi;
            for (int i = tid; i < nc; i += 256) {
                const unsigned long long k = cand[i];
                const unsigned vb = (unsigned)(k >> 32);
                if (__uint_as_float(vb) > thr) {
                    const int bin = (int)min((vb >> 13) - thr_bin, (unsigned)SEL_BINS - 1);
                    if (bin >= blo && bin < bhi) { const int d = atomicAdd(&s.n_sorted, 1); if (d < SORT_CAP) s.key[d] = k; }
                }
            }
            __syncthreads();
            const int ns = min(s.n_sorted, SORT_CAP);
            int np2 = 1;
            while (np2 < ns) np2 <<= 1;
            for (int i = ns + tid; i < np2; i += 256) s.key[i] = 0ull;
            __syncthreads();
            // bitonic sort, DESCENDING on (value bits, address): value desc then address desc == cv greaterThanPtr
            for (int k = 2; k <= np2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < np2; i += 256) {
                        const int l = i ^ j;
                        if (l > i) {
                            const unsigned long long a = s.key[i], c = s.key[l];
                            const bool desc = (i & k) == 0;
                            if (desc ? (a < c) : (a > c)) { s.key[i] = c; s.key[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            // greedy min-distance selection over the 3x3 cell neighbourhood (cell = rint(minDistance))
            if (tid == 0) {
                int taken = s.taken;
                for (int i = 0; i < ns && taken < n_max; i++) {
                    const unsigned lin = (unsigned)s.key[i];
                    const int y = lin / cols, x = lin - y * cols;
                    const int xc = x / min_dist, yc = y / min_dist;
                    const int x1 = max(0, xc - 1), y1 = max(0, yc - 1), x2 = min(gw - 1, xc + 1), y2 = min(gh - 1, yc + 1);
                    bool good = true;
                    for (int yy = y1; yy <= y2 && good; yy++)
                        for (int xx = x1; xx <= x2 && good; xx++) {
                            const int c = yy * gw + xx;
                            for (int q = 0; q < s.cell_cnt[c]; q++) {
                                const int dx = x - s.cell_pt[c][q][0], dy = y - s.cell_pt[c][q][1];
                                if (dx * dx + dy * dy < md2) { good = false; break; }
                            }
                        }
                    if (!good) continue;
                    const int c = yc * gw + xc;
                    if (s.cell_cnt[c] < 4) { s.cell_pt[c][s.cell_cnt[c]][0] = (short)x; s.cell_pt[c][s.cell_cnt[c]][1] = (short)y; s.cell_cnt[c]++; }
                    s.newxy[taken][0] = (short)x; s.newxy[taken][1] = (short)y;
                    taken++;
                }
                s.taken = taken;
                s.band_hi = s.band_lo;
            }
            __syncthreads();
            if (s.taken >= n_max || s.band_hi <= 0) break;
        }
        if (tid == 0) s.n_new = s.taken;
        __syncthreads();
    }
    const int n_new = s.n_new;
    if (tid == 0) { st[5] = n_new; }
    // addPoints(): id -1, track_cnt 1, parallax min=max=p ; then pre_pts = cur_pts = forw_pts
    for (int i = tid; i < n_new; i += 256) {
        const float2 p = make_float2((float)s.newxy[i][0], (float)s.newxy[i][1]);
        const size_t d = o + n_kept + i;
        A.forw_pts[d] = p; A.pmin[d] = p; A.pmax[d] = p; A.track_cnt[d] = 1;
        A.good_pts[2 * o + A.n_good[b] + i] = p; A.track_len[2 * o + A.n_good[b] + i] = 0.0;   // n_good <= max_cnt, n_new <= max_cnt
    }
    __syncthreads();
    const int n = n_kept + n_new;
    // updateID(): every -1 id (all new points; kept points already carry ids) gets n_id++ in order
    const int id0 = A.n_id[b];
    for (int i = tid; i < n; i += 256) {
        const size_t d = o + i;
        const float2 p = A.forw_pts[d];
        A.cur_pts[d] = p; A.pre_pts[d] = p;
        int id = i < n_kept ? A.ids[d] : id0 + (i - n_kept);
        A.ids[d] = id;
        A.msg_ids[d] = id;
        A.msg_xyz[3 * d + 0] = ((double)p.x - cx) / fx;
        A.msg_xyz[3 * d + 1] = ((double)p.y - cy) / fy;
        A.msg_xyz[3 * d + 2] = 1.0;
    }
    __syncthreads();
    if (tid == 0) { A.n[b] = n; A.msg_cnt[b] = n; A.n_id[b] = id0 + n_new; A.n_good[b] = A.n_good[b] + n_new; }
}

}  // namespace fe
