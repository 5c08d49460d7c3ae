// fe_ransac.cuh -- K6: cv::findFundamentalMat(p1, p2, FM_RANSAC, thresh, 0.99, mask) for one stream, executed by
// one CTA (reference call sites feature_tracker.cpp:95,198; oracle r_find_fundamental*).
//
// OpenCV's loop is sequential (each accepted model shrinks the iteration budget).  Here the RNG stream and the
// minimal subsets are generated sequentially by one thread exactly as getSubset() does, HB hypotheses are then
// solved (7-point: Householder null space + cubic) and scored in parallel, one warp each, and the accept/update
// rule is replayed in order -- the result is identical to the sequential loop.
#pragma once
#include "common.cuh"
#include <float.h>

namespace fe {

constexpr int RS_HB = 8;                       // hypotheses per round = warps per CTA
constexpr int RS_WORDS = VIO_MAXP / 32;

struct RansacScratch {
    int idx[RS_HB][7];
    int valid[RS_HB];
    double F[RS_HB][3][9];
    int nmodels[RS_HB];
    int good[RS_HB][3];
    float median[RS_HB][3];
    unsigned bits[RS_HB][3][RS_WORDS];
    unsigned best_bits[RS_WORDS];
    double bestF[9];
    unsigned long long rng;
    int niters, iter, max_good, done, have_best;
    double min_median;
};

__device__ __forceinline__ unsigned rng_next(unsigned long long &s) {
    s = (unsigned long long)(unsigned)s * 4164903690ull + (unsigned)(s >> 32);
    return (unsigned)s;
}
__device__ __forceinline__ int rng_uniform(unsigned long long &s, int a, int b) { return a == b ? a : (int)(rng_next(s) % (unsigned)(b - a)) + a; }

__device__ inline bool collinear_last(const float2 *p, const int *idx) {
    const double xi = p[idx[6]].x, yi = p[idx[6]].y;
    for (int j = 0; j < 6; j++) {
        const double dx1 = (double)p[idx[j]].x - xi, dy1 = (double)p[idx[j]].y - yi;
        for (int k = 0; k < j; k++) {
            const double dx2 = (double)p[idx[k]].x - xi, dy2 = (double)p[idx[k]].y - yi;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// getSubset(): sequential draws with duplicate rejection, then the collinearity check of the LAST point
__device__ inline bool get_subset(const float2 *p1, const float2 *p2, int count, unsigned long long &rng, int *idx, int max_attempts) {
    for (int it = 0; it < max_attempts; it++) {
        for (int i = 0; i < 7; i++) {
            int v;
            bool dup;
            do {
                v = rng_uniform(rng, 0, count);
                dup = false;
                for (int j = 0; j < i; j++) dup |= (idx[j] == v);
            } while (dup);
            idx[i] = v;
        }
        if (!collinear_last(p1, idx) && !collinear_last(p2, idx)) return true;
    }
    return false;
}

__device__ inline int solve_cubic(const double c[4], double r[3]) {          // cv::solveCubic
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) { if (a3 == 0) { r[0] = 0; return 1; } return 0; }
            r[0] = -a3 / a2; return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { r[0] = q1 / a1; r[1] = a3 / q1; } else { r[0] = q2 / a1; r[1] = a3 / q2; }
        return d == 0 ? 1 : 2;
    }
    a0 = 1.0 / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1.0 / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54);
    const double Qc = Q * Q * Q;
    double d = Qc - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qc)), sq = sqrt(Q), t0 = -2 * sq, t1 = theta * (1.0 / 3), t2 = a1 * (1.0 / 3);
        r[0] = t0 * cos(t1) - t2;
        r[1] = t0 * cos(t1 + (2.0 * 3.14159265358979323846 / 3)) - t2;
        r[2] = t0 * cos(t1 + (4.0 * 3.14159265358979323846 / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { r[0] = -2 * cbrt(R) - a1 / 3; r[1] = cbrt(R) - a1 / 3; }
        else { r[0] = 2 * cbrt(-R) - a1 / 3; r[1] = -cbrt(-R) - a1 / 3; }
        return 2;
    }
    d = sqrt(-d);
    double e = cbrt(d + fabs(R));
    if (R > 0) e = -e;
    r[0] = (e + Q / e) - a1 * (1.0 / 3);
    return 1;
}

// run7Point of OpenCV 4.13 (oracle r_run_7point): Hartley normalisation of the seven pairs, null space of the 7x9 epipolar system by
// Householder QR of A^T (columns 8, 9 of Q) rotated to the basis cv::SVDecomp(FULL_UV) generates for its two zero singular values
// (normalised null-space components of two constant +-1/9 vectors: oracle _SVD_FILL_SIGNS), cubic in lambda, de-normalisation.
// The basis and the normalisation fix the ORDER of the up-to-three candidates, and RANSAC keeps the first of equally good models.
__device__ inline int run_7point(const float2 *p1, const float2 *p2, const int *idx, double Fout[3][9]) {
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
    for (int i = 0; i < 7; i++) { c1x += (double)p1[idx[i]].x; c1y += (double)p1[idx[i]].y; c2x += (double)p2[idx[i]].x; c2y += (double)p2[idx[i]].y; }
    const double tt = 1.0 / 7;
    c1x *= tt; c1y *= tt; c2x *= tt; c2y *= tt;
    double sc1 = 0, sc2 = 0;
    for (int i = 0; i < 7; i++) {
        double dx = (double)p1[idx[i]].x - c1x, dy = (double)p1[idx[i]].y - c1y;
        sc1 += sqrt(dx * dx + dy * dy);
        dx = (double)p2[idx[i]].x - c2x; dy = (double)p2[idx[i]].y - c2y;
        sc2 += sqrt(dx * dx + dy * dy);
    }
    sc1 *= tt; sc2 *= tt;
    if (sc1 < (double)FLT_EPSILON || sc2 < (double)FLT_EPSILON) return 0;
    sc1 = sqrt(2.0) / sc1; sc2 = sqrt(2.0) / sc2;
    double M[9][7];                                    // A^T
    for (int i = 0; i < 7; i++) {
        const double x0 = ((double)p1[idx[i]].x - c1x) * sc1, y0 = ((double)p1[idx[i]].y - c1y) * sc1;
        const double x1 = ((double)p2[idx[i]].x - c2x) * sc2, y1 = ((double)p2[idx[i]].y - c2y) * sc2;
        M[0][i] = x1 * x0; M[1][i] = x1 * y0; M[2][i] = x1; M[3][i] = y1 * x0; M[4][i] = y1 * y0; M[5][i] = y1;
        M[6][i] = x0; M[7][i] = y0; M[8][i] = 1.0;
    }
    double beta[7];
    for (int k = 0; k < 7; k++) {                      // v_k overwrites M[k..8][k]
        double nrm = 0;
        for (int r = k; r < 9; r++) nrm += M[r][k] * M[r][k];
        nrm = sqrt(nrm);
        if (nrm == 0) { beta[k] = 0; continue; }
        const double alpha = M[k][k] >= 0 ? -nrm : nrm;
        M[k][k] -= alpha;
        double vv = 0;
        for (int r = k; r < 9; r++) vv += M[r][k] * M[r][k];
        beta[k] = vv > 0 ? 2.0 / vv : 0;
        for (int j = k + 1; j < 7; j++) {
            double s = 0;
            for (int r = k; r < 9; r++) s += M[r][k] * M[r][j];
            s *= beta[k];
            for (int r = k; r < 9; r++) M[r][j] -= s * M[r][k];
        }
    }
    double n1[9], n2[9];
    for (int c = 0; c < 2; c++) {
        double *y = c == 0 ? n1 : n2;
        for (int r = 0; r < 9; r++) y[r] = (r == 7 + c) ? 1.0 : 0.0;
        for (int k = 6; k >= 0; k--) {
            double s = 0;
            for (int r = k; r < 9; r++) s += M[r][k] * y[r];
            s *= beta[k];
            for (int r = k; r < 9; r++) y[r] -= s * M[r][k];
        }
    }
    // rotate (n1, n2) to the basis of cv::SVD: f1 ~ null-space part of r1, f2 ~ that of r2 made orthogonal to f1
    double f1[9], f2[9];
    {
        const double sg1[9] = {-1, -1, 1, -1, -1, -1, -1, 1, 1}, sg2[9] = {1, -1, 1, 1, 1, 1, 1, -1, 1};
        double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
        for (int r = 0; r < 9; r++) a1 += sg1[r] * n1[r];
        for (int r = 0; r < 9; r++) a2 += sg1[r] * n2[r];
        for (int r = 0; r < 9; r++) b1 += sg2[r] * n1[r];
        for (int r = 0; r < 9; r++) b2 += sg2[r] * n2[r];
        const double na = sqrt(a1 * a1 + a2 * a2);
        bool rot = na != 0;
        if (rot) {
            a1 /= na; a2 /= na;
            const double d = b1 * a1 + b2 * a2;
            b1 -= d * a1; b2 -= d * a2;
            const double nb = sqrt(b1 * b1 + b2 * b2);
            rot = nb != 0;
            if (rot) { b1 /= nb; b2 /= nb; }
        }
        for (int r = 0; r < 9; r++) {
            f1[r] = rot ? a1 * n1[r] + a2 * n2[r] : n1[r];
            f2[r] = rot ? b1 * n1[r] + b2 * n2[r] : n2[r];
        }
    }
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double c[4];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    double roots[3];
    const int n = solve_cubic(c, roots);
    for (int k = 0; k < n; k++) {
        double lambda = roots[k], mu = 1.0;
        const double s = f1[8] * lambda + f2[8];
        double G[9];
        G[8] = 0.0;
        if (fabs(s) > DBL_EPSILON) { mu = 1.0 / s; lambda *= mu; G[8] = 1.0; }
        for (int i = 0; i < 8; i++) G[i] = f1[i] * lambda + f2[i] * mu;
        // de-normalise F = T2^T G T1, T = [s 0 -s cx; 0 s -s cy; 0 0 1]
        double H[9];
        for (int r = 0; r < 3; r++) {
            const double g0 = G[3 * r], g1 = G[3 * r + 1], g2 = G[3 * r + 2];
            H[3 * r] = g0 * sc1; H[3 * r + 1] = g1 * sc1; H[3 * r + 2] = g2 - (g0 * c1x + g1 * c1y) * sc1;
        }
        double *F = Fout[k];
        for (int q = 0; q < 3; q++) {
            F[q] = H[q] * sc2; F[3 + q] = H[3 + q] * sc2; F[6 + q] = H[6 + q] - (H[q] * c2x + H[3 + q] * c2y) * sc2;
        }
        if (fabs(F[8]) > (double)FLT_EPSILON) { const double inv = 1.0 / F[8]; for (int i = 0; i < 9; i++) F[i] *= inv; }
    }
    return n;
}

__device__ __forceinline__ float fm_error(const double *F, float2 a, float2 b2) {
    const double x1 = a.x, y1 = a.y, x2 = b2.x, y2 = b2.y;
    double aa = F[0] * x1 + F[1] * y1 + F[2], bb = F[3] * x1 + F[4] * y1 + F[5], cc = F[6] * x1 + F[7] * y1 + F[8];
    const double s2 = 1.0 / (aa * aa + bb * bb), d2 = x2 * aa + y2 * bb + cc;
    aa = F[0] * x2 + F[3] * y2 + F[6]; bb = F[1] * x2 + F[4] * y2 + F[7]; cc = F[2] * x2 + F[5] * y2 + F[8];
    const double s1 = 1.0 / (aa * aa + bb * bb), d1 = x1 * aa + y1 * bb + cc;
    return (float)fmax(d1 * d1 * s1, d2 * d2 * s2);
}

__device__ inline int ransac_update_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.), 1.); ep = fmin(fmax(ep, 0.), 1.);
    double num = fmax(1. - p, DBL_MIN), denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// Whole-CTA routine (blockDim.x == RS_HB*32).  Writes mask[i] (0/1) for i < n and returns true when a model was found
// (cv::findFundamentalMat returns an empty mask otherwise and the reference then keeps everything, see
// feature_tracker.cpp:194-205 where reduceVector runs over an empty status: we keep all points).
// n >= 15: RANSAC;  8 <= n < 15: LMedS, as OpenCV silently does for FM_RANSAC.
__device__ inline bool find_fundamental_cta(const float2 *p1, const float2 *p2, int n, double thresh, double conf, uint8_t *mask,
                                            RansacScratch &S, int *iters_out) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool lmeds = n < 15;
    const float t = (float)(thresh * thresh);
    if (tid == 0) {
        S.rng = 0xFFFFFFFFFFFFFFFFull;
        S.niters = lmeds ? max(ransac_update_iters(conf, 0.45, 7, 1000), 3) : 1000;
        S.iter = 0; S.max_good = 0; S.done = 0; S.have_best = 0; S.min_median = DBL_MAX;
    }
    __syncthreads();
    while (true) {
        if (tid == 0) {
            for (int h = 0; h < RS_HB; h++) {
                S.valid[h] = 0;
                if (S.iter + h >= S.niters) break;
                if (!get_subset(p1, p2, n, S.rng, S.idx[h], lmeds ? 1000 : 10000)) { S.valid[h] = -1; break; }
                S.valid[h] = 1;
            }
        }
        __syncthreads();
        if (S.valid[warp] == 1) {
            if (lane == 0) S.nmodels[warp] = run_7point(p1, p2, S.idx[warp], S.F[warp]);
            __syncwarp();
            const int nm = S.nmodels[warp];
            for (int m = 0; m < nm; m++) {
                const double *F = S.F[warp][m];
                if (!lmeds) {
                    int cnt = 0;
                    for (int base = 0; base < n; base += 32) {
                        const int i = base + lane;
                        const bool in = i < n && fm_error(F, p1[i], p2[i]) <= t;
                        const unsigned bal = __ballot_sync(0xffffffffu, in);
                        if (lane == 0) S.bits[warp][m][base >> 5] = bal;
                        cnt += __popc(bal);
                    }
                    if (lane == 0) S.good[warp][m] = cnt;
                } else {                                   // n <= 14: one lane per point, median = element n/2 of the sorted errors
                    const float e = lane < n ? fm_error(F, p1[lane], p2[lane]) : FLT_MAX;
                    int rank = 0;
                    for (int j = 0; j < n; j++) {
                        const float ej = __shfl_sync(0xffffffffu, e, j);
                        rank += (ej < e) || (ej == e && j < lane);
                    }
                    if (lane < n && rank == n / 2) S.median[warp][m] = e;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            for (int h = 0; h < RS_HB && !S.done; h++) {
                if (S.iter >= S.niters) { S.done = 1; break; }
                if (S.valid[h] == -1) { S.done = 1; break; }          // getSubset failed: return false at iter 0, break otherwise
                if (S.valid[h] != 1) { S.done = 1; break; }
                for (int m = 0; m < S.nmodels[h]; m++) {
                    if (!lmeds) {
                        const int g = S.good[h][m];
                        if (g > max(S.max_good, 6)) {
                            for (int w = 0; w < RS_WORDS; w++) S.best_bits[w] = S.bits[h][m][w];
                            S.max_good = g; S.have_best = 1;
                            S.niters = ransac_update_iters(conf, (double)(n - g) / n, 7, S.niters);
                        }
                    } else {
                        const double med = S.median[h][m];
                        if (med < S.min_median) { S.min_median = med; S.have_best = 1; for (int i = 0; i < 9; i++) S.bestF[i] = S.F[h][m][i]; }
                    }
                }
                S.iter++;
            }
            if (S.iter >= S.niters) S.done = 1;
        }
        __syncthreads();
        if (S.done) break;
    }
    bool ok = S.have_best != 0;
    if (ok && lmeds) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(S.min_median);
        sigma = fmax(sigma, 0.001);
        const float ts = (float)(sigma * sigma);
        const bool in = tid < n && fm_error(S.bestF, p1[tid], p2[tid]) <= ts;
        const int cnt = __syncthreads_count(in);
        if (tid < n) mask[tid] = in;
        ok = cnt >= 7;
    } else if (ok) {
        for (int i = tid; i < n; i += blockDim.x) mask[i] = (S.best_bits[i >> 5] >> (i & 31)) & 1u;
    }
    if (tid == 0 && iters_out) *iters_out = S.iter;
    __syncthreads();
    return ok;
}

}  // namespace fe
