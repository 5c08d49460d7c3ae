"""Front-end ORACLE -- test infrastructure only, never on the product path.

CPU restatement of the arithmetic behind `FeatureTracker::readImage`
(/root/reference/VINS_ios/feature_tracker.cpp:162-321).  The reference keeps
that arithmetic in OpenCV, which is NOT vendored under /root/reference
(VINS_ThirdPartyLib/opencv2.version:1 "A weird customized version based on
3.0.0", .gitignore:1).  The stand-in binary is Python `cv2` 4.13.0
(opencv-python-headless) which is present in this image on both the build box
and the GPU box.

Two layers live here:

* `cv2_*`  : thin calls into the real OpenCV binary with exactly the arguments
             the reference passes (feature_tracker.cpp:95,181,198,263).
* `r_*`    : numpy restatements of the same routines with an EXPLICIT operation
             order (named f32 / f64 steps, cv2's SIMD-lane summation order).  These are
             what the CUDA kernels are compared against bit-for-bit, and they are
             themselves pinned against `cv2_*` (tests/test_oracle_frontend.py and
             the committed fixtures in tests/golden/).

PARITY PINNING: the reference ships no golden vectors for this path (SURVEY.md
section 4); the pin is cv2 4.13.0 itself, run in the same process.

Image convention: img[y, x]; reference frame is ROW=640 rows x COL=480 cols
(feature_tracker.hpp:26-27).
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is only needed for the cv2_* layer and the cross-checks
    import cv2  # type: ignore
except Exception:  # pragma: no cover
    cv2 = None

f32 = np.float32
f64 = np.float64

LK_WIN = 21          # cv::Size(21,21)           feature_tracker.cpp:181
LK_LEVELS = 3        # maxLevel = 3               feature_tracker.cpp:181
LK_MAX_ITERS = 30    # default TermCriteria(COUNT+EPS, 30, 0.01)
LK_EPS = 0.01
LK_MIN_EIG = 1e-4    # default minEigThreshold
W_BITS = 14
FLT_SCALE = f32(1.0 / (1 << 20))
FLT_EPSILON = f32(1.1920929e-07)


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _fma32(x, y, z):
    """float32 fused multiply-add emulated through float64 (exact for f32 inputs)."""
    return (np.asarray(x, f64) * np.asarray(y, f64) + np.asarray(z, f64)).astype(f32)


def reflect101(i, n):
    """BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba), valid for |overshoot| < n."""
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * n - 2 - i, i)
    return i


# --------------------------------------------------------------------------
# K1: pyramid  (cv::pyrDown inside cv::calcOpticalFlowPyrLK / buildOpticalFlowPyramid)
# --------------------------------------------------------------------------
def r_pyr_down(img: np.ndarray) -> np.ndarray:
    """5-tap [1 4 6 4 1] separable, REFLECT_101, decimate by 2, single rounding (s+128)>>8.
    Output ((H+1)//2, (W+1)//2).  Exact vs cv2.pyrDown."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    a = img.astype(np.int32)
    ys = 2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :]
    xs = 2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :]
    ys = reflect101(ys, h)
    xs = reflect101(xs, w)
    k = np.array([1, 4, 6, 4, 1], np.int32)
    rows = (a[:, xs] * k[None, None, :]).sum(-1)          # (h, ow) horizontal taps at even centres
    out = (rows[ys, :] * k[None, :, None]).sum(1)          # (oh, ow)
    return ((out + 128) >> 8).astype(np.uint8)


def r_build_pyramid(img: np.ndarray, levels: int = LK_LEVELS):
    pyr = [np.ascontiguousarray(img)]
    for _ in range(levels):
        pyr.append(r_pyr_down(pyr[-1]))
    return pyr


def r_scharr(img: np.ndarray):
    """calcSharrDeriv: int16 Ix, Iy, REFLECT_101 on the level itself, no normalisation."""
    h, w = img.shape
    p = np.pad(img.astype(np.int32), 1, mode="reflect")
    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    ix = 3 * (s(-1, 1) - s(-1, -1)) + 10 * (s(0, 1) - s(0, -1)) + 3 * (s(1, 1) - s(1, -1))
    iy = 3 * (s(1, -1) - s(-1, -1)) + 10 * (s(1, 0) - s(-1, 0)) + 3 * (s(1, 1) - s(-1, 1))
    return ix.astype(np.int16), iy.astype(np.int16)


# --------------------------------------------------------------------------
# K4: pyramidal LK  (cv::calcOpticalFlowPyrLK, LKTrackerInvoker)
# --------------------------------------------------------------------------
_PAD = 24


def _rint_i32(x):
    return np.rint(x).astype(np.int32)


def _weights(frac_x, frac_y):
    a = frac_x.astype(f32)
    b = frac_y.astype(f32)
    one = f32(1.0)
    s = f32(1 << W_BITS)
    w00 = _rint_i32(((one - a) * (one - b)) * s)
    w01 = _rint_i32((a * (one - b)) * s)
    w10 = _rint_i32(((one - a) * b) * s)
    w11 = (1 << W_BITS) - w00 - w01 - w10
    return w00, w01, w10, w11


def _gather(img_pad, iy, ix, n):
    """(N, n, n) window whose top-left is (iy, ix) in UNPADDED coordinates."""
    yy = (iy[:, None] + np.arange(n)[None, :] + _PAD)[:, :, None]
    xx = (ix[:, None] + np.arange(n)[None, :] + _PAD)[:, None, :]
    return img_pad[yy, xx]


def _bilin(win, w00, w01, w10, w11, shift):
    """sum of 4 taps with integer weights, (x + 2^(shift-1)) >> shift."""
    n = win.shape[1] - 1
    v = (win[:, :n, :n] * w00[:, None, None] + win[:, :n, 1:] * w01[:, None, None]
         + win[:, 1:, :n] * w10[:, None, None] + win[:, 1:, 1:] * w11[:, None, None])
    return (v + (1 << (shift - 1))) >> shift


def _seq_sum_f32(vals):
    """Sequential f32 accumulation along the last axis, starting from +0 (one rounding per add)."""
    acc = np.zeros(vals.shape[:-1], f32)
    for k in range(vals.shape[-1]):
        acc = (acc + vals[..., k]).astype(f32)
    return acc


def _lk_sum_cov(p):
    """sum of the (N,21,21) exact integer products gx*gx / gx*gy / gy*gy in the order of OpenCV 4.13's LKTrackerInvoker
    (modules/video/src/lkpyramid.cpp, the CV_SIMD128 branch of the baseline SSE3 build -- the file is not dispatched):
    columns 0..15 of every row go through two 4-lane groups, lane k accumulating RN_f32(product) of columns k, k+4, k+8, k+12 row
    after row with separate mul/add (no FMA); columns 16..20 are added one by one, row-major, into the scalar accumulator;
    iA += v_reduce_sum(q) with the SSE reduction ((l0+l2)+(l1+l3)).  Pinned bit-exact against cv2 in tests/test_oracle_frontend.py."""
    pf = p.astype(f32)                                                   # RN of the exact product == f32 multiply of the int16s
    n = p.shape[0]
    lanes = _seq_sum_f32(np.stack([pf[:, :, [k, k + 4, k + 8, k + 12]].reshape(n, -1) for k in range(4)], 1))    # (N,4)
    tail = _seq_sum_f32(pf[:, :, 16:21].reshape(n, -1))
    red = ((lanes[:, 0] + lanes[:, 2]).astype(f32) + (lanes[:, 1] + lanes[:, 3]).astype(f32)).astype(f32)
    return (tail + red).astype(f32)


def _lk_sum_mismatch(pd):
    """sum of the (N,21,21) exact integer products diff*gx (or diff*gy), same source: v_dotprod adds the products of columns (k, k+4)
    and (k+8, k+12) exactly in int32, v_cvt_f32 rounds, and four f32 lanes k = 0..3 accumulate them row after row; the scalar tail
    takes columns 16..20 one by one; ib += ((l0+l2) + (l1+l3))."""
    n = pd.shape[0]
    ch = []
    for k in range(4):
        a = (pd[:, :, k] + pd[:, :, k + 4]).astype(f32)
        b = (pd[:, :, k + 8] + pd[:, :, k + 12]).astype(f32)
        ch.append(np.stack([a, b], 2).reshape(n, -1))
    lanes = _seq_sum_f32(np.stack(ch, 1))
    tail = _seq_sum_f32(pd[:, :, 16:21].astype(f32).reshape(n, -1))
    red = ((lanes[:, 0] + lanes[:, 2]).astype(f32) + (lanes[:, 1] + lanes[:, 3]).astype(f32)).astype(f32)
    return (tail + red).astype(f32)


def r_lk_track(prev_pyr, next_pyr, prev_pts: np.ndarray):
    """Restatement of calcOpticalFlowPyrLK(prev, next, pts, Size(21,21), 3) with default
    criteria/flags (feature_tracker.cpp:181).  Window products are exact integers; they are ACCUMULATED in f32 in the lane
    order of cv2's SSE code (_lk_sum_cov / _lk_sum_mismatch), which makes positions bit-identical to cv2.
    Returns (next_pts f32 (N,2), status u8 (N,))."""
    n = len(prev_pts)
    prev_pts = np.asarray(prev_pts, f32).reshape(n, 2)
    next_pts = np.zeros((n, 2), f32)
    status = np.ones(n, np.uint8)
    if n == 0:
        return next_pts, status
    half = f32((LK_WIN - 1) * 0.5)
    win = LK_WIN
    eps2 = f64(LK_EPS) * f64(LK_EPS)
    # buildOpticalFlowPyramid stops at the last level whose successor would not be larger than the window in both directions
    top = 0
    while top < min(LK_LEVELS, len(prev_pyr) - 1) and min(prev_pyr[top + 1].shape) > win:
        top += 1
    for level in range(top, -1, -1):
        I = prev_pyr[level]
        J = next_pyr[level]
        rows, cols = I.shape
        Ipad = np.pad(I.astype(np.int32), _PAD, mode="reflect")
        Jpad = np.pad(J.astype(np.int32), _PAD, mode="reflect")
        dx, dy = r_scharr(I)
        dxp = np.pad(dx.astype(np.int32), _PAD, mode="constant")
        dyp = np.pad(dy.astype(np.int32), _PAD, mode="constant")

        scale = f32(1.0 / (1 << level))
        prev = prev_pts * scale
        if level == top:
            nxt = prev.copy()
        else:
            nxt = next_pts * f32(2.0)
        next_pts = nxt.copy()

        prev = prev - half
        ipx = np.floor(prev[:, 0]).astype(np.int32)
        ipy = np.floor(prev[:, 1]).astype(np.int32)
        oob = (ipx < -win) | (ipx >= cols) | (ipy < -win) | (ipy >= rows)
        if level == 0:
            status[oob] = 0
        act = ~oob
        # clamp the indices of dead points so that gathers stay in range
        ipx_s = np.where(act, ipx, 0)
        ipy_s = np.where(act, ipy, 0)
        w00, w01, w10, w11 = _weights(prev[:, 0] - ipx.astype(f32), prev[:, 1] - ipy.astype(f32))
        Iw = _bilin(_gather(Ipad, ipy_s, ipx_s, win + 1), w00, w01, w10, w11, W_BITS - 5)
        gx = _bilin(_gather(dxp, ipy_s, ipx_s, win + 1), w00, w01, w10, w11, W_BITS)
        gy = _bilin(_gather(dyp, ipy_s, ipx_s, win + 1), w00, w01, w10, w11, W_BITS)
        gx = gx.astype(np.int64)
        gy = gy.astype(np.int64)
        A11 = _lk_sum_cov(gx * gx) * FLT_SCALE
        A12 = _lk_sum_cov(gx * gy) * FLT_SCALE
        A22 = _lk_sum_cov(gy * gy) * FLT_SCALE
        D = A11 * A22 - A12 * A12
        d12 = A11 - A22
        min_eig = ((A22 + A11) - np.sqrt(d12 * d12 + (f32(4.0) * A12) * A12)) / f32(2 * win * win)
        bad = (min_eig < f32(LK_MIN_EIG)) | (D < FLT_EPSILON)
        bad &= act
        if level == 0:
            status[bad] = 0
        act &= ~bad
        with np.errstate(divide="ignore"):
            Dinv = f32(1.0) / D

        nxt = nxt - half
        prev_delta = np.zeros((n, 2), f32)
        live = act.copy()
        for j in range(LK_MAX_ITERS):
            if not live.any():
                break
            inx = np.floor(nxt[:, 0]).astype(np.int32)
            iny = np.floor(nxt[:, 1]).astype(np.int32)
            oob = (inx < -win) | (inx >= cols) | (iny < -win) | (iny >= rows)
            oob &= live
            if level == 0:
                status[oob] = 0
            live &= ~oob
            if not live.any():
                break
            inx_s = np.where(live, inx, 0)
            iny_s = np.where(live, iny, 0)
            v00, v01, v10, v11 = _weights(nxt[:, 0] - inx.astype(f32), nxt[:, 1] - iny.astype(f32))
            Jw = _bilin(_gather(Jpad, iny_s, inx_s, win + 1), v00, v01, v10, v11, W_BITS - 5)
            diff = (Jw - Iw).astype(np.int64)
            b1 = _lk_sum_mismatch(diff * gx) * FLT_SCALE
            b2 = _lk_sum_mismatch(diff * gy) * FLT_SCALE
            dxv = ((A12 * b2 - A22 * b1) * Dinv).astype(f32)
            dyv = ((A12 * b1 - A11 * b2) * Dinv).astype(f32)
            delta = np.stack([dxv, dyv], 1)
            upd = live
            nxt = np.where(upd[:, None], nxt + delta, nxt)
            next_pts = np.where(upd[:, None], nxt + half, next_pts)
            dd = delta[:, 0].astype(f64) ** 2 + delta[:, 1].astype(f64) ** 2
            conv = upd & (dd <= eps2)
            live = live & ~conv
            if j > 0:
                osc = live & (np.abs(delta[:, 0] + prev_delta[:, 0]) < f32(0.01)) \
                           & (np.abs(delta[:, 1] + prev_delta[:, 1]) < f32(0.01))
                next_pts = np.where(osc[:, None], next_pts - delta * f32(0.5), next_pts)
                live = live & ~osc
            prev_delta = np.where(upd[:, None], delta, prev_delta)

        if level == 0:
            fin = next_pts - half
            fx = np.floor(fin[:, 0]).astype(np.int32)
            fy = np.floor(fin[:, 1]).astype(np.int32)
            oob = (fx < -win) | (fx >= cols) | (fy < -win) | (fy >= rows)
            status[(status == 1) & oob] = 0
    return next_pts.astype(f32), status


def cv2_lk_track(prev_img, next_img, prev_pts):
    p = np.asarray(prev_pts, f32).reshape(-1, 1, 2)
    nxt, st, _err = cv2.calcOpticalFlowPyrLK(prev_img, next_img, p, None, winSize=(21, 21), maxLevel=3)
    return nxt.reshape(-1, 2), st.reshape(-1).astype(np.uint8)


# --------------------------------------------------------------------------
# K2/K3: Shi-Tomasi (cv::goodFeaturesToTrack, useHarrisDetector=false)
# --------------------------------------------------------------------------
def r_min_eig_map(img: np.ndarray) -> np.ndarray:
    """cornerMinEigenVal(img, blockSize=3, ksize=3), BORDER_REFLECT_101.  Operation order chosen
    to be bit-identical with cv2 4.13.0 (AVX2 dispatch) -- verified in the oracle tests:
      scale a = float(1/(255*4*3));
      dx = fma(r[-1]+r[+1], a, r[0]*(2a)),   r[k] = p[y+k,x+1]-p[y+k,x-1]          (f32)
      dy = s[+1]-s[-1],  s[k] = fma(p[y+k,x+1], a, fma(p[y+k,x], 2a, a*p[y+k,x-1]))  (f32)
      cov = (dx*dx, dx*dy, dy*dy) in f32;  3x3 box sum accumulated in f64, rounded to f32
      eig = (A/2 + C/2) - sqrt((A/2-C/2)^2 + B*B)   (f32, no fma)."""
    h, w = img.shape
    p = np.pad(img.astype(np.int32), 1, mode="reflect")
    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    a = f32(1.0 / 3060.0)
    a2 = f32(2.0) * a
    r = {k: (s(k, 1) - s(k, -1)).astype(f32) for k in (-1, 0, 1)}
    dx = _fma32(r[-1] + r[1], a, r[0] * a2)
    def srow(k):
        return _fma32(s(k, 1).astype(f32), a, _fma32(s(k, 0).astype(f32), a2, a * s(k, -1).astype(f32)))
    dy = srow(1) - srow(-1)
    out = []
    for c in (dx * dx, dx * dy, dy * dy):
        pc = np.pad(c.astype(f64), 1, mode="reflect")
        rows = (pc[:, 0:w] + pc[:, 1:w + 1]) + pc[:, 2:w + 2]
        out.append(((rows[0:h] + rows[1:h + 1]) + rows[2:h + 2]).astype(f32))
    A, B, C = out
    ha = A * f32(0.5)
    hc = C * f32(0.5)
    t = ha - hc
    return (ha + hc) - np.sqrt(t * t + B * B)


def r_good_features(img, mask, max_corners, quality=0.01, min_dist=30.0, eig=None):
    """goodFeaturesToTrack(img, maxCorners, 0.01, 30, mask) with blockSize 3, min-eig detector
    (SURVEY Appendix A.3).  Ties in the value sort break towards the HIGHER address
    (later pixel first), as OpenCV's greaterThanPtr does.  Returns (K,2) f32 (x,y)."""
    if eig is None:
        eig = r_min_eig_map(img)
    h, w = eig.shape
    m = np.ones((h, w), bool) if mask is None else (mask != 0)
    if max_corners <= 0 or not m.any():
        return np.zeros((0, 2), f32)
    max_val = eig[m].max()
    thr = f32(f64(max_val) * f64(quality))            # threshold(eig, eig, maxVal*qualityLevel, 0, TOZERO)
    e = np.where(eig > thr, eig, f32(0))
    pe = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.full((h, w), -np.inf, f32)
    for dy in range(3):
        for dx in range(3):
            dil = np.maximum(dil, pe[dy:dy + h, dx:dx + w])
    cand = (e != 0) & (e == dil) & m
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = e[ys, xs]
    lin = ys.astype(np.int64) * w + xs
    order = np.lexsort((-lin, -vals.astype(f64)))       # value desc, then address desc
    ys, xs = ys[order], xs[order]
    cell = int(np.rint(min_dist))
    gw = (w + cell - 1) // cell
    gh = (h + cell - 1) // cell
    grid = [[] for _ in range(gw * gh)]
    md2 = min_dist * min_dist
    out = []
    for y, x in zip(ys.tolist(), xs.tolist()):
        if min_dist >= 1:
            xc, yc = x // cell, y // cell
            x1, y1 = max(0, xc - 1), max(0, yc - 1)
            x2, y2 = min(gw - 1, xc + 1), min(gh - 1, yc + 1)
            good = True
            for yy in range(y1, y2 + 1):
                for xx in range(x1, x2 + 1):
                    for (px, py) in grid[yy * gw + xx]:
                        ddx, ddy = x - px, y - py
                        if ddx * ddx + ddy * ddy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if not good:
                continue
            grid[yc * gw + xc].append((x, y))
        out.append((x, y))
        if len(out) >= max_corners:
            break
    return np.array(out, f32).reshape(-1, 2)


def cv2_good_features(img, mask, max_corners, quality=0.01, min_dist=30.0):
    if max_corners <= 0:
        return np.zeros((0, 2), f32)
    c = cv2.goodFeaturesToTrack(img, max_corners, quality, min_dist, mask=mask)
    return np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2).astype(f32)


# --------------------------------------------------------------------------
# K6: RANSAC fundamental matrix (cv::findFundamentalMat FM_RANSAC, 1.0, 0.99)
# --------------------------------------------------------------------------
class CvRNG:
    """cv::RNG multiply-with-carry; RANSACPointSetRegistrator seeds it with (uint64)-1."""
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _collinear_last(pts, count):
    """haveCollinearPoints(m, count): last point vs all pairs of earlier ones."""
    i = count - 1
    xi, yi = f64(pts[i, 0]), f64(pts[i, 1])
    for j in range(i):
        dx1 = f64(pts[j, 0]) - xi
        dy1 = f64(pts[j, 1]) - yi
        for k in range(j):
            dx2 = f64(pts[k, 0]) - xi
            dy2 = f64(pts[k, 1]) - yi
            if abs(dx2 * dy1 - dy2 * dx1) <= f64(FLT_EPSILON) * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _solve_cubic(c):
    """cv::solveCubic restated: coefficients c[0]x^3+c[1]x^2+c[2]x+c[3]; returns real roots list."""
    a0, a1, a2, a3 = (f64(v) for v in c)
    roots = []
    if a0 == 0:
        if a1 == 0:
            if a2 == 0:
                return [0.0] if a3 == 0 else []
            return [-a3 / a2]
        d = a2 * a2 - 4 * a1 * a3
        if d >= 0:
            d = np.sqrt(d)
            q1 = (-a2 + d) * 0.5
            q2 = (a2 + d) * -0.5
            if abs(q1) > abs(q2):
                roots = [q1 / a1, a3 / q1]
            else:
                roots = [q2 / a1, a3 / q2]
            if d == 0:
                roots = roots[:1]
        return roots
    a0 = 1.0 / a0
    a1 *= a0
    a2 *= a0
    a3 *= a0
    Q = (a1 * a1 - 3 * a2) * (1.0 / 9)
    R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54)
    Qcubed = Q * Q * Q
    d = Qcubed - R * R
    if d > 0:
        theta = np.arccos(R / np.sqrt(Qcubed))
        sqrtQ = np.sqrt(Q)
        t0 = -2 * sqrtQ
        t1 = theta * (1.0 / 3)
        t2 = a1 * (1.0 / 3)
        roots = [t0 * np.cos(t1) - t2,
                 t0 * np.cos(t1 + (2.0 * np.pi / 3)) - t2,
                 t0 * np.cos(t1 + (4.0 * np.pi / 3)) - t2]
    elif d == 0:
        if R >= 0:
            x0 = -2 * np.cbrt(R) - a1 / 3
            x1 = np.cbrt(R) - a1 / 3
        else:
            x0 = 2 * np.cbrt(-R) - a1 / 3
            x1 = -np.cbrt(-R) - a1 / 3
        roots = [x0, x1]
    else:
        d = np.sqrt(-d)
        e = np.cbrt(d + abs(R))
        if R > 0:
            e = -e
        roots = [(e + Q / e) - a1 * (1.0 / 3)]
    return [float(r) for r in roots]


# cv::SVD (JacobiSVDImpl_, modules/core/src/lapack.cpp) completes the two null-space rows of the FULL_UV decomposition of the 7x9
# system from pseudo-random vectors: entries +-1/9 with the sign of bit 8 of successive draws of cv::RNG(0x12345678), projected off
# the rows found so far and normalised.  The generator is re-seeded on every call, so the two vectors are constants:
_SVD_FILL_SIGNS = ((-1, -1, 1, -1, -1, -1, -1, 1, 1), (1, -1, 1, 1, 1, 1, 1, -1, 1))


def _null_space_7x9(A):
    """The basis (f1, f2) of the null space of the 7x9 epipolar system that run7Point() takes from cv::SVDecomp(FULL_UV): rows 7 and
    8 of Vt.  Both singular values are zero there, so OpenCV's Jacobi SVD GENERATES those rows (see _SVD_FILL_SIGNS): f1 is the
    normalised null-space component of the constant vector r1, f2 that of r2 made orthogonal to f1.  The basis matters: it fixes the
    ORDER in which solveCubic emits the up-to-three F candidates, and RANSAC keeps the first of equally good models.  The null space
    itself comes from an explicit Householder QR of A^T (9x7), columns 8 and 9 of Q -- the same algorithm, step for step, as
    run_7point() in csrc/fe_ransac.cuh; (f1, f2) agree with cv2.SVDecomp to round-off (checked in tests/test_oracle_frontend.py)."""
    M = A.T.astype(f64).copy()                        # 9 x 7
    beta = np.zeros(7)
    for k in range(7):
        nrm = np.sqrt(sum(M[r, k] * M[r, k] for r in range(k, 9)))
        if nrm == 0:
            continue
        alpha = -nrm if M[k, k] >= 0 else nrm
        M[k, k] -= alpha
        vv = sum(M[r, k] * M[r, k] for r in range(k, 9))
        beta[k] = 2.0 / vv if vv > 0 else 0.0
        for j in range(k + 1, 7):
            s = sum(M[r, k] * M[r, j] for r in range(k, 9)) * beta[k]
            for r in range(k, 9):
                M[r, j] -= s * M[r, k]
    out = []
    for c in range(2):
        y = np.zeros(9)
        y[7 + c] = 1.0
        for k in range(6, -1, -1):
            s = sum(M[r, k] * y[r] for r in range(k, 9)) * beta[k]
            for r in range(k, 9):
                y[r] -= s * M[r, k]
        out.append(y)
    n1, n2 = out
    # coordinates of r1, r2 in the orthonormal null-space basis (n1, n2); the common factor 1/9 drops out in the normalisation
    a1 = sum(_SVD_FILL_SIGNS[0][r] * n1[r] for r in range(9))
    a2 = sum(_SVD_FILL_SIGNS[0][r] * n2[r] for r in range(9))
    b1 = sum(_SVD_FILL_SIGNS[1][r] * n1[r] for r in range(9))
    b2 = sum(_SVD_FILL_SIGNS[1][r] * n2[r] for r in range(9))
    na = np.sqrt(a1 * a1 + a2 * a2)
    if na == 0:
        return n1, n2
    a1 /= na
    a2 /= na
    d = b1 * a1 + b2 * a2
    b1 -= d * a1
    b2 -= d * a2
    nb = np.sqrt(b1 * b1 + b2 * b2)
    if nb == 0:
        return n1, n2
    b1 /= nb
    b2 /= nb
    return a1 * n1 + a2 * n2, b1 * n1 + b2 * n2


def r_run_7point(m1, m2, null_space=_null_space_7x9):
    """run7Point of OpenCV 4.13 (modules/calib3d/src/fundam.cpp): the seven pairs are first normalised like the 8-point algorithm
    (centroid to the origin, mean distance sqrt 2), the null-space basis is rows 7, 8 of the FULL_UV SVD (_null_space_7x9), the
    cubic det(lambda f1' + f2) = 0 gives up to three F, each scaled to F[8] = 1, de-normalised (T2^T F T1) and scaled to F[8] = 1
    again.  The normalisation matters for parity: it changes the parametrisation of the pencil and with it the ORDER of the roots
    (checked against cv2.findFundamentalMat(FM_7POINT), which returns the candidates in that order).  Returns row-major 9-vectors."""
    m1 = np.asarray(m1, f32)
    m2 = np.asarray(m2, f32)
    c1x = c1y = c2x = c2y = f64(0)
    for i in range(7):
        c1x += f64(m1[i, 0]); c1y += f64(m1[i, 1]); c2x += f64(m2[i, 0]); c2y += f64(m2[i, 1])
    t = f64(1.0) / 7
    c1x *= t; c1y *= t; c2x *= t; c2y *= t
    sc1 = sc2 = f64(0)
    for i in range(7):
        dx, dy = f64(m1[i, 0]) - c1x, f64(m1[i, 1]) - c1y
        sc1 += np.sqrt(dx * dx + dy * dy)
        dx, dy = f64(m2[i, 0]) - c2x, f64(m2[i, 1]) - c2y
        sc2 += np.sqrt(dx * dx + dy * dy)
    sc1 *= t
    sc2 *= t
    if sc1 < f64(FLT_EPSILON) or sc2 < f64(FLT_EPSILON):
        return []
    sc1 = np.sqrt(f64(2.0)) / sc1
    sc2 = np.sqrt(f64(2.0)) / sc2
    A = np.zeros((7, 9), f64)
    for i in range(7):
        x0, y0 = (f64(m1[i, 0]) - c1x) * sc1, (f64(m1[i, 1]) - c1y) * sc1
        x1, y1 = (f64(m2[i, 0]) - c2x) * sc2, (f64(m2[i, 1]) - c2y) * sc2
        A[i] = [x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1.0]
    f1, f2 = null_space(A)
    # f1 := f1 - f2 so that F = lambda*f1' + f2  (det(lambda*f1 + (1-lambda)*f2) = 0)
    f1 = f1 - f2
    t0 = f2[4] * f2[8] - f2[5] * f2[7]
    t1 = f2[3] * f2[8] - f2[5] * f2[6]
    t2 = f2[3] * f2[7] - f2[4] * f2[6]
    c = [0.0] * 4
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2
    c[2] = (f1[0] * t0 - f1[1] * t1 + f1[2] * t2
            - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7])
            + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6])
            - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6])
            + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4])
            - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3])
            + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]))
    t0 = f1[4] * f1[8] - f1[5] * f1[7]
    t1 = f1[3] * f1[8] - f1[5] * f1[6]
    t2 = f1[3] * f1[7] - f1[4] * f1[6]
    c[1] = (f2[0] * t0 - f2[1] * t1 + f2[2] * t2
            - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7])
            + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6])
            - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6])
            + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4])
            - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3])
            + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]))
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2
    roots = _solve_cubic(c)
    out = []
    for lam in roots:
        mu = 1.0
        s = f1[8] * lam + f2[8]
        G = np.zeros(9, f64)
        if abs(s) > np.finfo(f64).eps:
            mu = 1.0 / s
            lam *= mu
            G[8] = 1.0
        for i in range(8):
            G[i] = f1[i] * lam + f2[i] * mu
        # de-normalise: F = T2^T G T1 with T = [s 0 -s cx; 0 s -s cy; 0 0 1]
        #   H = G T1 (columns), then F = T2^T H (rows)
        H = np.zeros(9, f64)
        for r in range(3):
            g0, g1, g2 = G[3 * r], G[3 * r + 1], G[3 * r + 2]
            H[3 * r] = g0 * sc1
            H[3 * r + 1] = g1 * sc1
            H[3 * r + 2] = g2 - (g0 * c1x + g1 * c1y) * sc1
        F = np.zeros(9, f64)
        for k in range(3):
            F[k] = H[k] * sc2
            F[3 + k] = H[3 + k] * sc2
            F[6 + k] = H[6 + k] - (H[k] * c2x + H[3 + k] * c2y) * sc2
        if abs(F[8]) > f64(FLT_EPSILON):
            F = F * (1.0 / F[8])
        out.append(F)
    return out


def r_fm_error(m1, m2, F):
    """FMEstimatorCallback::computeError: max of the two squared point-line distances, as f32."""
    x1 = m1[:, 0].astype(f64); y1 = m1[:, 1].astype(f64)
    x2 = m2[:, 0].astype(f64); y2 = m2[:, 1].astype(f64)
    a = F[0] * x1 + F[1] * y1 + F[2]
    b = F[3] * x1 + F[4] * y1 + F[5]
    c = F[6] * x1 + F[7] * y1 + F[8]
    s2 = 1.0 / (a * a + b * b)
    d2 = x2 * a + y2 * b + c
    a = F[0] * x2 + F[3] * y2 + F[6]
    b = F[1] * x2 + F[4] * y2 + F[7]
    c = F[2] * x2 + F[5] * y2 + F[8]
    s1 = 1.0 / (a * a + b * b)
    d1 = x1 * a + y1 * b + c
    return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(f32)


def _ransac_update_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(f64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(f64).tiny:
        return 0
    num = np.log(num)
    denom = np.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def r_find_fundamental_ransac(p1, p2, thresh=1.0, conf=0.99, max_iters=1000, null_space=_null_space_7x9,
                              trace=None):
    """findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask) for N >= 15 points
    (RANSACPointSetRegistrator::run with the 7-point kernel).  Returns mask u8 (N,) or None if no model.
    For 8 <= N < 15 OpenCV silently switches to LMedS: see r_find_fundamental."""
    p1 = np.asarray(p1, f32).reshape(-1, 2)
    p2 = np.asarray(p2, f32).reshape(-1, 2)
    count = len(p1)
    rng = CvRNG()
    niters = max(max_iters, 1)
    best_mask = None
    max_good = 0
    t2 = f64(thresh) * f64(thresh)
    it = 0
    while it < niters:
        # getSubset(m1, m2, ms1, ms2, rng, 10000)
        found = False
        idx = [0] * 7
        for _attempt in range(10000):
            i = 0
            ok = True
            while i < 7:
                idx_i = rng.uniform(0, count)
                while idx_i in idx[:i]:
                    idx_i = rng.uniform(0, count)
                idx[i] = idx_i
                i += 1
            ms1 = p1[idx]
            ms2 = p2[idx]
            if _collinear_last(ms1, 7) or _collinear_last(ms2, 7):
                ok = False
            if ok:
                found = True
                break
        if not found:
            if it == 0:
                return None
            break
        models = r_run_7point(ms1, ms2, null_space)
        if trace is not None:
            trace.append((list(idx), [m.copy() for m in models]))
        for F in models:
            err = r_fm_error(p1, p2, F)
            mask = err <= t2
            good = int(mask.sum())
            if good > max(max_good, 6):
                best_mask = mask
                max_good = good
                niters = _ransac_update_iters(conf, (count - good) / count, 7, niters)
        it += 1
    if best_mask is None:
        return None
    return best_mask.astype(np.uint8)


def cv2_find_fundamental(p1, p2, thresh=1.0, conf=0.99):
    F, mask = cv2.findFundamentalMat(np.asarray(p1, f32).reshape(-1, 1, 2), np.asarray(p2, f32).reshape(-1, 1, 2),
                                     cv2.FM_RANSAC, thresh, conf)
    if mask is None:
        return None
    return mask.reshape(-1).astype(np.uint8)


def r_find_fundamental_lmeds(p1, p2, conf=0.99, max_iters=1000, null_space=_null_space_7x9):
    """What findFundamentalMat(..., FM_RANSAC, ...) really runs for 8 <= N < 15 points:
    LMeDSPointSetRegistrator::run (outlierRatio 0.45 -> 300 iterations, getSubset maxAttempts 1000,
    median of f32 errors at index N/2, sigma = 2.5*1.4826*(1+5/(N-7))*sqrt(median), floor 0.001).
    Returns mask u8 (N,) or None when no model / fewer than 7 inliers."""
    p1 = np.asarray(p1, f32).reshape(-1, 2)
    p2 = np.asarray(p2, f32).reshape(-1, 2)
    count = len(p1)
    rng = CvRNG()
    niters = max(_ransac_update_iters(conf, 0.45, 7, max_iters), 3)
    min_median = np.inf
    best = None
    for it in range(niters):
        found = False
        idx = [0] * 7
        for _attempt in range(1000):
            for i in range(7):
                idx_i = rng.uniform(0, count)
                while idx_i in idx[:i]:
                    idx_i = rng.uniform(0, count)
                idx[i] = idx_i
            ms1 = p1[idx]
            ms2 = p2[idx]
            if not (_collinear_last(ms1, 7) or _collinear_last(ms2, 7)):
                found = True
                break
        if not found:
            if it == 0:
                return None
            break
        for F in r_run_7point(ms1, ms2, null_space):
            err = r_fm_error(p1, p2, F)
            med = f64(np.sort(err)[count // 2])
            if med < min_median:
                min_median = med
                best = F
    if best is None:
        return None
    sigma = 2.5 * 1.4826 * (1 + 5.0 / (count - 7)) * np.sqrt(min_median)
    sigma = max(sigma, 0.001)
    t = f32(sigma * sigma)
    mask = r_fm_error(p1, p2, best) <= t
    if int(mask.sum()) < 7:
        return None
    return mask.astype(np.uint8)


def r_find_fundamental(p1, p2, thresh=1.0, conf=0.99):
    """Dispatch exactly as cv::findFundamentalMat does for method FM_RANSAC (N >= 8 at the call sites,
    feature_tracker.cpp:91,194): RANSAC when N >= 15, LMedS otherwise."""
    n = len(p1)
    if n >= 15:
        return r_find_fundamental_ransac(p1, p2, thresh, conf)
    return r_find_fundamental_lmeds(p1, p2, conf)


# --------------------------------------------------------------------------
# a1: FeatureTracker::readImage restated (feature_tracker.cpp:162-321)
# --------------------------------------------------------------------------
def r_in_border(pts, rows, cols):
    """inBorder(): cvRound (round-half-even) then 1-px border test (feature_tracker.cpp:18-24)."""
    x = np.rint(pts[:, 0]).astype(np.int64)
    y = np.rint(pts[:, 1]).astype(np.int64)
    return (1 <= x) & (x < cols - 1) & (1 <= y) & (y < rows - 1)


def r_set_mask(pts, track_cnt, rows, cols, min_dist):
    """setMask() (feature_tracker.cpp:50-87).  std::sort there is unstable; canonicalised to
    (track_cnt desc, index asc) = stable sort.  Mask lookup and circle centre
    use cvRound of the float point; cv::circle(r, filled) == Euclidean disc dx^2+dy^2 <= r^2
    (SURVEY Appendix A.4).  An out-of-image rounded centre cannot occur (inBorder ran before).
    Returns (order-of-kept-indices, mask u8)."""
    n = len(pts)
    order = sorted(range(n), key=lambda i: (-int(track_cnt[i]), i))
    mask = np.full((rows, cols), 255, np.uint8)
    keep = []
    rx = np.rint(pts[:, 0]).astype(np.int64) if n else np.zeros(0, np.int64)
    ry = np.rint(pts[:, 1]).astype(np.int64) if n else np.zeros(0, np.int64)
    if cv2 is not None:                     # what the reference itself calls (feature_tracker.cpp:73,80)
        for i in order:
            cx, cy = int(rx[i]), int(ry[i])
            if mask[cy, cx] != 255:
                continue
            keep.append(i)
            cv2.circle(mask, (cx, cy), min_dist, 0, -1)
        return keep, mask
    yy, xx = np.mgrid[-min_dist:min_dist + 1, -min_dist:min_dist + 1]
    disc = (xx * xx + yy * yy) <= min_dist * min_dist
    for i in order:
        cx, cy = int(rx[i]), int(ry[i])
        if mask[cy, cx] != 255:
            continue
        keep.append(i)
        y0, y1 = max(0, cy - min_dist), min(rows, cy + min_dist + 1)
        x0, x1 = max(0, cx - min_dist), min(cols, cx + min_dist + 1)
        sub = disc[y0 - (cy - min_dist):y1 - (cy - min_dist), x0 - (cx - min_dist):x1 - (cx - min_dist)]
        mask[y0:y1, x0:x1][sub] = 0
    return keep, mask


class FeatureTrackerOracle:
    """Field-for-field restatement of FeatureTracker (feature_tracker.hpp:52-90).

    backend='cv2'      -> OpenCV binary does KLT / RANSAC-F / goodFeaturesToTrack (closest thing to
                          the reference that can run here);
    backend='restated' -> the r_* numpy restatements above (what the CUDA path is bit-compared to).
    `n_id` is per-instance (the reference's process-wide static, feature_tracker.cpp:11, made
    per-stream: SURVEY section 8(e)).  solveVinsPnP (feature_tracker.cpp:207) is out of scope
    (use_pnp defaults false)."""

    def __init__(self, rows=640, cols=480, max_cnt=70, min_dist=30, freq=3,
                 fx=526.600, fy=526.678, cx=243.481, cy=315.280, f_threshold=1.0, backend="cv2"):
        self.rows, self.cols = rows, cols
        self.max_cnt, self.min_dist, self.freq = max_cnt, min_dist, freq
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.f_threshold = f_threshold
        self.backend = backend
        self.cur_img = None
        self.forw_img = None
        self.cur_pyr = None
        self.e2 = lambda n: np.zeros((n, 2), f32)
        self.cur_pts = self.e2(0)
        self.pre_pts = self.e2(0)
        self.forw_pts = self.e2(0)
        self.ids = np.zeros(0, np.int32)
        self.track_cnt = np.zeros(0, np.int32)
        self.pmin = self.e2(0)      # parallax_cnt[i].min
        self.pmax = self.e2(0)      # parallax_cnt[i].max
        self.n_id = 0
        self.img_cnt = 0
        self.image_msg = {}
        self.mask = None
        self.stats = {}

    # -- primitives routed by backend -------------------------------------
    def _lk(self, cur_img, forw_img, cur_pyr, forw_pyr, pts):
        if self.backend == "cv2":
            return cv2_lk_track(cur_img, forw_img, pts)
        return r_lk_track(cur_pyr, forw_pyr, pts)

    def _fmat(self, p1, p2):
        if self.backend == "cv2":
            return cv2_find_fundamental(p1, p2, self.f_threshold, 0.99)
        return r_find_fundamental(p1, p2, self.f_threshold, 0.99)

    def _gftt(self, img, mask, n):
        if self.backend == "cv2":
            return cv2_good_features(img, mask, n, 0.01, float(self.min_dist))
        return r_good_features(img, mask, n, 0.01, float(self.min_dist))

    def _reduce(self, status, with_pre=True):
        s = status.astype(bool)
        if len(self.pre_pts) == len(s):     # reduceVector on an empty pre_pts is a no-op
            self.pre_pts = self.pre_pts[s]
        self.cur_pts = self.cur_pts[s]
        self.forw_pts = self.forw_pts[s]
        self.ids = self.ids[s]
        self.track_cnt = self.track_cnt[s]
        self.pmin = self.pmin[s]
        self.pmax = self.pmax[s]

    def _parallax_update(self, div):
        """feature_tracker.cpp:209-226 / :237-250 (UI-only outputs good_pts, track_len)."""
        p = self.forw_pts
        if len(p) == 0:
            return []
        lo = (p[:, 0] < self.pmin[:, 0]) | (p[:, 1] < self.pmin[:, 1])
        hi = ~lo & ((p[:, 0] > self.pmax[:, 0]) | (p[:, 1] > self.pmax[:, 1]))
        self.pmin[lo] = p[lo]
        self.pmax[hi] = p[hi]
        d = self.pmax.astype(f64) - self.pmin.astype(f64)
        nrm = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        parallax = np.where(nrm < 2.0, 0.0, nrm)
        track_len = np.minimum(1.0, 1.0 * parallax / div).tolist()
        return track_len

    def read_image(self, img):
        """readImage(); returns (good_pts, track_len).  Publishes image_msg on detect frames."""
        img = np.ascontiguousarray(img)
        forw_pyr = r_build_pyramid(img) if self.backend != "cv2" else None
        if self.forw_img is None:
            self.cur_img = img
            self.cur_pyr = forw_pyr
        self.forw_img = img
        self.forw_pts = self.e2(0)
        good_pts, track_len = [], []
        st = {}
        if len(self.cur_pts) > 0:
            nxt, status = self._lk(self.cur_img, self.forw_img, self.cur_pyr, forw_pyr, self.cur_pts)
            self.forw_pts = nxt
            st["lk_in"] = len(status)
            status = status & r_in_border(nxt, self.rows, self.cols).astype(np.uint8)
            st["lk_ok"] = int(status.sum())
            self._reduce(status)
            if len(self.forw_pts) >= 8:
                m = self._fmat(self.cur_pts, self.forw_pts)
                if m is not None:
                    self._reduce(m)
                st["f1_ok"] = len(self.forw_pts)
            if self.img_cnt != 0:
                track_len += self._parallax_update(30.0)
                good_pts += [p.copy() for p in self.forw_pts]
        if self.img_cnt == 0:
            if len(self.forw_pts) >= 8:                          # rejectWithF()
                m = self._fmat(self.pre_pts, self.forw_pts)
                if m is not None:
                    self._reduce(m)
                st["f2_ok"] = len(self.forw_pts)
            track_len += self._parallax_update(50.0)
            good_pts += [p.copy() for p in self.forw_pts]
            self.track_cnt = self.track_cnt + 1
            keep, mask = r_set_mask(self.forw_pts, self.track_cnt, self.rows, self.cols, self.min_dist)
            self.mask = mask
            self.forw_pts = self.forw_pts[keep]
            self.ids = self.ids[keep]
            self.track_cnt = self.track_cnt[keep]
            self.pmin = self.pmin[keep]
            self.pmax = self.pmax[keep]
            st["kept"] = len(keep)
            n_max = self.max_cnt - len(self.forw_pts)
            n_pts = self._gftt(self.forw_img, mask, n_max) if n_max > 0 else self.e2(0)
            st["new"] = len(n_pts)
            k = len(n_pts)                                       # addPoints()
            self.forw_pts = np.concatenate([self.forw_pts, n_pts.astype(f32)])
            self.ids = np.concatenate([self.ids, np.full(k, -1, np.int32)])
            self.track_cnt = np.concatenate([self.track_cnt, np.ones(k, np.int32)])
            self.pmin = np.concatenate([self.pmin, n_pts.astype(f32)])
            self.pmax = np.concatenate([self.pmax, n_pts.astype(f32)])
            self.pre_pts = self.forw_pts.copy()
            good_pts += [p.copy() for p in n_pts]
            track_len += [0.0] * k
        self.cur_img = self.forw_img
        self.cur_pyr = forw_pyr
        self.cur_pts = self.forw_pts.copy()
        if self.img_cnt == 0:
            for i in range(len(self.ids)):                       # updateID()
                if self.ids[i] == -1:
                    self.ids[i] = self.n_id
                    self.n_id += 1
            self.image_msg = {}
            for i in range(len(self.ids)):
                x = (f64(self.cur_pts[i, 0]) - self.cx) / self.fx
                y = (f64(self.cur_pts[i, 1]) - self.cy) / self.fy
                self.image_msg[int(self.ids[i])] = (float(x), float(y), 1.0)
        self.stats = st
        published = self.img_cnt == 0
        self.img_cnt = (self.img_cnt + 1) % self.freq          # ViewController.mm:494
        return good_pts, track_len, published


# ---------------------------------------------------------------------------------------------------------------------------------
# CLAHE -- the pre-processing the reference applies to the camera frame right before readImage (ViewController.mm:438-441:
# cv::createCLAHE(); setClipLimit(3); apply()).  Restatement of OpenCV's clahe.cpp (CLAHE_CalcLut_Body / CLAHE_Interpolation_Body) for
# 8-bit images whose size is divisible by the tile grid; pinned bit-exact against the cv2 binary in tests/test_oracle_frontend.py.
# ---------------------------------------------------------------------------------------------------------------------------------
def cv2_clahe(img, clip_limit=3.0, tiles=(8, 8)):
    import cv2
    c = cv2.createCLAHE(clipLimit=float(clip_limit), tileGridSize=(int(tiles[0]), int(tiles[1])))
    return c.apply(np.ascontiguousarray(img, np.uint8))


def r_clahe(img, clip_limit=3.0, tiles=(8, 8)):
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape
    tx, ty = int(tiles[0]), int(tiles[1])
    assert cols % tx == 0 and rows % ty == 0, "OpenCV pads with BORDER_REFLECT_101 otherwise (not restated)"
    tw, th = cols // tx, rows // ty
    area = tw * th
    lut_scale = np.float32(np.float32(255) / np.float32(area))
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    lut = np.zeros((ty * tx, 256), np.uint8)
    for k in range(tx * ty):
        y0, x0 = (k // tx) * th, (k % tx) * tw
        h = np.bincount(img[y0:y0 + th, x0:x0 + tw].ravel(), minlength=256).astype(np.int64)
        if clip > 0:
            clipped = int(np.maximum(h - clip, 0).sum())
            h = np.minimum(h, clip)
            batch = clipped // 256
            resid = clipped - batch * 256
            h += batch
            if resid:
                step = max(256 // resid, 1)
                i = 0
                while i < 256 and resid > 0:
                    h[i] += 1
                    i += step
                    resid -= 1
        v = np.cumsum(h).astype(np.float32) * lut_scale
        lut[k] = np.clip(np.rint(v), 0, 255).astype(np.uint8)                 # saturate_cast<uchar>(float) = cvRound
    one, half = np.float32(1.0), np.float32(0.5)
    inv_tw, inv_th = one / np.float32(tw), one / np.float32(th)
    txf = np.arange(cols, dtype=np.float32) * inv_tw - half
    tx1 = np.floor(txf).astype(np.int32)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = one - xa
    tx2 = np.minimum(tx1 + 1, tx - 1)
    tx1 = np.maximum(tx1, 0)
    tyf = np.arange(rows, dtype=np.float32) * inv_th - half
    ty1 = np.floor(tyf).astype(np.int32)
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = one - ya
    ty2 = np.minimum(ty1 + 1, ty - 1)
    ty1 = np.maximum(ty1, 0)
    L = lut.astype(np.float32)
    out = np.zeros_like(img)
    for y in range(rows):
        v = img[y].astype(np.int64)
        l11, l12 = L[ty1[y] * tx + tx1, v], L[ty1[y] * tx + tx2, v]
        l21, l22 = L[ty2[y] * tx + tx1, v], L[ty2[y] * tx + tx2, v]
        res = (l11 * xa1 + l12 * xa) * ya1[y] + (l21 * xa1 + l22 * xa) * ya[y]     # every f32 operation rounded separately
        out[y] = np.clip(np.rint(res), 0, 255).astype(np.uint8)
    return out
