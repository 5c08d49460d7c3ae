"""ctypes loader for oracle/_ref/libvins_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own factor code + vendored Ceres 1.12.0 driven by the restated
estimator loop in oracle/backend_ref.cpp; it is built by oracle/Makefile from /root/reference and
travels to the GPU box as a prebuilt file (oracle/_ref/ is git-ignored, not gpurun-ignored)."""
import ctypes as C
import importlib
import os

import numpy as np

_abi = importlib.import_module("vins-mobile_b200.abi")
DP, IP = _abi.DP, _abi.IP
_HERE = os.path.dirname(os.path.abspath(__file__))
# VINS_REF_LIB selects another build of the same sources (bench.py times libvins_ref_o3.so, -O3 -march=x86-64-v3, next to the default -O2)
LIB_PATH = os.path.join(_HERE, "_ref", os.environ.get("VINS_REF_LIB", "libvins_ref.so"))
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.vref_create.restype = C.c_void_p
        L.vref_create.argtypes = [C.POINTER(_abi.VioConfig)]
        L.vref_destroy.argtypes = [C.c_void_p]
        L.vref_clear.argtypes = [C.c_void_p]
        L.vref_process_imu.argtypes = [C.c_void_p, C.c_double, DP, DP]
        L.vref_set_init_window.argtypes = [C.c_void_p, DP, DP, DP, DP, DP]
        L.vref_process_image.argtypes = [C.c_void_p, C.c_int, IP, DP, C.c_double]
        L.vref_get_state.argtypes = [C.c_void_p, DP, DP, DP, DP, DP, DP]
        L.vref_get_post_solve.argtypes = [C.c_void_p, DP]
        L.vref_get_info.argtypes = [C.c_void_p, IP, DP]
        L.vref_get_features.argtypes = [C.c_void_p, C.c_int, IP, IP, IP, IP, DP, IP]
        L.vref_get_prior.argtypes = [C.c_void_p, DP, DP, IP, DP]
        L.vref_stage_seconds.argtypes = [C.c_void_p, DP, C.c_int]
        L.vref_solve.argtypes = [C.c_void_p]
        L.vref_set_loop_match.argtypes = [C.c_void_p, C.c_int, C.c_double, IP, DP, DP]
        L.vref_get_loop_result.argtypes = [C.c_void_p, DP]
        L.vref_preintegrate.argtypes = [C.c_int, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP]
        L.vref_imu_factor.argtypes = [DP, DP, DP, C.c_double, DP, DP, DP, DP, DP, DP, DP, DP]
        if hasattr(L, "vref_imu_sqrt_info"):
            L.vref_imu_sqrt_info.argtypes = [DP, DP]
        L.vref_projection_factor.argtypes = [C.c_double, DP, DP, DP, DP, DP, DP, C.c_double, DP, DP]
        if hasattr(L, "vref_set_init_sfm"):
            L.vref_set_init_sfm.argtypes = [C.c_void_p, DP, DP]
            L.vref_set_init_sfm_frames.argtypes = [C.c_void_p, C.c_int, DP, DP]
            L.vref_get_init_frames.argtypes = [C.c_void_p, C.c_int, DP]
            L.vref_get_init_result.argtypes = [C.c_void_p, IP, DP, DP]
        if hasattr(L, "vref_visual_imu_align"):
            L.vref_visual_imu_align.argtypes = [C.c_int, DP, DP, IP, DP, C.c_int, DP, DP, DP, DP, DP, DP]
        if hasattr(L, "vpnp_create"):
            L.vpnp_create.restype = C.c_void_p
            L.vpnp_create.argtypes = [C.POINTER(_abi.VioConfig)]
            L.vpnp_destroy.argtypes = [C.c_void_p]
            L.vpnp_set_init.argtypes = [C.c_void_p, C.c_double, DP, DP, DP, DP, DP]
            L.vpnp_process_imu.argtypes = [C.c_void_p, C.c_double, DP, DP]
            L.vpnp_process_image.argtypes = [C.c_void_p, C.c_int, IP, DP, DP, IP, C.c_double, C.c_int]
            L.vpnp_get_state.argtypes = [C.c_void_p, DP, DP, DP, DP, DP, DP, IP, IP]
            L.vpnp_perspective_factor.argtypes = [DP, DP, C.c_int, C.c_double, DP, DP, DP, DP, DP]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, np.float64)


class RefEstimator:
    """One VINS object (single stream) of the reference back end."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.W = cfg.window_size
        self.h = lib().vref_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("vref_create failed")

    def close(self):
        if self.h:
            lib().vref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self):
        return lib().vref_solve(self.h)

    def set_loop_match(self, header, ids, xy, pose_old):
        ids = np.ascontiguousarray(ids, np.int32); xy = _d(xy); pose_old = _d(pose_old)
        lib().vref_set_loop_match(self.h, len(ids), float(header), ids.ctypes.data_as(IP), xy.ctypes.data_as(DP), pose_old.ctypes.data_as(DP))

    def loop_result(self):
        out = np.zeros(12)
        n = lib().vref_get_loop_result(self.h, out.ctypes.data_as(DP))
        return out, n

    def stage_seconds(self, reset=False):
        """wall-clock seconds in processImage (total), ceres::Solve, marginalisation, processIMU since creation / the last reset"""
        out = np.zeros(4)
        lib().vref_stage_seconds(self.h, out.ctypes.data_as(DP), int(reset))
        return dict(process_image=float(out[0]), ceres_solve=float(out[1]), marginalise=float(out[2]), process_imu=float(out[3]))

    def process_imu(self, dt, acc, gyr):
        a, g = _d(acc), _d(gyr)
        lib().vref_process_imu(self.h, float(dt), _abi.ptr(a, C.c_double), _abi.ptr(g, C.c_double))

    def set_init_window(self, P, Q, V, Ba, Bg):
        arrs = [_d(x) for x in (P, Q, V, Ba, Bg)]
        lib().vref_set_init_window(self.h, *[_abi.ptr(x, C.c_double) for x in arrs])

    def set_init_sfm(self, R, T):
        """ImageFrame::R / T of the window's W + 1 frames as solveInitial leaves them (VINS.cpp:889-905); consumed by the process_image
        call that fills the window: visualInitialAlign (VINS.cpp:1022-1102) + the first solve."""
        R, T = _d(R), _d(T)
        lib().vref_set_init_sfm(self.h, _abi.ptr(R, C.c_double), _abi.ptr(T, C.c_double))

    def set_init_sfm_frames(self, R, T):
        R, T = _d(R), _d(T)
        lib().vref_set_init_sfm_frames(self.h, R.shape[0], _abi.ptr(R, C.c_double), _abi.ptr(T, C.c_double))

    def init_frames(self):
        h = np.zeros(256)
        n = lib().vref_get_init_frames(self.h, 256, _abi.ptr(h, C.c_double))
        return h[:n].copy()

    def init_result(self):
        ok = np.zeros(1, np.int32); g = np.zeros(3); sc = np.zeros(1)
        lib().vref_get_init_result(self.h, _abi.ptr(ok, C.c_int32), _abi.ptr(g, C.c_double), _abi.ptr(sc, C.c_double))
        return int(ok[0]), g, float(sc[0])

    def process_image(self, ids, xyz, header):
        ids = np.ascontiguousarray(ids, np.int32)
        xyz = _d(xyz)
        return lib().vref_process_image(self.h, len(ids), _abi.ptr(ids, C.c_int32), _abi.ptr(xyz, C.c_double), float(header))

    def state(self):
        n = self.W + 1
        P, Q, V, Ba, Bg, H = (np.zeros((n, k)) for k in (3, 4, 3, 3, 3, 1))
        lib().vref_get_state(self.h, *[_abi.ptr(x, C.c_double) for x in (P, Q, V, Ba, Bg, H)])
        return dict(P=P, Q=Q, V=V, Ba=Ba, Bg=Bg, headers=H[:, 0])

    def post_solve(self):
        out = np.zeros((self.W + 1, 16))
        rc = lib().vref_get_post_solve(self.h, _abi.ptr(out, C.c_double))
        return None if rc else out

    def info(self):
        i = np.zeros(8, np.int32)
        d = np.zeros(4)
        lib().vref_get_info(self.h, _abi.ptr(i, C.c_int32), _abi.ptr(d, C.c_double))
        return dict(solver_flag=int(i[0]), marg_flag=int(i[1]), frame_count=int(i[2]), failure=int(i[3]),
                    n_feat=int(i[4]), n_proj=int(i[5]), iters=int(i[6]), last_track_num=int(i[7]),
                    cost0=float(d[0]), cost1=float(d[1]), prior_n=int(d[2]))

    def features(self, cap=4096):
        n = C.c_int(0)
        ids, st, no, fl = (np.zeros(cap, np.int32) for _ in range(4))
        dep = np.zeros(cap)
        lib().vref_get_features(self.h, cap, C.byref(n), _abi.ptr(ids, C.c_int32), _abi.ptr(st, C.c_int32),
                                _abi.ptr(no, C.c_int32), _abi.ptr(dep, C.c_double), _abi.ptr(fl, C.c_int32))
        k = n.value
        return dict(ids=ids[:k], start=st[:k], n_obs=no[:k], depth=dep[:k], solve_flag=fl[:k])

    def prior(self):
        N = 15 * (self.W + 1) + 6
        H = np.zeros((N, N)); b = np.zeros(N); pres = np.zeros(2 * (self.W + 1) + 1, np.int32); c0 = np.zeros(1)
        rc = lib().vref_get_prior(self.h, _abi.ptr(H, C.c_double), _abi.ptr(b, C.c_double), _abi.ptr(pres, C.c_int32),
                                  _abi.ptr(c0, C.c_double))
        return None if rc else dict(H=H, b=b, present=pres, c0=float(c0[0]))


def preintegrate(dt, acc, gyr, acc0, gyr0, ba, bg):
    dt, acc, gyr, acc0, gyr0, ba, bg = (_d(x) for x in (dt, acc, gyr, acc0, gyr0, ba, bg))
    pqv = np.zeros(10); jac = np.zeros((15, 15)); cov = np.zeros((15, 15)); sdt = np.zeros(1)
    p = lambda a: _abi.ptr(a, C.c_double)
    lib().vref_preintegrate(len(dt), p(dt), p(acc), p(gyr), p(acc0), p(gyr0), p(ba), p(bg), p(pqv), p(jac), p(cov), p(sdt))
    return pqv, jac, cov, float(sdt[0])


def imu_factor(pqv, jac, cov, sum_dt, lba, lbg, pi, sbi, pj, sbj):
    a = [_d(x) for x in (pqv, jac, cov)]
    b = [_d(x) for x in (lba, lbg, pi, sbi, pj, sbj)]
    res = np.zeros(15); J = np.zeros((15, 30))
    p = lambda x: _abi.ptr(x, C.c_double)
    lib().vref_imu_factor(p(a[0]), p(a[1]), p(a[2]), float(sum_dt), *[p(x) for x in b], p(res), p(J))
    return res, J


def imu_sqrt_info(cov):
    """LLT(cov^-1).matrixL()^T as imu_factor.h:72 computes it (Eigen 3.3, the oracle's compiler flags)."""
    cov = _d(cov); U = np.zeros((15, 15))
    lib().vref_imu_sqrt_info(_abi.ptr(cov, C.c_double), _abi.ptr(U, C.c_double))
    return U


def projection_factor(fx, tic, ric, pts_i, pts_j, pi, pj, inv_dep):
    a = [_d(x) for x in (tic, ric, pts_i, pts_j, pi, pj)]
    res = np.zeros(2); J = np.zeros((2, 13))
    p = lambda x: _abi.ptr(x, C.c_double)
    lib().vref_projection_factor(float(fx), *[p(x) for x in a], float(inv_dep), p(res), p(J))
    return res, J


class RefPnP:
    """The reference's motion-only PnP tracker (vinsPnP, vins_pnp.cpp; FeatureTracker::solveVinsPnP feeds it, feature_tracker.cpp:107-160)
    behind oracle/pnp_ref.cpp.  SURVEY.md section 8(f) rank 3: the oracle for the next widening step."""

    def __init__(self, cfg):
        self.h = lib().vpnp_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("vpnp_create failed")
        self.n = lib().vpnp_size() + 1

    def close(self):
        if self.h:
            lib().vpnp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_init(self, header, P, R, V, Ba, Bg):
        a = [_d(x) for x in (P, R, V, Ba, Bg)]
        lib().vpnp_set_init(self.h, float(header), *[_abi.ptr(x, C.c_double) for x in a])

    def process_imu(self, dt, acc, gyr):
        a, g = _d(acc), _d(gyr)
        lib().vpnp_process_imu(self.h, float(dt), _abi.ptr(a, C.c_double), _abi.ptr(g, C.c_double))

    def process_image(self, ids, obs_xy, pos_xyz, track_num, header, use_pnp=True):
        ids = np.ascontiguousarray(ids, np.int32); tn = np.ascontiguousarray(track_num, np.int32)
        o, p = _d(obs_xy), _d(pos_xyz)
        lib().vpnp_process_image(self.h, len(ids), _abi.ptr(ids, C.c_int32), _abi.ptr(o, C.c_double), _abi.ptr(p, C.c_double),
                                 _abi.ptr(tn, C.c_int32), float(header), int(use_pnp))

    def state(self):
        n = self.n
        P, V, Ba, Bg = (np.zeros((n, 3)) for _ in range(4))
        R = np.zeros((n, 3, 3)); H = np.zeros(n); fs = np.zeros(n, np.int32); fc = np.zeros(1, np.int32)
        lib().vpnp_get_state(self.h, *[_abi.ptr(x, C.c_double) for x in (P, R, V, Ba, Bg, H)], _abi.ptr(fs, C.c_int32), _abi.ptr(fc, C.c_int32))
        return dict(P=P, R=R, V=V, Ba=Ba, Bg=Bg, headers=H, find_solved=fs, frame_count=int(fc[0]))


def perspective_factor(obs_xy, pos_xyz, track_num, fx, pose7, ex7):
    """PerspectiveFactor::Evaluate (perspective_factor.cpp:16-67): residual (2,), d/dpose (2,6), d/dex_pose (2,6)."""
    res = np.zeros(2); Jp = np.zeros((2, 6)); Je = np.zeros((2, 6))
    a = [_d(x) for x in (obs_xy, pos_xyz)]
    b = [_d(x) for x in (pose7, ex7)]
    lib().vpnp_perspective_factor(_abi.ptr(a[0], C.c_double), _abi.ptr(a[1], C.c_double), int(track_num), float(fx), _abi.ptr(b[0], C.c_double),
                                  _abi.ptr(b[1], C.c_double), *[_abi.ptr(x, C.c_double) for x in (res, Jp, Je)])
    return res, Jp, Je


def visual_imu_align(n, R, T, counts, imu0, imu, bg0, tic):
    """The reference's VisualIMUAlignment (initial_aligment.cpp:222-229, compiled unmodified) on ONE stream.
    R [n][3][3], T [n][3], counts [n], imu0 [n][6], imu [n][M][7], bg0 [3], tic [3] -> (bgs [3], g [3], x [3 n + 3], ok)."""
    R, T, imu0, imu, bg0, tic = (_d(a) for a in (R, T, imu0, imu, bg0, tic))
    counts = np.ascontiguousarray(counts, np.int32)
    bgs = np.zeros(3); g = np.zeros(3); x = np.zeros(3 * n + 3)
    p = lambda a: a.ctypes.data_as(DP)
    ok = lib().vref_visual_imu_align(n, p(R), p(T), counts.ctypes.data_as(IP), p(imu0), imu.shape[1], p(imu), p(bg0), p(tic), p(bgs), p(g), p(x))
    return bgs, g, x, int(ok)
