"""ctypes binding of libvio_b200.so -- host-side mirror of the reference call surface
(FeatureTracker::readImage, VINS::processIMU / processImage / solve_ceres) in batched form.

There is NO CPU fallback here: if the CUDA library is missing or no device is present every call
raises.  Nothing in this module imports or calls anything under oracle/."""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import DP, FP, IP, UP, VioConfig, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("VIO_LIB_NAME", "libvio_b200.so"))      # VIO_LIB_NAME: debug builds (tools/)
_lib = None


class VioError(RuntimeError):
    pass


_CODES = {1: "VIO_ERR_ARG", 2: "VIO_ERR_CUDA", 3: "VIO_ERR_STATE", 4: "VIO_ERR_CAPACITY"}


def _check(rc, what):
    if rc != 0:
        raise VioError(f"{what} failed: {_CODES.get(rc, rc)}")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VioError(f"{LIB_PATH} not built -- run __graft_entry__.build(); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        cfgp = C.POINTER(VioConfig)
        L.vio_config_default.argtypes = [cfgp]
        L.vio_frontend_create.argtypes = [cfgp, C.POINTER(vp)]
        L.vio_frontend_destroy.argtypes = [vp]
        L.vio_frontend_read_images.argtypes = [vp, vp, C.POINTER(C.c_int)]
        L.vio_frontend_read_images_dev.argtypes = [vp, vp, C.POINTER(C.c_int)]
        L.vio_frontend_next_image_buffer.argtypes = [vp]
        L.vio_frontend_next_image_buffer.restype = vp
        L.vio_frontend_get_stream.argtypes = [vp, C.c_int, C.POINTER(C.c_int), IP, FP, IP, DP]
        L.vio_frontend_get_ui.argtypes = [vp, C.c_int, C.POINTER(C.c_int), FP, DP]
        L.vio_frontend_get_stats.argtypes = [vp, C.c_int, IP]
        L.vio_frontend_image_msg_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.vio_frontend_launch_count.argtypes = [vp]
        L.vio_frontend_launch_count.restype = C.c_int64
        L.vio_frontend_sync.argtypes = [vp]
        L.vio_frontend_stream.argtypes = [vp]
        L.vio_frontend_stream.restype = vp
        L.vio_frontend_use_stream.argtypes = [vp, vp]
        L.vio_frontend_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        L.vio_backend_solve.argtypes = [vp]
        L.vio_backend_set_loop_match.argtypes = [vp, IP, DP, IP, DP, DP]
        L.vio_backend_get_loop_result.argtypes = [vp, C.c_int, DP, C.POINTER(C.c_int32)]
        L.vio_backend_get_error.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        L.vio_prim_pyramid.argtypes = [cfgp, UP, UP, UP, UP]
        L.vio_frontend_set_clahe.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int]
        L.vio_prim_clahe.argtypes = [cfgp, UP, C.c_double, C.c_int, C.c_int, UP]
        L.vio_prim_min_eig_candidates.argtypes = [cfgp, UP, FP, C.c_int, C.c_int, FP, C.POINTER(C.c_int), FP]
        L.vio_prim_lk.argtypes = [cfgp, UP, UP, FP, C.c_int, FP, UP]
        L.vio_prim_ransac_f.argtypes = [cfgp, FP, FP, C.c_int, UP, C.POINTER(C.c_int)]
        if hasattr(L, "vio_pnp_create"):
            L.vio_pnp_create.argtypes = [cfgp, C.POINTER(vp)]
            L.vio_pnp_destroy.argtypes = [vp]
            L.vio_pnp_set_init.argtypes = [vp, DP, DP, DP, DP, DP, DP]
            L.vio_pnp_process_imu.argtypes = [vp, C.c_int, DP, DP, DP]
            L.vio_pnp_process_image.argtypes = [vp, IP, IP, DP, DP, IP, DP, C.c_int]
            L.vio_pnp_get_state.argtypes = [vp, C.c_int, DP, DP, DP, DP, IP, IP, DP]
            L.vio_pnp_launch_count.restype = C.c_int64
            L.vio_pnp_launch_count.argtypes = [vp]
        if hasattr(L, "vio_backend_create"):
            L.vio_backend_create.argtypes = [cfgp, C.POINTER(vp)]
            L.vio_backend_destroy.argtypes = [vp]
            L.vio_backend_clear.argtypes = [vp]
            L.vio_backend_process_imu.argtypes = [vp, C.c_int, DP, DP, DP]
            L.vio_backend_process_imu_dev.argtypes = [vp, C.c_int, vp, vp, vp]
            L.vio_backend_set_init_window.argtypes = [vp, DP, DP, DP, DP, DP]
            L.vio_backend_process_image.argtypes = [vp, IP, IP, DP, DP]
            L.vio_backend_process_image_dev.argtypes = [vp, vp, vp, vp, DP]
            L.vio_backend_process_image_from_frontend.argtypes = [vp, vp, DP]
            L.vio_backend_get_state.argtypes = [vp, C.c_int, DP, DP, DP, DP, DP, DP]
            L.vio_backend_state_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
            L.vio_backend_get_info.argtypes = [vp, C.c_int, IP, DP]
            L.vio_backend_get_features.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), IP, IP, IP, DP, IP]
            L.vio_backend_get_observations.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), DP]
            L.vio_backend_get_prior.argtypes = [vp, C.c_int, DP, DP, IP, DP]
            L.vio_backend_get_post_solve.argtypes = [vp, C.c_int, DP]
            L.vio_backend_launch_count.argtypes = [vp]
            L.vio_backend_launch_count.restype = C.c_int64
            L.vio_backend_sync.argtypes = [vp]
            L.vio_backend_use_stream.argtypes = [vp, vp]
            L.vio_backend_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
            L.vio_backend_copy_state.argtypes = [vp, vp, C.c_int]
            L.vio_backend_phase_cycles.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
            L.vio_prim_preintegrate.argtypes = [cfgp, C.c_int, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP]
            L.vio_prim_imu_factor.argtypes = [cfgp, DP, DP, DP, C.c_double, DP, DP, DP, DP, DP, DP, DP, DP]
            L.vio_prim_projection_factor.argtypes = [cfgp, DP, DP, DP, DP, C.c_double, DP, DP]
            L.vio_prim_imu_factor_sqi.argtypes = [cfgp, DP, DP, DP, C.c_double, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP]
            L.vio_backend_set_init_sfm.argtypes = [vp, DP, DP]
            L.vio_backend_set_init_sfm_frames.argtypes = [vp, IP, C.c_int, DP, DP]
            L.vio_backend_get_init_frames.argtypes = [vp, C.c_int, C.c_int, IP, DP]
            L.vio_backend_get_init_result.argtypes = [vp, C.c_int, IP, DP, DP]
            L.vio_visual_imu_align.argtypes = [cfgp, C.c_int, C.c_int, C.c_int, IP, DP, DP, IP, DP, DP, DP, DP, DP, DP, IP]
        _lib = L
    return _lib


def _parse_profile(txt):
    out = {}
    for item in txt.split(";"):
        if item:
            name, cnt, ms = item.split(":")
            out[name] = (int(cnt), float(ms))
    return out


def exported_symbols():
    """Names declared in include/vio_b200.h that the library must export."""
    hdr = os.path.join(_HERE, "..", "include", "vio_b200.h")
    import re
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(vio_[a-z0-9_]+)\s*\(", txt)))


class FrontEnd:
    """Batched FeatureTracker (feature_tracker.hpp:52-90): `batch` trackers advancing in lock-step."""

    def __init__(self, cfg: VioConfig):
        self.cfg = cfg
        self.B, self.maxp = cfg.batch, cfg.max_cnt
        self.h = C.c_void_p()
        _check(lib().vio_frontend_create(C.byref(cfg), C.byref(self.h)), "vio_frontend_create")

    def close(self):
        if getattr(self, "h", None):
            lib().vio_frontend_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read_images(self, images: np.ndarray) -> bool:
        """readImage for every stream; images (B, rows, cols) u8 HOST array.  Returns `published`."""
        images = np.ascontiguousarray(images, np.uint8)
        assert images.shape == (self.B, self.cfg.rows, self.cfg.cols)
        pub = C.c_int(0)
        _check(lib().vio_frontend_read_images(self.h, images.ctypes.data, C.byref(pub)), "vio_frontend_read_images")
        return bool(pub.value)

    def read_images_dev(self, dev_ptr: int) -> bool:
        pub = C.c_int(0)
        _check(lib().vio_frontend_read_images_dev(self.h, dev_ptr, C.byref(pub)), "vio_frontend_read_images_dev")
        return bool(pub.value)

    def next_image_buffer(self) -> int:
        return lib().vio_frontend_next_image_buffer(self.h)

    def stream(self, s: int) -> dict:
        n = C.c_int(0)
        ids = np.zeros(self.maxp, np.int32); pts = np.zeros((self.maxp, 2), np.float32)
        cnt = np.zeros(self.maxp, np.int32); xyz = np.zeros((self.maxp, 3), np.float64)
        _check(lib().vio_frontend_get_stream(self.h, s, C.byref(n), ptr(ids, C.c_int32), ptr(pts, C.c_float), ptr(cnt, C.c_int32),
                                             ptr(xyz, C.c_double)), "vio_frontend_get_stream")
        k = n.value
        return dict(n=k, ids=ids[:k], pts=pts[:k], track_cnt=cnt[:k], norm_xyz=xyz[:k])

    def ui(self, s: int):
        n = C.c_int(0)
        g = np.zeros((2 * self.maxp, 2), np.float32); t = np.zeros(2 * self.maxp)
        _check(lib().vio_frontend_get_ui(self.h, s, C.byref(n), ptr(g, C.c_float), ptr(t, C.c_double)), "vio_frontend_get_ui")
        return g[:n.value], t[:n.value]

    def stats(self, s: int) -> dict:
        st = np.zeros(8, np.int32)
        _check(lib().vio_frontend_get_stats(self.h, s, ptr(st, C.c_int32)), "vio_frontend_get_stats")
        keys = ["lk_in", "lk_ok", "f1_ok", "f2_ok", "kept", "new", "n_cand", "ransac_iters"]
        return dict(zip(keys, st.tolist()))

    def use_stream(self, cuda_stream: int):
        _check(lib().vio_frontend_use_stream(self.h, cuda_stream), "vio_frontend_use_stream")

    def set_clahe(self, enable=True, clip_limit=3.0, tiles_x=8, tiles_y=8):
        """cv::createCLAHE(); setClipLimit(3); apply() on every incoming frame (ViewController.mm:438-441)."""
        _check(lib().vio_frontend_set_clahe(self.h, int(enable), float(clip_limit), tiles_x, tiles_y), "vio_frontend_set_clahe")

    def profile(self, enable: bool) -> dict:
        buf = C.create_string_buffer(4096)
        _check(lib().vio_frontend_profile(self.h, int(enable), buf, 4096), "vio_frontend_profile")
        return _parse_profile(buf.value.decode())

    def image_msg_dev(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().vio_frontend_image_msg_dev(self.h, C.byref(a), C.byref(b), C.byref(c)), "vio_frontend_image_msg_dev")
        return a.value, b.value, c.value

    def launch_count(self) -> int:
        return lib().vio_frontend_launch_count(self.h)

    def sync(self):
        _check(lib().vio_frontend_sync(self.h), "vio_frontend_sync")


# ---- primitives (single image, host in/out) used by the parity tests -------------------------------------
def prim_pyramid(cfg, img):
    img = np.ascontiguousarray(img, np.uint8)
    r, c = img.shape
    outs = []
    for _ in range(3):
        r, c = (r + 1) // 2, (c + 1) // 2
        outs.append(np.zeros((r, c), np.uint8))
    _check(lib().vio_prim_pyramid(C.byref(cfg), ptr(img, C.c_uint8), *[ptr(o, C.c_uint8) for o in outs]), "vio_prim_pyramid")
    return outs


def prim_clahe(cfg, img, clip_limit=3.0, tiles_x=8, tiles_y=8):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    _check(lib().vio_prim_clahe(C.byref(cfg), ptr(img, C.c_uint8), float(clip_limit), tiles_x, tiles_y, ptr(out, C.c_uint8)), "vio_prim_clahe")
    return out


def prim_good_features(cfg, img, kept_xy, max_corners):
    img = np.ascontiguousarray(img, np.uint8)
    kept = np.ascontiguousarray(kept_xy, np.float32).reshape(-1, 2)
    out = np.zeros((max(max_corners, 1), 2), np.float32)
    n = C.c_int(0)
    mv = np.zeros(1, np.float32)
    _check(lib().vio_prim_min_eig_candidates(C.byref(cfg), ptr(img, C.c_uint8), ptr(kept, C.c_float), len(kept), max_corners,
                                             ptr(out, C.c_float), C.byref(n), ptr(mv, C.c_float)), "vio_prim_min_eig_candidates")
    return out[:n.value], float(mv[0])


def prim_lk(cfg, prev, nxt, pts):
    prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(pts); st = np.zeros(len(pts), np.uint8)
    _check(lib().vio_prim_lk(C.byref(cfg), ptr(prev, C.c_uint8), ptr(nxt, C.c_uint8), ptr(pts, C.c_float), len(pts), ptr(out, C.c_float),
                             ptr(st, C.c_uint8)), "vio_prim_lk")
    return out, st


def prim_ransac_f(cfg, p1, p2):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2); p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    m = np.zeros(len(p1), np.uint8); it = C.c_int(0)
    rc = lib().vio_prim_ransac_f(C.byref(cfg), ptr(p1, C.c_float), ptr(p2, C.c_float), len(p1), ptr(m, C.c_uint8), C.byref(it))
    if rc == 3:
        return None, it.value
    _check(rc, "vio_prim_ransac_f")
    return m, it.value


class BackEnd:
    """Batched VINS estimator (VINS.hpp:51-172): `batch` sliding windows advancing in lock-step."""

    def __init__(self, cfg: VioConfig):
        self.cfg = cfg
        self.B, self.W, self.maxp = cfg.batch, cfg.window_size, cfg.max_cnt
        self.h = C.c_void_p()
        _check(lib().vio_backend_create(C.byref(cfg), C.byref(self.h)), "vio_backend_create")

    def close(self):
        if getattr(self, "h", None):
            lib().vio_backend_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        _check(lib().vio_backend_clear(self.h), "vio_backend_clear")

    def process_imu(self, dt, acc, gyr):
        """processIMU for n consecutive samples: dt (n,B), acc (n,B,3), gyr (n,B,3) HOST arrays."""
        dt = np.ascontiguousarray(dt, np.float64).reshape(-1, self.B)
        n = dt.shape[0]
        acc = np.ascontiguousarray(acc, np.float64).reshape(n, self.B, 3)
        gyr = np.ascontiguousarray(gyr, np.float64).reshape(n, self.B, 3)
        _check(lib().vio_backend_process_imu(self.h, n, ptr(dt, C.c_double), ptr(acc, C.c_double), ptr(gyr, C.c_double)), "vio_backend_process_imu")

    def process_imu_dev(self, n, dt_ptr, acc_ptr, gyr_ptr):
        _check(lib().vio_backend_process_imu_dev(self.h, n, dt_ptr, acc_ptr, gyr_ptr), "vio_backend_process_imu_dev")

    def set_init_window(self, P, Q, V, Ba, Bg):
        n = self.W + 1
        P = np.ascontiguousarray(P, np.float64).reshape(self.B, n, 3); Q = np.ascontiguousarray(Q, np.float64).reshape(self.B, n, 4)
        V = np.ascontiguousarray(V, np.float64).reshape(self.B, n, 3)
        Ba = np.ascontiguousarray(Ba, np.float64).reshape(self.B, 3); Bg = np.ascontiguousarray(Bg, np.float64).reshape(self.B, 3)
        _check(lib().vio_backend_set_init_window(self.h, *[ptr(x, C.c_double) for x in (P, Q, V, Ba, Bg)]), "vio_backend_set_init_window")

    def process_image(self, counts, ids, xyz, headers):
        """processImage: counts (B,), ids (B,max_cnt), xyz (B,max_cnt,3), headers (B,) HOST arrays."""
        counts = np.ascontiguousarray(counts, np.int32).reshape(self.B)
        ids = np.ascontiguousarray(ids, np.int32).reshape(self.B, self.maxp)
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(self.B, self.maxp, 3)
        headers = np.ascontiguousarray(headers, np.float64).reshape(self.B)
        _check(lib().vio_backend_process_image(self.h, ptr(counts, C.c_int32), ptr(ids, C.c_int32), ptr(xyz, C.c_double), ptr(headers, C.c_double)),
               "vio_backend_process_image")

    def process_image_dev(self, counts_ptr, ids_ptr, xyz_ptr, headers):
        headers = np.ascontiguousarray(headers, np.float64).reshape(self.B)
        _check(lib().vio_backend_process_image_dev(self.h, counts_ptr, ids_ptr, xyz_ptr, ptr(headers, C.c_double)), "vio_backend_process_image_dev")

    def process_image_from_frontend(self, fe, headers):
        headers = np.ascontiguousarray(headers, np.float64).reshape(self.B)
        _check(lib().vio_backend_process_image_from_frontend(self.h, fe.h, ptr(headers, C.c_double)), "vio_backend_process_image_from_frontend")

    def process_image_single(self, ids, xyz, header):
        """convenience for batch 1: variable-length ids/xyz"""
        assert self.B == 1
        n = len(ids)
        I = np.zeros((1, self.maxp), np.int32); X = np.zeros((1, self.maxp, 3)); X[..., 2] = 1.0
        I[0, :n] = ids; X[0, :n] = xyz
        self.process_image([n], I, X, [header])

    def state(self, s=0):
        n = self.W + 1
        P, Q, V, Ba, Bg, H = (np.zeros((n, k)) for k in (3, 4, 3, 3, 3, 1))
        _check(lib().vio_backend_get_state(self.h, s, *[ptr(x, C.c_double) for x in (P, Q, V, Ba, Bg, H)]), "vio_backend_get_state")
        return dict(P=P, Q=Q, V=V, Ba=Ba, Bg=Bg, headers=H[:, 0])

    def post_solve(self, s=0):
        out = np.zeros((self.W + 1, 16))
        _check(lib().vio_backend_get_post_solve(self.h, s, ptr(out, C.c_double)), "vio_backend_get_post_solve")
        return out

    def state_dev(self):
        p = C.c_void_p(); n = C.c_int64(0)
        _check(lib().vio_backend_state_dev(self.h, C.byref(p), C.byref(n)), "vio_backend_state_dev")
        return p.value, n.value

    def info(self, s=0):
        i = np.zeros(8, np.int32); d = np.zeros(4)
        _check(lib().vio_backend_get_info(self.h, s, ptr(i, C.c_int32), ptr(d, C.c_double)), "vio_backend_get_info")
        return dict(solver_flag=int(i[0]), marg_flag=int(i[1]), frame_count=int(i[2]), failure=int(i[3]), n_feat=int(i[4]), n_proj=int(i[5]),
                    iters=int(i[6]), last_track_num=int(i[7]), cost0=float(d[0]), cost1=float(d[1]), prior_n=int(d[2]), err=int(d[3]) & 15,
                    marg_fast=(int(d[3]) >> 4) & 1, marg_sweeps=(int(d[3]) >> 5) & 127, marg_m=(int(d[3]) >> 12) & 1023, chol_retry=int(d[3]) >> 22)

    def features(self, s=0, cap=8192):
        n = C.c_int(0)
        ids, st, no, fl = (np.zeros(cap, np.int32) for _ in range(4)); dep = np.zeros(cap)
        _check(lib().vio_backend_get_features(self.h, s, cap, C.byref(n), ptr(ids, C.c_int32), ptr(st, C.c_int32), ptr(no, C.c_int32),
                                              ptr(dep, C.c_double), ptr(fl, C.c_int32)), "vio_backend_get_features")
        k = n.value
        return dict(ids=ids[:k], start=st[:k], n_obs=no[:k], depth=dep[:k], solve_flag=fl[:k])

    def observations(self, s=0, cap=8192):
        """feature_per_frame points (x, y) of every feature, [n][W + 1][2], rows beyond n_obs are stale"""
        n = C.c_int(0)
        obs = np.zeros((cap, self.W + 1, 2))
        _check(lib().vio_backend_get_observations(self.h, s, cap, C.byref(n), ptr(obs, C.c_double)), "vio_backend_get_observations")
        return obs[:n.value]

    def corresponding(self, l, r, s=0):
        """FeatureManager::getCorresponding(l, r) (feature_manager.cpp:157-176): (a, b) point pairs of the features seen in both frames"""
        f = self.features(s); o = self.observations(s)
        sel = (f["start"] <= l) & (f["start"] + f["n_obs"] - 1 >= r)
        idx = np.nonzero(sel)[0]
        a = np.stack([o[i, l - f["start"][i]] for i in idx]) if len(idx) else np.zeros((0, 2))
        b = np.stack([o[i, r - f["start"][i]] for i in idx]) if len(idx) else np.zeros((0, 2))
        return f["ids"][idx], a, b

    def prior(self, s=0):
        N = 15 * (self.W + 1) + 6
        H = np.zeros((N, N)); b = np.zeros(N); pres = np.zeros(2 * (self.W + 1) + 1, np.int32); c0 = np.zeros(1)
        rc = lib().vio_backend_get_prior(self.h, s, ptr(H, C.c_double), ptr(b, C.c_double), ptr(pres, C.c_int32), ptr(c0, C.c_double))
        if rc == 3:
            return None
        _check(rc, "vio_backend_get_prior")
        return dict(H=H, b=b, present=pres, c0=float(c0[0]))

    def use_stream(self, cuda_stream: int):
        _check(lib().vio_backend_use_stream(self.h, cuda_stream), "vio_backend_use_stream")

    def profile(self, enable: bool) -> dict:
        buf = C.create_string_buffer(4096)
        _check(lib().vio_backend_profile(self.h, int(enable), buf, 4096), "vio_backend_profile")
        return _parse_profile(buf.value.decode())

    def phase_cycles(self, reset=True):
        out = np.zeros((self.B, 32), np.int64)
        _check(lib().vio_backend_phase_cycles(self.h, out.ctypes.data_as(C.POINTER(C.c_longlong)), int(reset)), "vio_backend_phase_cycles")
        return out

    def copy_state(self, dst_ptr: int, is_device):
        """is_device: False/0 host (synchronous), True/1 device (stream-ordered), 2 pinned host (stream-ordered)."""
        _check(lib().vio_backend_copy_state(self.h, dst_ptr, int(is_device)), "vio_backend_copy_state")

    def state_all(self):
        out = np.zeros((self.B, self.W + 1, 16))
        self.copy_state(out.ctypes.data, False)
        return out

    def launch_count(self):
        return lib().vio_backend_launch_count(self.h)

    def sync(self):
        _check(lib().vio_backend_sync(self.h), "vio_backend_sync")

    def set_loop_match(self, counts, headers, ids, xy, pose_old):
        """retrive_pose_data per stream: counts (B,), headers (B,), ids (B,max_cnt) ascending, xy (B,max_cnt,2), pose_old (B,7) = P_old, Q_old xyzw"""
        counts = np.ascontiguousarray(counts, np.int32).reshape(self.B); headers = np.ascontiguousarray(headers, np.float64).reshape(self.B)
        ids = np.ascontiguousarray(ids, np.int32).reshape(self.B, self.maxp); xy = np.ascontiguousarray(xy, np.float64).reshape(self.B, self.maxp, 2)
        pose_old = np.ascontiguousarray(pose_old, np.float64).reshape(self.B, 7)
        _check(lib().vio_backend_set_loop_match(self.h, ptr(counts, C.c_int32), ptr(headers, C.c_double), ptr(ids, C.c_int32), ptr(xy, C.c_double),
                                                ptr(pose_old, C.c_double)), "vio_backend_set_loop_match")

    def loop_result(self, s=0):
        out = np.zeros(12); n = C.c_int32(0)
        _check(lib().vio_backend_get_loop_result(self.h, s, ptr(out, C.c_double), C.byref(n)), "vio_backend_get_loop_result")
        return out, n.value

    def solve(self):
        """VINS::solve_ceres() alone on the current window of every NON_LINEAR stream"""
        _check(lib().vio_backend_solve(self.h), "vio_backend_solve")

    def set_init_sfm(self, R, T):
        """ImageFrame::R [B][W+1][3][3] / T [B][W+1][3] after the global SfM (VINS.cpp:889-905); visualInitialAlign then runs on the device
        inside the process_image call that fills the window."""
        R = np.ascontiguousarray(R, np.float64); T = np.ascontiguousarray(T, np.float64)
        assert R.shape == (self.B, self.W + 1, 3, 3) and T.shape == (self.B, self.W + 1, 3)
        _check(lib().vio_backend_set_init_sfm(self.h, ptr(R, C.c_double), ptr(T, C.c_double)), "vio_backend_set_init_sfm")

    def set_init_sfm_frames(self, n_frames, R, T):
        """The general form: R [B][F][3][3] / T [B][F][3] of EVERY frame of all_image_frame (init_frames()), n_frames [B] of them per stream."""
        n_frames = np.ascontiguousarray(n_frames, np.int32)
        R = np.ascontiguousarray(R, np.float64); T = np.ascontiguousarray(T, np.float64)
        F = R.shape[1]
        assert R.shape == (self.B, F, 3, 3) and T.shape == (self.B, F, 3) and n_frames.shape == (self.B,)
        _check(lib().vio_backend_set_init_sfm_frames(self.h, ptr(n_frames, C.c_int32), F, ptr(R, C.c_double), ptr(T, C.c_double)), "vio_backend_set_init_sfm_frames")

    def init_frames(self, s=0):
        """headers of the frames in all_image_frame for stream s, oldest first"""
        cap = 3 * (self.W + 1)
        n = np.zeros(1, np.int32); h = np.zeros(cap)
        _check(lib().vio_backend_get_init_frames(self.h, s, cap, ptr(n, C.c_int32), ptr(h, C.c_double)), "vio_backend_get_init_frames")
        return h[:int(n[0])].copy()

    def init_result(self, s=0):
        ok = np.zeros(1, np.int32); g = np.zeros(3); sc = np.zeros(1)
        _check(lib().vio_backend_get_init_result(self.h, s, ptr(ok, C.c_int32), ptr(g, C.c_double), ptr(sc, C.c_double)), "vio_backend_get_init_result")
        return int(ok[0]), g, float(sc[0])

    def error(self, s=0, clear=False):
        """latched per-stream error code (VIO_ERR_CAPACITY ...), optionally cleared"""
        code = C.c_int32(0)
        _check(lib().vio_backend_get_error(self.h, s, int(clear), C.byref(code)), "vio_backend_get_error")
        return code.value


def prim_preintegrate(cfg, dt, acc, gyr, acc0, gyr0, ba, bg):
    d = lambda a: np.ascontiguousarray(a, np.float64)
    dt, acc, gyr, acc0, gyr0, ba, bg = (d(x) for x in (dt, acc, gyr, acc0, gyr0, ba, bg))
    pqv = np.zeros(10); jac = np.zeros((15, 15)); cov = np.zeros((15, 15)); sdt = np.zeros(1)
    p = lambda a: ptr(a, C.c_double)
    _check(lib().vio_prim_preintegrate(C.byref(cfg), len(dt), p(dt), p(acc), p(gyr), p(acc0), p(gyr0), p(ba), p(bg), p(pqv), p(jac), p(cov), p(sdt)),
           "vio_prim_preintegrate")
    return pqv, jac, cov, float(sdt[0])


def prim_imu_factor(cfg, pqv, jac, cov, sum_dt, lba, lbg, pi, sbi, pj, sbj):
    d = lambda a: np.ascontiguousarray(a, np.float64)
    a = [d(x) for x in (pqv, jac, cov)]; b = [d(x) for x in (lba, lbg, pi, sbi, pj, sbj)]
    res = np.zeros(15); J = np.zeros((15, 30))
    p = lambda x: ptr(x, C.c_double)
    _check(lib().vio_prim_imu_factor(C.byref(cfg), p(a[0]), p(a[1]), p(a[2]), float(sum_dt), *[p(x) for x in b], p(res), p(J)), "vio_prim_imu_factor")
    return res, J


def prim_imu_factor_sqi(cfg, pqv, jac, cov, sum_dt, lba, lbg, pi, sbi, pj, sbj, sqrt_info=None):
    """IMUFactor::Evaluate with the weighting matrix explicit: returns (residual, J, sqrt_info used); `sqrt_info` overrides the device's own."""
    d = lambda a: np.ascontiguousarray(a, np.float64)
    a = [d(x) for x in (pqv, jac, cov)]; b = [d(x) for x in (lba, lbg, pi, sbi, pj, sbj)]
    res = np.zeros(15); J = np.zeros((15, 30)); U = np.zeros((15, 15))
    p = lambda x: ptr(x, C.c_double)
    sin = d(sqrt_info) if sqrt_info is not None else None
    _check(lib().vio_prim_imu_factor_sqi(C.byref(cfg), p(a[0]), p(a[1]), p(a[2]), float(sum_dt), *[p(x) for x in b],
                                         p(sin) if sin is not None else None, p(U), p(res), p(J)), "vio_prim_imu_factor_sqi")
    return res, J, U


def prim_projection_factor(cfg, pts_i, pts_j, pi, pj, inv_dep):
    d = lambda a: np.ascontiguousarray(a, np.float64)
    a = [d(x) for x in (pts_i, pts_j, pi, pj)]
    res = np.zeros(2); J = np.zeros((2, 13))
    p = lambda x: ptr(x, C.c_double)
    _check(lib().vio_prim_projection_factor(C.byref(cfg), *[p(x) for x in a], float(inv_dep), p(res), p(J)), "vio_prim_projection_factor")
    return res, J


def visual_imu_align(cfg, n_frames, R, T, imu_counts, imu0, imu, bg0):
    """Batched VisualIMUAlignment (initial_aligment.cpp:222-229).  R [B][F][3][3], T [B][F][3], imu_counts [B][F], imu0 [B][F][6],
    imu [B][F][M][7] (dt, acc, gyr), bg0 [B][3]  ->  (bgs [B][3], g [B][3], x [B][3 F + 4], ok [B])."""
    d = lambda a: np.ascontiguousarray(a, np.float64)
    n_frames = np.ascontiguousarray(n_frames, np.int32); imu_counts = np.ascontiguousarray(imu_counts, np.int32)
    R, T, imu0, imu, bg0 = d(R), d(T), d(imu0), d(imu), d(bg0)
    B, F = imu_counts.shape
    M = imu.shape[2]
    assert R.shape == (B, F, 3, 3) and T.shape == (B, F, 3) and imu0.shape == (B, F, 6) and imu.shape == (B, F, M, 7) and bg0.shape == (B, 3)
    bgs = np.zeros((B, 3)); g = np.zeros((B, 3)); x = np.zeros((B, 3 * F + 4)); ok = np.zeros(B, np.int32)
    p = lambda a: ptr(a, C.c_double)
    _check(lib().vio_visual_imu_align(C.byref(cfg), B, F, M, ptr(n_frames, C.c_int32), p(R), p(T), ptr(imu_counts, C.c_int32), p(imu0), p(imu),
                                      p(bg0), p(bgs), p(g), p(x), ptr(ok, C.c_int32)), "vio_visual_imu_align")
    return bgs, g, x, ok


class PnP:
    """Batched vinsPnP (vins_pnp.hpp:40-91): the motion-only tracker FeatureTracker::solveVinsPnP drives (feature_tracker.cpp:107-160)."""
    N = 7

    def __init__(self, cfg: VioConfig):
        self.cfg, self.B, self.maxp = cfg, cfg.batch, cfg.max_cnt
        self.h = C.c_void_p()
        _check(lib().vio_pnp_create(C.byref(cfg), C.byref(self.h)), "vio_pnp_create")

    def close(self):
        if getattr(self, "h", None):
            lib().vio_pnp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_init(self, header, P, R, V, Ba, Bg):
        d = lambda a, *shape: np.ascontiguousarray(a, np.float64).reshape(self.B, *shape)
        a = [d(header), d(P, 3), d(R, 9), d(V, 3), d(Ba, 3), d(Bg, 3)]
        _check(lib().vio_pnp_set_init(self.h, *[ptr(x, C.c_double) for x in a]), "vio_pnp_set_init")

    def process_imu(self, dt, acc, gyr):
        dt = np.ascontiguousarray(dt, np.float64).reshape(-1, self.B)
        n = dt.shape[0]
        acc = np.ascontiguousarray(acc, np.float64).reshape(n, self.B, 3); gyr = np.ascontiguousarray(gyr, np.float64).reshape(n, self.B, 3)
        _check(lib().vio_pnp_process_imu(self.h, n, ptr(dt, C.c_double), ptr(acc, C.c_double), ptr(gyr, C.c_double)), "vio_pnp_process_imu")

    def process_image(self, counts, ids, obs_xy, pos_xyz, track_num, headers, use_pnp=True):
        counts = np.ascontiguousarray(counts, np.int32).reshape(self.B)
        ids = np.ascontiguousarray(ids, np.int32).reshape(self.B, self.maxp); tn = np.ascontiguousarray(track_num, np.int32).reshape(self.B, self.maxp)
        obs = np.ascontiguousarray(obs_xy, np.float64).reshape(self.B, self.maxp, 2); pos = np.ascontiguousarray(pos_xyz, np.float64).reshape(self.B, self.maxp, 3)
        hdr = np.ascontiguousarray(headers, np.float64).reshape(self.B)
        _check(lib().vio_pnp_process_image(self.h, ptr(counts, C.c_int32), ptr(ids, C.c_int32), ptr(obs, C.c_double), ptr(pos, C.c_double),
                                           ptr(tn, C.c_int32), ptr(hdr, C.c_double), int(use_pnp)), "vio_pnp_process_image")

    def state(self, s=0):
        n = self.N
        P, V = np.zeros((n, 3)), np.zeros((n, 3))
        R = np.zeros((n, 3, 3)); H = np.zeros(n); fs = np.zeros(n, np.int32); info = np.zeros(3, np.int32); cost = np.zeros(2)
        _check(lib().vio_pnp_get_state(self.h, s, ptr(P, C.c_double), ptr(R, C.c_double), ptr(V, C.c_double), ptr(H, C.c_double), ptr(fs, C.c_int32),
                                       ptr(info, C.c_int32), ptr(cost, C.c_double)), "vio_pnp_get_state")
        return dict(P=P, R=R, V=V, headers=H, find_solved=fs, frame_count=int(info[0]), err=int(info[1]), iters=int(info[2]), cost0=float(cost[0]),
                    cost1=float(cost[1]))
