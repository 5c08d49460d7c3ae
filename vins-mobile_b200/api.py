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
LIB_PATH = os.path.join(_HERE, "libvio_b200.so")
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
        L.vio_prim_pyramid.argtypes = [cfgp, UP, UP, UP, UP]
        L.vio_prim_min_eig_candidates.argtypes = [cfgp, UP, FP, C.c_int, C.c_int, FP, C.POINTER(C.c_int), FP]
        L.vio_prim_lk.argtypes = [cfgp, UP, UP, FP, C.c_int, FP, UP]
        L.vio_prim_ransac_f.argtypes = [cfgp, FP, FP, C.c_int, UP, C.POINTER(C.c_int)]
        if hasattr(L, "vio_backend_create"):
            L.vio_backend_create.argtypes = [cfgp, C.POINTER(vp)]
            L.vio_backend_destroy.argtypes = [vp]
            L.vio_backend_clear.argtypes = [vp]
            L.vio_backend_process_imu.argtypes = [vp, C.c_int, DP, DP, DP]
            L.vio_backend_process_imu_dev.argtypes = [vp, C.c_int, vp, vp, vp]
            L.vio_backend_set_init_window.argtypes = [vp, DP, DP, DP, DP, DP]
            L.vio_backend_process_image.argtypes = [vp, IP, IP, DP, DP]
            L.vio_backend_process_image_dev.argtypes = [vp, vp, vp, vp, DP]
            L.vio_backend_get_state.argtypes = [vp, C.c_int, DP, DP, DP, DP, DP, DP]
            L.vio_backend_state_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
            L.vio_backend_get_info.argtypes = [vp, C.c_int, IP, DP]
            L.vio_backend_get_features.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), IP, IP, IP, DP, IP]
            L.vio_backend_get_prior.argtypes = [vp, C.c_int, DP, DP, IP, DP]
            L.vio_backend_get_post_solve.argtypes = [vp, C.c_int, DP]
            L.vio_backend_launch_count.argtypes = [vp]
            L.vio_backend_launch_count.restype = C.c_int64
            L.vio_backend_sync.argtypes = [vp]
            L.vio_prim_preintegrate.argtypes = [cfgp, C.c_int, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP, DP]
            L.vio_prim_imu_factor.argtypes = [cfgp, DP, DP, DP, C.c_double, DP, DP, DP, DP, DP, DP, DP, DP]
            L.vio_prim_projection_factor.argtypes = [cfgp, DP, DP, DP, DP, C.c_double, DP, DP]
        _lib = L
    return _lib


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
        g = np.zeros((self.maxp, 2), np.float32); t = np.zeros(self.maxp)
        _check(lib().vio_frontend_get_ui(self.h, s, C.byref(n), ptr(g, C.c_float), ptr(t, C.c_double)), "vio_frontend_get_ui")
        return g[:n.value], t[:n.value]

    def stats(self, s: int) -> dict:
        st = np.zeros(8, np.int32)
        _check(lib().vio_frontend_get_stats(self.h, s, ptr(st, C.c_int32)), "vio_frontend_get_stats")
        keys = ["lk_in", "lk_ok", "f1_ok", "f2_ok", "kept", "new", "n_cand", "ransac_iters"]
        return dict(zip(keys, st.tolist()))

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
