"""ctypes mirror of include/vio_b200.h (struct vio_config) -- shared by the product wrapper and by the
test-only oracle loader so both sides are configured from the same bytes."""
import ctypes as C

import numpy as np


class VioConfig(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("tic", C.c_double * 3), ("ric", C.c_double * 9),
        ("max_cnt", C.c_int32), ("min_dist", C.c_int32),
        ("f_threshold", C.c_double),
        ("freq", C.c_int32), ("window_size", C.c_int32), ("num_of_f", C.c_int32),
        ("acc_n", C.c_double), ("acc_w", C.c_double), ("gyr_n", C.c_double), ("gyr_w", C.c_double),
        ("gravity", C.c_double),
        ("max_iters", C.c_int32),
        ("min_parallax", C.c_double), ("init_depth", C.c_double),
        ("max_imu_per_frame", C.c_int32), ("batch", C.c_int32), ("device", C.c_int32),
        ("marg_mode", C.c_int32), ("marg_eig", C.c_int32), ("marg_amm_eig", C.c_int32), ("solve_path", C.c_int32), ("be_threads", C.c_int32),
        ("loop_closure", C.c_int32),
    ]


def default_config(batch=1, max_cnt=150, window_size=10, rows=640, cols=480, freq=3, device=0) -> VioConfig:
    """iPhone7P entry of setGlobalParam (global_param.cpp:26-39) + BASELINE.json bench values.
    For other resolutions the intrinsics scale with the image width."""
    c = VioConfig()
    s = cols / 480.0
    c.rows, c.cols = rows, cols
    c.fx, c.fy, c.cx, c.cy = 526.600 * s, 526.678 * s, 243.481 * s, 315.280 * (rows / 640.0)
    c.tic[:] = [0.0, 0.092, 0.01]
    c.ric[:] = [1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, -1.0]      # ypr2R(0,0,180 deg)
    c.max_cnt, c.min_dist, c.f_threshold = max_cnt, 30, 1.0
    c.freq, c.window_size, c.num_of_f = freq, window_size, 1000
    c.acc_n, c.acc_w, c.gyr_n, c.gyr_w, c.gravity = 0.5, 0.002, 0.2, 4.0e-5, 9.805
    c.max_iters = 10
    c.min_parallax, c.init_depth = 10.0 / 549.0, 5.0
    c.max_imu_per_frame = 256
    c.batch, c.device = batch, device
    return c


def ptr(a, t):
    """numpy array -> ctypes pointer of type t (array must be C-contiguous and of the matching dtype)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(t))


DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int32)
FP = C.POINTER(C.c_float)
UP = C.POINTER(C.c_uint8)
