"""Generates tests/golden/align_golden.npz: inputs of VisualIMUAlignment for two synthetic streams and the outputs of the reference's
UNMODIFIED initial_aligment.cpp (through oracle/_ref/libvins_ref.so, built by oracle/Makefile from /root/reference).
Run in the build container:  python tests/golden/make_align_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import be_common  # noqa: E402
from oracle import backend_oracle as bo  # noqa: E402

out = {}
for i, (sid, n) in enumerate(((7, 11), (8, 16))):
    c = be_common.align_case(sid, n)
    bg0 = np.array([0.001, 0.0, -0.001]) * i
    bgs, g, x, ok = bo.visual_imu_align(n, c["R"], c["T"], c["counts"], c["imu0"], c["imu"], bg0, c["tic"])
    for k in ("R", "T", "counts", "imu0", "imu", "tic"):
        out[f"{k}{i}"] = c[k]
    out[f"bg0{i}"] = bg0; out[f"bgs{i}"] = bgs; out[f"g{i}"] = g; out[f"x{i}"] = x; out[f"ok{i}"] = np.int32(ok)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "align_golden.npz"), **out)
print("written", {k: np.shape(v) for k, v in out.items() if k[0] in "xg"})
