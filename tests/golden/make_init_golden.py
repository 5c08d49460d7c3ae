"""Generates tests/golden/init_sfm_golden.npz: the window the reference oracle (oracle/_ref: restated visualInitialAlign around the
reference's UNMODIFIED VisualIMUAlignment + reference factors + Ceres) holds right after an initialisation from SfM poses, for
(a) stream 0: a mirrored (rejected) attempt at keyframe W, an accepted one at W + 1;  (b) stream 1: a non-keyframe in all_image_frame.
Run in the build container:  python tests/golden/make_init_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import backend_oracle as bo  # noqa: E402
import be_common  # noqa: E402
from be_common import Quiet, init_scenario  # noqa: E402

abi = importlib.import_module("vins-mobile_b200.abi")
cfg = abi.default_config(batch=1, max_cnt=150)
out = {}
for name in ("rejected_then_accepted", "non_keyframe"):
    ref = bo.RefEstimator(cfg)
    with Quiet():
        info = init_scenario(ref, cfg, name)
    st = ref.state()
    ok, g, sc = ref.init_result()
    assert ok == 1 and ref.info()["solver_flag"] == 1
    for k in ("P", "Q", "V", "Bg"):
        out[f"{name}_{k}"] = st[k]
    out[f"{name}_headers"] = st["headers"]; out[f"{name}_scale"] = sc; out[f"{name}_g"] = g; out[f"{name}_cost1"] = ref.info()["cost1"]
    ref.close()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "init_sfm_golden.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})
