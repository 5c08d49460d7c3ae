"""Diagnostic: SfM-based initialisation, GPU vs reference, with / without a rejected first attempt.  python tools/init_sfm_probe.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import backend_oracle as bo  # noqa: E402
import be_common  # noqa: E402
from be_common import Quiet, drive_sfm, rel_err, sfm_window  # noqa: E402

api = importlib.import_module("vins-mobile_b200.api")
abi = importlib.import_module("vins-mobile_b200.abi")
synth = importlib.import_module("vins-mobile_b200.synth")
cfg = abi.default_config(batch=1, max_cnt=150)
W = cfg.window_size
import copy
for first_bad, iters in ((False, 10),):
    cfg.max_iters = iters
    tr = synth.make_tracks(0, 20, max_cnt=cfg.max_cnt)
    ref, gpu = bo.RefEstimator(cfg), api.BackEnd(cfg)
    for k in range(14):
        sfm = None
        if k == W:
            sfm = sfm_window(tr, k, W, mirrored=first_bad)
        if k == W + 1 and first_bad:
            sfm = sfm_window(tr, k, W)
        with Quiet():
            drive_sfm(ref, tr, k, W, sfm)
        drive_sfm(gpu, tr, k, W, sfm)
        if k >= W:
            rs, gs = ref.state(), gpu.state()
            print(f"first_bad={first_bad} kf {k}: ok {ref.init_result()[0]} {gpu.init_result()[0]} scale {ref.init_result()[2]:.9f} {gpu.init_result()[2]:.9f} "
                  f"P {rel_err(gs['P'], rs['P']):.2e} V {rel_err(gs['V'], rs['V']):.2e} Bg {np.abs(gs['Bg'] - rs['Bg']).max():.2e} "
                  f"cost {ref.info()['cost1']:.6f} {gpu.info()['cost1']:.6f}")
            rp, gp = ref.post_solve(), gpu.post_solve()
            if rp is not None and gp is not None:
                rp = np.asarray(rp).reshape(W + 1, 16); gp = np.asarray(gp).reshape(W + 1, 16)
                print(f"   iters={iters} post_solve rel err: P {rel_err(gp[:, :3], rp[:, :3]):.2e} Q {np.abs(gp[:, 3:7] - rp[:, 3:7]).max():.2e} V {rel_err(gp[:, 7:10], rp[:, 7:10]):.2e} "
                      f"Ba {np.abs(gp[:, 10:13] - rp[:, 10:13]).max():.2e} Bg {np.abs(gp[:, 13:] - rp[:, 13:]).max():.2e}")
                if k == W:
                    print("   ref P[1], Q[0]:", rp[1, :3], rp[0, 3:7], "\n   gpu P[1], Q[0]:", gp[1, :3], gp[0, 3:7])
                    rf, gf = ref.features(), gpu.features()
                    m = min(len(rf["depth"]), len(gf["depth"]))
                    print("   depth rel err", rel_err(gf["depth"][:m], rf["depth"][:m]), "ids equal", np.array_equal(rf["ids"], gf["ids"]))
    ref.close(); gpu.close()
