#!/usr/bin/env python
"""Replay dumped front-end divergence cases (tools/fe_parity_long.py) primitive by primitive on the GPU against cv2."""
import glob, importlib, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import frontend_oracle as fo
abi = importlib.import_module("vins-mobile_b200.abi")
api = importlib.import_module("vins-mobile_b200.api")
cfg = abi.default_config(batch=1, max_cnt=200)
for f in sorted(glob.glob(sys.argv[1] + "/*.npz")):
    d = np.load(f); k = int(re.search(r"_f(\d+)", f).group(1))
    pts = d["cur_pts"]
    nc, sc = fo.cv2_lk_track(d["img_prev"], d["img"], pts)
    ng, sg = api.prim_lk(cfg, d["img_prev"], d["img"], pts)
    ok = sc == 1
    bad = np.where((sg != sc) | (ok & (ng.view(np.uint32) != nc.view(np.uint32)).any(1)))[0]
    print(os.path.basename(f), "LK mismatches:", len(bad))
    for i in bad:
        print("    idx", i, "cur", pts[i], "cv2", nc[i], sc[i], "gpu", ng[i], sg[i])
    st = sc & fo.r_in_border(nc, 640, 480).astype(np.uint8)
    p1, p2 = pts[st == 1], nc[st == 1]
    mc = fo.cv2_find_fundamental(p1, p2)
    mg, it = api.prim_ransac_f(cfg, p1, p2)
    print("    F1 masks equal:", mg is not None and np.array_equal(mg, mc), "iters", it, "n", len(p1), "cv inl", int(mc.sum()), "gpu inl", None if mg is None else int(mg.sum()))
