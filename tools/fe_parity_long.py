#!/usr/bin/env python
"""Front-end parity over long streams: the CUDA tracker (one batch-NB handle) against NB cv2-backed FeatureTracker restatements
(oracle/frontend_oracle.py, backend="cv2": KLT / RANSAC-F / goodFeaturesToTrack are the OpenCV binary), NF frames each.
Prints the per-stream first-divergence table (ids / bitwise positions / track counts) and writes it to OUT (json); the first
diverging frame of each stream is dumped (images k-1, k, points) next to it for offline analysis.

    python tools/fe_parity_long.py [NB=8] [NF=300] [OUT=gpurun_out/fe_parity_long.json] [device=cuda]
"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import frontend_oracle as fo
abi = importlib.import_module("vins-mobile_b200.abi")
api = importlib.import_module("vins-mobile_b200.api")
synth = importlib.import_module("vins-mobile_b200.synth")

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 300
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "fe_parity_long.json")
dev = sys.argv[4] if len(sys.argv) > 4 else "cuda"
os.makedirs(os.path.dirname(out), exist_ok=True)
streams = [synth.make_stream(i, nf, device=dev) for i in range(nb)]
fe = api.FrontEnd(abi.default_config(batch=nb, max_cnt=150))
trs = [fo.FeatureTrackerOracle(max_cnt=150, backend="cv2") for _ in range(nb)]
first = {b: None for b in range(nb)}
prev = None
prev_pts = [None] * nb
for k in range(nf):
    ims = np.stack([s.images[k].cpu().numpy() for s in streams])
    fe.read_images(ims)
    for b, tr in enumerate(trs):
        if first[b] is not None:
            continue
        cur_before = tr.cur_pts.copy()
        tr.read_image(ims[b])
        g = fe.stream(b)
        ids_ok = np.array_equal(g["ids"], tr.ids)
        pts_ok = ids_ok and np.array_equal(g["pts"].view(np.uint32), tr.cur_pts.view(np.uint32))
        cnt_ok = ids_ok and np.array_equal(g["track_cnt"], tr.track_cnt)
        if not (ids_ok and pts_ok and cnt_ok):
            nd = int((g["pts"].view(np.uint32) != tr.cur_pts.view(np.uint32)).any(1).sum()) if ids_ok else -1
            first[b] = {"frame": k, "ids_equal": bool(ids_ok), "points_bitwise_equal": bool(pts_ok), "n_points_differ": nd,
                        "max_abs_diff_px": float(np.abs(g["pts"] - tr.cur_pts).max()) if ids_ok else None, "n_gpu": int(len(g["ids"])), "n_cv2": int(len(tr.ids))}
            np.savez_compressed(out.replace(".json", f"_case_s{b}_f{k}.npz"), img_prev=prev[b], img=ims[b], cur_pts=cur_before,
                                gpu_ids=g["ids"], gpu_pts=g["pts"], cv_ids=tr.ids, cv_pts=tr.cur_pts)
    prev = ims
table = {"streams": nb, "frames": nf, "first_divergence": {str(b): first[b] for b in range(nb)},
         "identical_streams": int(sum(v is None for v in first.values()))}
json.dump(table, open(out, "w"), indent=1)
print(json.dumps(table))
