"""Diagnostic: the same streams through a small-batch and a large-batch pipeline in lock step (sync after every frame); reports the
first frame at which a stream's front-end or back-end result differs between the two."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B1, B2 = int(sys.argv[1]), int(sys.argv[2])
n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 60
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B2, n_frames + 3, 0, "cuda:0")
pipes = []
for B in (B1, B2):
    cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
    f = frames[:, :B].contiguous()
    d, a, g = (torch.as_tensor(np.ascontiguousarray(x), device="cuda:0") for x in (dt[:, :, :B], acc[:, :, :B], gyr[:, :, :B]))
    st = torch.cuda.Stream()
    pipes.append((bench.Pipeline(api, cfg, st.cuda_stream, st.cuda_stream, gt[:B], False), f, (d, a, g), st))
reported = set()
for i in range(n_frames):
    res = []
    for pipe, f, (d, a, g), st in pipes:
        with torch.cuda.stream(st):
            pub = pipe.step(f[i].data_ptr(), lambda k: (d[k].data_ptr(), a[k].data_ptr(), g[k].data_ptr()))
            torch.cuda.synchronize()
        fe = [pipe.fe.stream(b) for b in range(B1)]
        be = [dict(pipe.be.info(b), feat=pipe.be.features(b)) for b in range(B1)] if pub else None
        stt = pipe.be.state_all()[:B1] if pub else None
        res.append((fe, be, stt, pub))
    (fe1, be1, st1, pub), (fe2, be2, st2, _) = res
    for b in range(B1):
        if b in reported:
            continue
        why = None
        if not np.array_equal(fe1[b]["ids"], fe2[b]["ids"]) or not np.array_equal(fe1[b]["pts"].view(np.uint32), fe2[b]["pts"].view(np.uint32)):
            why = f"FE differs (n {len(fe1[b]['ids'])} vs {len(fe2[b]['ids'])})"
        elif pub:
            x, y = be1[b], be2[b]
            fx, fy = x.pop("feat"), y.pop("feat")
            if len(fx["ids"]) != len(fy["ids"]) or any(not np.array_equal(fx[k], fy[k]) for k in ("ids", "start", "n_obs")):
                common = min(len(fx["ids"]), len(fy["ids"]))
                idx = [q for q in range(common) if fx["ids"][q] != fy["ids"][q] or fx["start"][q] != fy["start"][q] or fx["n_obs"][q] != fy["n_obs"][q]]
                why = f"feature table differs: n {len(fx['ids'])} vs {len(fy['ids'])}; first rows " + str([(q, int(fx['ids'][q]), int(fx['start'][q]), int(fx['n_obs'][q]), int(fy['ids'][q]), int(fy['start'][q]), int(fy['n_obs'][q])) for q in idx[:6]])
                why += f" | fe n {len(fe1[b]['ids'])} {len(fe2[b]['ids'])} track_cnt equal {np.array_equal(fe1[b]['track_cnt'], fe2[b]['track_cnt'])} xyz equal {np.array_equal(fe1[b]['norm_xyz'], fe2[b]['norm_xyz'])}"
            if why:
                pass
            elif x["n_feat"] != y["n_feat"] or x["n_proj"] != y["n_proj"] or x["iters"] != y["iters"] or not (abs(x["cost0"] - y["cost0"]) <= 1e-6 * abs(x["cost0"])):
                why = f"BE differs: {x} vs {y}"
            elif not np.allclose(st1[b], st2[b], rtol=1e-6, atol=1e-9):
                why = f"BE state differs max {np.abs(st1[b] - st2[b]).max():.3e}"
        if why:
            reported.add(b)
            print(f"frame {i} (kf {pipes[0][0].kf - 1}) stream {b}: {why}"[:600], flush=True)
print("done; streams that diverged:", sorted(reported))
