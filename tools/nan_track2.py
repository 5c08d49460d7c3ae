"""Diagnostic: rep 0 syncs after every keyframe (reference); later reps run un-synced up to keyframe S and then compare cost0 of all streams."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B, n_frames = 128, 75
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames + 3, 0, "cuda:0")
dt_d, acc_d, gyr_d = (torch.as_tensor(x, device="cuda:0").contiguous() for x in (dt, acc, gyr))
imu_dev = lambda k: (dt_d[k].data_ptr(), acc_d[k].data_ptr(), gyr_d[k].data_ptr())
ref = {}
for rep, S in enumerate([0, 11, 11, 12, 12, 13, 13, 14, 14, 16, 16, 20, 20]):
    s_fe = torch.cuda.Stream()
    pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_fe.cuda_stream, gt, False)
    msg = ""
    with torch.cuda.stream(s_fe):
        for i in range(n_frames):
            pub = pipe.step(frames[i].data_ptr(), imu_dev)
            k = pipe.kf - 1
            if pub and k >= 10 and (S == 0 or k == S):
                torch.cuda.synchronize()
                info = [pipe.be.info(b) for b in range(B)]
                c0 = np.array([x['cost0'] for x in info]); nf = np.array([x['n_feat'] for x in info]); it = np.array([x['iters'] for x in info])
                if S == 0:
                    ref[k] = (c0, nf, it)
                else:
                    r0, rn, ri = ref[k]
                    bad = [b for b in range(B) if not (abs(c0[b] - r0[b]) <= 1e-6 * abs(r0[b])) or nf[b] != rn[b]]
                    msg = f"kf {k}: nan {int(np.isnan(c0).sum())} differing {len(bad)} {bad[:12]}" + (f" e.g. b={bad[0]} c0 {c0[bad[0]]:.5f} vs {r0[bad[0]]:.5f} nf {nf[bad[0]]} vs {rn[bad[0]]} it {it[bad[0]]} vs {ri[bad[0]]}" if bad else "")
                    break
    print(f"rep {rep} S={S}: {msg}", flush=True)
    pipe.close()
