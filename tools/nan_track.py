"""Diagnostic: single CUDA stream, device inputs, sync + inspect after every keyframe; prints per keyframe how many streams have NaN
cost0/cost1, NaN state, NaN prior (stream 0..3)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 45
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames + 3, 0, "cuda:0")
dt_d, acc_d, gyr_d = (torch.as_tensor(x, device="cuda:0").contiguous() for x in (dt, acc, gyr))
imu_dev = lambda k: (dt_d[k].data_ptr(), acc_d[k].data_ptr(), gyr_d[k].data_ptr())
for rep in range(reps):
    s_fe = torch.cuda.Stream()
    pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_fe.cuda_stream, gt, False)
    out = []
    with torch.cuda.stream(s_fe):
        for i in range(n_frames):
            pub = pipe.step(frames[i].data_ptr(), imu_dev)
            if pub and pipe.kf > 9 and (not os.environ.get("SYNC_FROM") or pipe.kf > int(os.environ["SYNC_FROM"])):
                torch.cuda.synchronize()
                info = [pipe.be.info(b) for b in range(B)]
                st = pipe.be.state_all()
                c0 = np.array([x['cost0'] for x in info]); c1 = np.array([x['cost1'] for x in info])
                pn = 0
                for b in range(min(B, 4)):
                    p = pipe.be.prior(b)
                    if p is not None and (np.isnan(p['H']).any() or np.isnan(p['b']).any() or np.isnan(p['c0'])):
                        pn += 1
                out.append(f"kf{pipe.kf - 1}: nanc0 {int(np.isnan(c0).sum())} nanc1 {int(np.isnan(c1).sum())} nanstate {int(np.isnan(st).any(axis=(1, 2)).sum())} nanprior4 {pn} "
                           f"it {info[0]['iters']} retry {info[0]['chol_retry']} flag {info[0]['solver_flag']} c0[0] {c0[0]:.4f} c1[0] {c1[0]:.4f}")
    bad = [o for o in out if "nanc0 0 nanc1 0 nanstate 0 nanprior4 0" not in o]
 
    print("   first: " + out[0])
    print(f"rep {rep}: {len(out)} kfs; first bad: " + " | ".join(bad[:3]) + " || last: " + out[-1])
    pipe.close()
