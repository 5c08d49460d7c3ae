"""Small e2e (host-input) or device-input pipeline run that prints a digest of the results -- for determinism checks and
compute-sanitizer runs.  Usage: python tools/e2e_small.py [B] [frames] [host|dev] [reps]"""
import hashlib, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 45
host = (sys.argv[3] if len(sys.argv) > 3 else "host") == "host"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames + 3, 0, "cuda:0")
dt_p, acc_p, gyr_p = (torch.as_tensor(np.ascontiguousarray(x)).pin_memory().numpy() for x in (dt, acc, gyr))
dt_d, acc_d, gyr_d = (torch.as_tensor(x, device="cuda:0").contiguous() for x in (dt, acc, gyr))
src = frames.cpu().pin_memory().numpy()
imu_host = lambda k: (dt_p[k], acc_p[k], gyr_p[k])
imu_dev = lambda k: (dt_d[k].data_ptr(), acc_d[k].data_ptr(), gyr_d[k].data_ptr())
for rep in range(reps):
    s_fe = torch.cuda.Stream()
    s_be = s_fe if os.environ.get("ONE_STREAM") else torch.cuda.Stream()
    pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_be.cuda_stream, gt, host)
    fe_dig = hashlib.sha1()
    with torch.cuda.stream(s_fe):
        for i in range(n_frames):
            pipe.step(src[i] if host else frames[i].data_ptr(), imu_host if host else imu_dev)
            if os.environ.get("FE_DIGEST"):
                for b in range(B):
                    g = pipe.fe.stream(b); fe_dig.update(g["ids"].tobytes()); fe_dig.update(g["pts"].tobytes())
        torch.cuda.synchronize()
    st = pipe.be.state_all()
    info = [pipe.be.info(b) for b in range(B)]
    print(f"rep {rep} host={host}: state sha {hashlib.sha1(st.tobytes()).hexdigest()[:12]} nan={int(np.isnan(st).sum())} fe {fe_dig.hexdigest()[:12]} "
          f"n_feat sha {hashlib.sha1(str([i['n_feat'] for i in info]).encode()).hexdigest()[:8]} {[i['n_feat'] for i in info][:3]} n_proj {[i['n_proj'] for i in info][:3]} cost0 {[round(i['cost0'], 6) for i in info][:3]} n_nan_cost {sum(1 for i in info if i['cost0'] != i['cost0'])}")
    costs = np.array([i['cost0'] for i in info]); nf = np.array([i['n_feat'] for i in info])
    if rep == 0:
        costs0, nf0 = costs, nf
    else:
        bad = [b for b in range(B) if nf[b] != nf0[b] or not (abs(costs[b] - costs0[b]) <= 1e-3 * abs(costs0[b]))]
        print("   streams differing from rep 0:", bad[:20], "count", len(bad))
    pipe.close()
