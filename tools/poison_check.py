"""Debug: run the pipeline with the debug library (shared memory poisoned at kernel entry, one kernel per run) and report which
kernel's results change -> that kernel reads shared memory it never wrote.  Build the library first:
  nvcc ... -DVIO_DEBUG_POISON -o vins-mobile_b200/libvio_b200_dbg.so frontend.cu backend.cu
Usage (GPU box): VIO_LIB_NAME=libvio_b200_dbg.so python tools/poison_check.py [B] [frames]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames + 3, 0, "cuda:0")
dt_d, acc_d, gyr_d = (torch.as_tensor(x, device="cuda:0").contiguous() for x in (dt, acc, gyr))
imu_dev = lambda k: (dt_d[k].data_ptr(), acc_d[k].data_ptr(), gyr_d[k].data_ptr())
names = ["imu", "addfeat", "triangulate", "prepare", "post_solve", "finish", "marg", "solve", "lk", "eig_candidates", "post_track", "select"]
ref = None
for mask, nm in [(0, "none")] + [(1 << i, n) for i, n in enumerate(names)]:
    os.environ["VIO_POISON_MASK"] = str(mask)
    s_fe = torch.cuda.Stream()
    pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_fe.cuda_stream, gt, False)
    with torch.cuda.stream(s_fe):
        for i in range(n_frames):
            pipe.step(frames[i].data_ptr(), imu_dev)
        torch.cuda.synchronize()
    info = [pipe.be.info(b) for b in range(B)]
    c0 = np.array([x['cost0'] for x in info]); nf = np.array([x['n_feat'] for x in info]); npj = np.array([x['n_proj'] for x in info])
    if ref is None:
        ref = (c0, nf, npj)
        print(f"mask none: cost0[:4] {c0[:4]} nan {int(np.isnan(c0).sum())}")
    else:
        bad = [b for b in range(B) if not (abs(c0[b] - ref[0][b]) <= 1e-5 * abs(ref[0][b])) or nf[b] != ref[1][b] or npj[b] != ref[2][b]]
        print(f"poison {nm:15s}: nan {int(np.isnan(c0).sum()):3d} differing streams {len(bad):3d} {bad[:8]}", flush=True)
    pipe.close()
