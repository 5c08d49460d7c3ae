"""Diagnostic: run the bench's e2e (host-input) pipeline several times in one process with the per-kernel event timers on and
per-step CUDA events on both streams, to find what stalls in the slow mode."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B, steps, warm = 128, 30, 6
prof_on = len(sys.argv) > 1 and sys.argv[1] == "prof"
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
W = cfg.window_size
prologue = 3 * (W + 1)
n_frames = prologue + warm + steps + 3
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames, 0, "cuda:0")
dt_p, acc_p, gyr_p = (torch.as_tensor(np.ascontiguousarray(x)).pin_memory().numpy() for x in (dt, acc, gyr))
src = frames.cpu().pin_memory().numpy()
imu_host = lambda k: (dt_p[k], acc_p[k], gyr_p[k])
for rep in range(4):
    s_fe, s_be = torch.cuda.Stream(), torch.cuda.Stream()
    pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_be.cuda_stream, gt, True)
    with torch.cuda.stream(s_fe):
        for i in range(prologue + warm):
            pipe.step(src[i], imu_host)
        torch.cuda.synchronize()
        if prof_on:
            pipe.fe.profile(True); pipe.be.profile(True)
        ev_fe, ev_be, host = [], [], []
        e0 = torch.cuda.Event(enable_timing=True); e0.record(s_fe)
        s_be.wait_stream(s_fe)
        t0 = time.perf_counter()
        for i in range(prologue + warm, prologue + warm + steps):
            pipe.step(src[i], imu_host)
            a = torch.cuda.Event(enable_timing=True); a.record(s_fe); ev_fe.append(a)
            b = torch.cuda.Event(enable_timing=True); b.record(s_be); ev_be.append(b)
            host.append(time.perf_counter() - t0)
        s_fe.wait_stream(s_be)
        e1 = torch.cuda.Event(enable_timing=True); e1.record(s_fe)
        torch.cuda.synchronize()
    print(f"rep {rep}: total {e0.elapsed_time(e1):.1f} ms")
    print("  fe stream step-end ms:", " ".join(f"{e0.elapsed_time(a):.1f}" for a in ev_fe))
    print("  be stream step-end ms:", " ".join(f"{e0.elapsed_time(a):.1f}" for a in ev_be))
    print("  host step-end ms     :", " ".join(f"{h*1e3:.1f}" for h in host))
    if prof_on:
        p = {}; p.update(pipe.fe.profile(False)); p.update(pipe.be.profile(False))
        print("  kernels:", {k: (c, round(t / c, 3)) for k, (c, t) in p.items() if c})
    print("  info0:", pipe.be.info(0))
    pipe.close()
