"""Debug: run the pipeline on the debug library (guard bands around every device array) and report out-of-bounds writes.
Usage (GPU box): VIO_LIB_NAME=libvio_b200_dbg.so python tools/guard_check.py [B] [frames] [one|two]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api"); synth = importlib.import_module("vins-mobile_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
two = len(sys.argv) > 3 and sys.argv[3] == "two"
cfg = abi.default_config(batch=B, max_cnt=150, window_size=10, device=0)
frames, dt, acc, gyr, gt, cam = bench.make_data(synth, B, n_frames + 3, 0, "cuda:0")
dt_d, acc_d, gyr_d = (torch.as_tensor(x, device="cuda:0").contiguous() for x in (dt, acc, gyr))
imu_dev = lambda k: (dt_d[k].data_ptr(), acc_d[k].data_ptr(), gyr_d[k].data_ptr())
s_fe = torch.cuda.Stream(); s_be = torch.cuda.Stream() if two else s_fe
pipe = bench.Pipeline(api, cfg, s_fe.cuda_stream, s_be.cuda_stream, gt, False)
with torch.cuda.stream(s_fe):
    for i in range(n_frames):
        pipe.step(frames[i].data_ptr(), imu_dev)
        if i in (0, 1, 2, 3, 4, 6, 9, 30, 33):
            torch.cuda.synchronize()
            print(f"after frame {i}: corrupted guard bands = {api.lib().vio_debug_check_guards()}", flush=True)
    torch.cuda.synchronize()
print("final: corrupted guard bands =", api.lib().vio_debug_check_guards())
info = [pipe.be.info(b) for b in range(B)]
print("cost0[:3]", [round(i['cost0'], 4) for i in info[:3]], "n_proj[:3]", [i['n_proj'] for i in info[:3]])
