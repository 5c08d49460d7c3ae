"""Probe: H2D bandwidth from a large torch-pinned buffer in 39 MB slices (the e2e bench's access pattern)."""
import sys, time, torch
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 72
B, R, C = 128, 640, 480
t0 = time.time()
host = torch.randint(0, 255, (n_frames, B, R, C), dtype=torch.uint8)
t1 = time.time()
pin = host.pin_memory()
t2 = time.time()
dst = torch.empty((B, R, C), dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
res = []
with torch.cuda.stream(s):
    for rep in range(2):
        for i in range(n_frames):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); dst.copy_(pin[i], non_blocking=True); e1.record(s)
            res.append((e0, e1))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in res]
print(f"alloc {t1-t0:.1f}s pin {t2-t1:.1f}s is_pinned={pin.is_pinned()}  H2D 39MB ms: min {min(ms):.2f} med {sorted(ms)[len(ms)//2]:.2f} max {max(ms):.2f}  first10 {[round(x,2) for x in ms[:10]]}")
