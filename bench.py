#!/usr/bin/env python
"""bench.py -- VIO frames/sec (640x480, 150 feats, 10-KF window) on B200 vs the reference CPU path.

A "step" is ONE camera frame for every stream of the batch: FeatureTracker::readImage on all streams and, on every FREQ-th
frame, 20 IMU samples through VINS::processIMU + one VINS::processImage (triangulate, <=10 dogleg iterations, marginalisation,
slideWindow).  Workload at N GPUs: 128 independent streams per GPU (BASELINE.json configs[2]; configs[3] is the same shape at
N=8), weak scaling, streams sharded by rank with no data-path collective; the only collective is an NCCL all-gather of the
packed window states after each solve when N > 1.

  python bench.py --gpus 1 --steps 30 --warmup 6          # ours
  python bench.py --impl reference --steps 30 --warmup 6  # reference CPU arm (cv2 KLT/RANSAC/GFTT + reference factors + Ceres)

Timed region: CUDA events on the stream the kernels run on, barrier + synchronize on both sides, max over ranks.  Inputs of
successive steps are different frames (39 MB per step at B=128) and the working set (pyramids, windows, scratch ~0.6 GB) is
larger than the 126 MB L2, so no explicit L2 flush is needed (config.l2 says so).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FREQ = 3
IMU_PER_KF = 20
# BASELINE.json configs -> concrete shapes.  c2 (= c3 at 8 GPUs) is the workload `value` is quoted on; c1 and c4 are measured as
# extra lines ("configs" in the JSON) or on their own with --config.
CONFIGS = {
    "c1": dict(rows=640, cols=480, max_cnt=150, window=10, batch=1,
               name="BASELINE.json configs[1]: single stream 640x480@30fps + 200 Hz IMU, 150 feats, 10-KF window"),
    "c2": dict(rows=640, cols=480, max_cnt=150, window=10, batch=128,
               name="batch 128 independent streams per GPU, 640x480@30fps + 200 Hz IMU, 150 feats, 10-KF window, FREQ=3 "
                    "(BASELINE.json configs[2]; configs[3] at 8 GPUs)"),
    "c4": dict(rows=720, cols=1280, max_cnt=300, window=20, batch=32,
               name="BASELINE.json configs[4]: 1280x720, 300 feats, 20-KF window, 32 streams per GPU (batch 256 across 8 GPUs)"),
}


def algo_bytes(rows, cols, n):
    """algorithmic bytes per camera frame -- pyramid 1.328 HW (read L0, write L1..L3), KLT 4232 N,
    detection 2 HW on a detect frame."""
    hw = rows * cols
    lv = [hw]
    r, c = rows, cols
    for _ in range(3):
        r, c = (r + 1) // 2, (c + 1) // 2
        lv.append(r * c)
    pyr = hw + sum(lv[1:])
    return dict(pyr=pyr, detect=2 * hw, klt=4232 * n, frame=pyr + 4232 * n + 2 * hw / FREQ)


def _rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class Clocks(threading.Thread):
    """SM-clock / throttle-reason sampler running during the timed region (B200_PROFILING.md clocks line).  NVML in-process (a sample
    every 5 ms; the timed region of the default run lasts ~70 ms, shorter than one nvidia-smi start-up), nvidia-smi as fallback."""
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.stop_flag, self.rows = dev, False, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _reasons(self, mask):
        n = self.nvml
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        return [k for k, v in names.items() if mask & getattr(n, v, 0)]

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            while not self.stop_flag:
                try:
                    sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                    try:
                        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((sm, self._reasons(mask)))
                except Exception:
                    pass
                time.sleep(0.005)
            return
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")] if out else []
                if len(r) >= 7 and r[0].replace(".", "").isdigit():
                    self.max_sm = float(r[1])
                    self.rows.append((float(r[0]), [nm for i, nm in enumerate(names) if r[3 + i].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[1]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": getattr(self, "max_sm", None), "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------- data
def make_data(synth, B, n_frames, stream0, device, cam=None):
    """(n_frames, B, rows, cols) u8 frames on `device`, IMU (n_kf-1, 20, B, .) and ground-truth init windows."""
    import torch
    cam = cam or synth.Camera()
    surf = synth.Surface(1000, device=device)                        # one texture, per-stream trajectories
    frames = torch.empty((n_frames, B, cam.rows, cam.cols), dtype=torch.uint8, device=device)
    n_kf = (n_frames + FREQ - 1) // FREQ
    dt = np.full((n_kf - 1, IMU_PER_KF, B), 1.0 / 200.0)
    acc = np.zeros((n_kf - 1, IMU_PER_KF, B, 3))
    gyr = np.zeros((n_kf - 1, IMU_PER_KF, B, 3))
    gt = []
    for b in range(B):
        s = synth.make_stream(stream0 + b, n_frames, cam=cam, device=device, surface=surf)
        frames[:, b] = s.images
        m = (n_kf - 1) * IMU_PER_KF
        acc[:, :, b] = s.acc[:m].reshape(n_kf - 1, IMU_PER_KF, 3)
        gyr[:, :, b] = s.gyr[:m].reshape(n_kf - 1, IMU_PER_KF, 3)
        gt.append((s.P[::FREQ], synth.rot_to_quat_xyzw(s.R[::FREQ]), s.V[::FREQ]))
    return frames, dt, acc, gyr, gt, cam


# ---------------------------------------------------------------------------------------------------- our arm
class Pipeline:
    """One FeatureTracker + VINS pair per stream, batched, both on one CUDA stream (front end hands image_msg to the back end
    in device memory)."""

    def __init__(self, api, cfg, fe_stream, be_stream, gt, host_inputs):
        self.api, self.cfg, self.W, self.B = api, cfg, cfg.window_size, cfg.batch
        self.fe = api.FrontEnd(cfg)
        self.be = api.BackEnd(cfg)
        self.fe.use_stream(fe_stream)          # tracker of frames k+1.. overlaps the solve of keyframe k (event-ordered hand-over)
        self.be.use_stream(be_stream)
        self.gt, self.host = gt, host_inputs
        self.kf = 0
        self.frame = 0
        self.state_host = None
        self.be_stream = be_stream
        if host_inputs:
            import torch
            self.state_pin = [torch.empty((self.B, self.W + 1, 16), dtype=torch.float64).pin_memory() for _ in range(2)]
            self.state_evt = [None, None]

    def step(self, img, imu):
        """img: device pointer (device-resident run) or pinned host ndarray (e2e run).  imu(kf) -> (dt, acc, gyr) device pointers
        or host arrays."""
        pub = self.fe.read_images(img) if self.host else self.fe.read_images_dev(img)
        if pub:
            k = self.kf
            if k > 0:
                d, a, g = imu(k - 1)
                if self.host:
                    self.be.process_imu(d, a, g)
                else:
                    self.be.process_imu_dev(IMU_PER_KF, d, a, g)
            if k == self.W:
                P = np.stack([g[0][:self.W + 1] for g in self.gt]); Q = np.stack([g[1][:self.W + 1] for g in self.gt])
                V = np.stack([g[2][:self.W + 1] for g in self.gt])
                self.be.set_init_window(P, Q, V, np.zeros((self.B, 3)), np.zeros((self.B, 3)))
            self.be.process_image_from_frontend(self.fe, np.full(self.B, self.frame / 30.0))
            if self.host:
                # device -> host read of the step's result: stream-ordered copy into pinned memory; the host consumes the result of
                # the PREVIOUS keyframe (its event has long fired), so the tracker of the next frames is enqueued while this
                # keyframe's solve runs
                import torch
                k2 = self.kf & 1
                if self.state_evt[k2 ^ 1] is not None:
                    self.state_evt[k2 ^ 1].synchronize()
                    self.state_host = float(self.state_pin[k2 ^ 1][:, -1, :3].sum())
                self.be.copy_state(self.state_pin[k2].data_ptr(), 2)
                self.state_evt[k2] = torch.cuda.Event()
                self.state_evt[k2].record(torch.cuda.ExternalStream(self.be_stream))
            self.kf += 1
        self.frame += 1
        return pub

    def close(self):
        self.fe.close()
        self.be.close()


def _solve_flops(i, W):
    """SURVEY.md section 8(d) flop model of one window solve with the stream's own P, L, iterations."""
    NPd = 15 * (W + 1)
    P, L, it, npr = i["n_proj"], max(i["n_feat"], 1), i["iters"], i["prior_n"]
    lin = 1500.0 * P + 45000.0 * W + 2.0 * npr * npr
    schur = 2.0 * L * (6.0 * (P / L + 1.0)) ** 2
    chol = NPd ** 3 / 3.0 + 2.0 * NPd * NPd
    cost_only = 0.4 * (1500.0 * P + 45000.0 * W) + 2.0 * npr * npr
    return lin + it * (lin + schur + chol + cost_only)          # first linearisation + one (re)linearisation per iteration


class Bench:
    """Everything one rank needs to time a configuration: data, handles, streams."""

    def __init__(self, args, conf, rank, local, world):
        import torch
        self.torch = torch
        self.args, self.conf, self.rank, self.local, self.world = args, conf, rank, local, world
        self.dev = f"cuda:{local}"
        self.abi = importlib.import_module("vins-mobile_b200.abi")
        self.api = importlib.import_module("vins-mobile_b200.api")
        self.synth = importlib.import_module("vins-mobile_b200.synth")
        self.B, self.W = conf["batch"], conf["window"]
        self.prologue = FREQ * (self.W + 1)                  # fills the window and runs the first (initialising) solve
        self.n_frames = self.prologue + args.warmup + args.steps + FREQ
        self.cam = self.synth.Camera() if (conf["rows"], conf["cols"]) == (640, 480) else self.synth.Camera().scaled(conf["rows"], conf["cols"])
        t0 = time.time()
        self.frames, self.dt, self.acc, self.gyr, self.gt, _ = make_data(self.synth, self.B, self.n_frames, rank * self.B, self.dev, self.cam)
        self.t_data = time.time() - t0
        self.dt_d, self.acc_d, self.gyr_d = (torch.as_tensor(x, device=self.dev).contiguous() for x in (self.dt, self.acc, self.gyr))
        self.dt_p, self.acc_p, self.gyr_p = (torch.as_tensor(np.ascontiguousarray(x)).pin_memory().numpy() for x in (self.dt, self.acc, self.gyr))
        self.stream = torch.cuda.Stream(device=self.dev)          # front end
        # the back end is the critical path (one CTA per stream, a full SM each): its kernels get the free SMs first
        self.stream_b = torch.cuda.Stream(device=self.dev, priority=-1) if not args.no_overlap else self.stream
        self.stream_c = torch.cuda.Stream(device=self.dev)        # NCCL gather of the window states, off the critical path
        if world > 1:
            self.gather_buf = torch.empty((world, self.B, self.W + 1, 16), dtype=torch.float64, device=self.dev)
            self.send_buf = [torch.empty((self.B, self.W + 1, 16), dtype=torch.float64, device=self.dev) for _ in range(2)]

    def cfg(self, **over):
        c = self.abi.default_config(batch=self.B, max_cnt=self.conf["max_cnt"], window_size=self.W, rows=self.conf["rows"], cols=self.conf["cols"],
                                    device=self.local)
        for k, v in over.items():
            setattr(c, k, v)
        return c

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed_run(self, host_inputs, cfg=None, clahe=False, profile=True):
        torch, args, B, W, world = self.torch, self.args, self.B, self.W, self.world
        cfg = cfg or self.cfg()
        stream, stream_b, stream_c = self.stream, self.stream_b, self.stream_c
        pipe = Pipeline(self.api, cfg, stream.cuda_stream, stream_b.cuda_stream, self.gt, host_inputs)
        if clahe:
            pipe.fe.set_clahe(True, 3.0, 8, 8)
        src = self.frames.cpu().pin_memory().numpy() if host_inputs else None
        imu_host = lambda k: (self.dt_p[k], self.acc_p[k], self.gyr_p[k])
        imu_dev = lambda k: (self.dt_d[k].data_ptr(), self.acc_d[k].data_ptr(), self.gyr_d[k].data_ptr())
        ev_copy, ev_gather, nkf = [None, None], [None, None], [0]
        prologue = self.prologue
        with torch.cuda.stream(stream):
            def one(i):
                pub = pipe.step(src[i] if host_inputs else self.frames[i].data_ptr(), imu_host if host_inputs else imu_dev)
                if pub and world > 1:
                    # the only collective: all-gather of the packed window states.  The copy into the (double-buffered) send buffer is
                    # stream-ordered behind the solve; the gather itself runs on a third stream behind an event, so the next keyframe's
                    # back-end kernels never queue behind NCCL
                    import torch.distributed as dist
                    k2 = nkf[0] & 1
                    nkf[0] += 1
                    if ev_gather[k2] is not None:
                        stream_b.wait_event(ev_gather[k2])
                    with torch.cuda.stream(stream_b):
                        pipe.be.copy_state(self.send_buf[k2].data_ptr(), True)
                    ev_copy[k2] = torch.cuda.Event(); ev_copy[k2].record(stream_b)
                    stream_c.wait_event(ev_copy[k2])
                    with torch.cuda.stream(stream_c):
                        dist.all_gather_into_tensor(self.gather_buf.view(world * B, W + 1, 16), self.send_buf[k2])
                    ev_gather[k2] = torch.cuda.Event(); ev_gather[k2].record(stream_c)
            for i in range(prologue + args.warmup):
                one(i)
            self.barrier()
            l0 = pipe.fe.launch_count() + pipe.be.launch_count()
            clocks = Clocks(self.local)
            if self.rank == 0 and not os.environ.get("VIO_BENCH_NO_CLOCKS"):
                clocks.start()
            trace = [] if os.environ.get("VIO_BENCH_TRACE") else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            stream_b.wait_stream(stream)
            for i in range(prologue + args.warmup, prologue + args.warmup + args.steps):
                one(i)
                if trace is not None:
                    trace.append(time.perf_counter())
            stream.wait_stream(stream_b)
            stream.wait_stream(stream_c)
            e1.record(stream)
            self.barrier()
            clocks.stop_flag = True
            ms = e0.elapsed_time(e1)
            if trace:
                print("host ms per step (host_inputs=%s): " % host_inputs + " ".join(f"{(b - a) * 1e3:.2f}" for a, b in zip(trace, trace[1:])), file=sys.stderr)
            launches = pipe.fe.launch_count() + pipe.be.launch_count() - l0
            prof, info, phase_us = {}, [], None
            if profile:
                # per-kernel CUDA-event pass (separate, untimed): one more keyframe period with both handles on ONE stream, so that every
                # kernel is timed alone (under the two-stream overlap a front-end kernel's event time includes waiting for SMs the solve holds)
                torch.cuda.synchronize()
                pipe.be.use_stream(stream.cuda_stream)
                pipe.fe.profile(True); pipe.be.profile(True); pipe.be.phase_cycles(True)
                base = prologue + args.warmup + args.steps
                for i in range(base, base + FREQ):
                    one(i)
                prof.update(pipe.fe.profile(False)); prof.update(pipe.be.profile(False))
                info = [pipe.be.info(b) for b in range(B)]
                ph = pipe.be.phase_cycles(True).astype(float)
                names = ["solve.linearize", "solve.scale_cauchy", "solve.schur", "solve.cholesky", "solve.dogleg_model", "solve.cost_eval", "solve.accept", "",
                         "marg.setup", "marg.accumulate", "marg.slow_amm", "marg.amm_inv+schur", "marg.eig", "marg.recompose", "", "",
                         "lin.prior", "lin.imu", "lin.projection", "lin.cost_sum", "chol.trailing_update", "chol.diag_kloop", "chol.diag_factor",
                         "chol.backward", "chol.wait_A(eig.tred2)", "chol.phase_B(eig.accumulate)", "eig.tql2", "proj.pair_tables", "proj.jacobians", "proj.block_sums", "cost.prior", "cost.imu"]
                phase_us = {nm: round(float(ph[:, i].max()) / 1.9e3, 1) for i, nm in enumerate(names) if nm}
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        pipe.close()
        return dict(ms=ms, launches=launches, prof=prof, info=info, phases=phase_us, clocks=clocks.summary() if self.rank == 0 else None)

    def n_frames_parity_cap(self):
        return 100000

    def parity_states_on(self, frames_host, dt, acc, gyr, gt, nb, last_frame):
        """Untimed pass for `parity_in_bench`: a batch-nb pipeline of this library through the host frames the CPU arm consumed
        (frames_host [n][nb][rows][cols]); returns the packed window states and the tracker state after the last frame."""
        torch = self.torch
        cfg = self.abi.default_config(batch=nb, max_cnt=150, window_size=10, device=self.local)
        pipe = Pipeline(self.api, cfg, self.stream.cuda_stream, self.stream.cuda_stream, gt[:nb], True)
        imu = lambda k: (np.ascontiguousarray(dt[k][:, :nb]), np.ascontiguousarray(acc[k][:, :nb]), np.ascontiguousarray(gyr[k][:, :nb]))
        with torch.cuda.stream(self.stream):
            for i in range(last_frame + 1):
                pipe.step(np.ascontiguousarray(frames_host[i]), imu)
            torch.cuda.synchronize()
            st = pipe.be.state_all()
            tr = [pipe.fe.stream(b) for b in range(nb)]
        pipe.close()
        return st, tr


def _kernel_table(prof, B, ab):
    kern = {k: {"launches": c, "ms_per_launch": t / c} for k, (c, t) in prof.items() if c}
    cand = {"pyr_down_kernel": ab["pyr"] / 3.0, "eig_candidates_kernel": ab["detect"], "lk_kernel": ab["klt"]}
    for k in kern:
        if k in cand:
            kern[k]["achieved_GBps"] = cand[k] * B / (kern[k]["ms_per_launch"] * 1e-3) / 1e9
    return kern, cand


def pnp_bench(bn, args, n_frames=40):
    """vio_pnp_* through the host API: per camera frame the IMU samples since the last frame + the frame's solved landmarks (60 per stream) +
    every third frame a (lagging) estimator result, as feature_tracker.cpp:107-160 feeds vinsPnP.  Frames per second over B windows."""
    torch = bn.torch
    B = bn.B
    cfg = bn.abi.default_config(batch=B, max_cnt=150, device=bn.local)
    seqs = [bn.synth.make_pnp_sequence(b, n_frames) for b in range(min(B, 8))]        # 8 distinct sequences, tiled over the batch
    pnp = bn.api.PnP(cfg)
    n_lm = len(seqs[0]["ids"])
    cnt = np.full(B, n_lm, np.int32)
    ids = np.zeros((B, 150), np.int32); tn = np.zeros((B, 150), np.int32); pos = np.zeros((B, 150, 3)); obs = np.zeros((B, 150, 2))
    for b in range(B):
        q = seqs[b % len(seqs)]
        ids[b, :n_lm] = q["ids"]; tn[b, :n_lm] = q["track_num"]; pos[b, :n_lm] = q["X"]
    last_t = 0.0
    t0 = None
    for k in range(n_frames):
        if k == 8:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        q0 = seqs[0]
        if k >= q0["lag"] and k % 3 == 0:
            j = k - q0["lag"]
            pnp.set_init(np.full(B, q0["t"][j]), np.stack([seqs[b % len(seqs)]["P"][j] for b in range(B)]), np.stack([seqs[b % len(seqs)]["R"][j] for b in range(B)]),
                         np.stack([seqs[b % len(seqs)]["V"][j] for b in range(B)]), np.zeros((B, 3)), np.zeros((B, 3)))
        sel = (q0["imu_t"] > last_t + 1e-9) & (q0["imu_t"] <= q0["t"][k] + 1e-9)
        if sel.any():
            tt = np.concatenate([[last_t], q0["imu_t"][sel]])
            pnp.process_imu(np.repeat(np.diff(tt)[:, None], B, 1), np.stack([seqs[b % len(seqs)]["acc"][sel] for b in range(B)], 1),
                            np.stack([seqs[b % len(seqs)]["gyr"][sel] for b in range(B)], 1))
        for b in range(B):
            obs[b, :n_lm] = seqs[b % len(seqs)]["obs"][k]
        pnp.process_image(cnt, ids, obs, pos, tn, np.full(B, q0["t"][k]), True)
        last_t = q0["t"][k]
    st = pnp.state(0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pnp.close()
    return {"value": (n_frames - 8) * B / dt, "unit": "frames/s", "windows": B, "landmarks_per_frame": n_lm, "frames_timed": n_frames - 8,
            "ms_per_frame_all_windows": 1e3 * dt / (n_frames - 8), "iters_last": st["iters"], "err": st["err"],
            "what": "vinsPnP::processIMU + processImage (5 dogleg iterations on the 7-frame window) through the host C-ABI, host loop included"}


def align_bench(bn, args, n=11, reps=5):
    """vio_visual_imu_align (VisualIMUAlignment, initial_aligment.cpp:222-229) for B streams x n frames through the host C-ABI (host copies inside
    the call), next to the reference's own code (oracle/_ref) on the first 8 of the same streams, one thread."""
    B = bn.B
    cases = [bn.synth.make_align_case(b, n) for b in range(min(B, 16))]
    pick = [cases[b % len(cases)] for b in range(B)]
    nf = np.full(B, n, np.int32)
    R = np.stack([c["R"] for c in pick]); T = np.stack([c["T"] for c in pick]); cnt = np.stack([c["counts"] for c in pick])
    i0 = np.stack([c["imu0"] for c in pick]); imu = np.stack([c["imu"] for c in pick]); bg0 = np.zeros((B, 3))
    cfg = bn.abi.default_config(batch=1, device=bn.local)
    bn.api.visual_imu_align(cfg, nf, R, T, cnt, i0, imu, bg0)
    t0 = time.perf_counter()
    for _ in range(reps):
        bgs, g, x, ok = bn.api.visual_imu_align(cfg, nf, R, T, cnt, i0, imu, bg0)
    gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
    out = {"streams": B, "frames_per_stream": n, "imu_samples_per_interval": int(cnt[0, 1]), "gpu_ms_all_streams": gpu_ms, "ok": int(ok.sum()),
           "what": "batched VisualIMUAlignment through the host C-ABI (H2D, one CTA per stream, D2H)"}
    try:                                   # cpu_baseline leg of this variant (the checker, timed beside): skipped with --no-cpu
        if args.no_cpu:
            return out
        import oracle.backend_oracle as bo
        if bo.available() and hasattr(bo.lib(), "vref_visual_imu_align"):
            t0 = time.perf_counter()
            worst = 0.0
            for b, c in enumerate(cases[:8]):
                rb, rg, rx, rok = bo.visual_imu_align(n, c["R"], c["T"], c["counts"], c["imu0"], c["imu"], np.zeros(3), c["tic"])
                worst = max(worst, float(np.abs(x[b, :3 * n + 3] - rx).max() / np.abs(rx).max()))
            out["cpu_baseline"] = {"ms_per_stream": 1e3 * (time.perf_counter() - t0) / 8, "cores": 1, "kind": "reference",
                                   "sample": "the reference's own initial_aligment.cpp (oracle/_ref) on the first 8 of the same streams"}
            out["max_rel_err_vs_reference"] = worst
    except Exception as e:
        out["cpu_reference_error"] = repr(e)
    return out


def run_ours(args):
    import torch
    rank, local, world = _rank_world()
    # rank 0's stdout must carry exactly ONE JSON line: libraries that print there (NCCL's version banner at communicator creation)
    # are sent to stderr for the duration of the run; the descriptor is restored right before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version banner on stdout; rank 0's stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    conf = dict(CONFIGS[args.config])
    if args.batch:
        conf["batch"] = args.batch
    bn = Bench(args, conf, rank, local, world)
    B, W = bn.B, bn.W
    ab = algo_bytes(conf["rows"], conf["cols"], conf["max_cnt"])
    main = bn.timed_run(False)
    e2e = bn.timed_run(True, profile=False)
    extras = {}
    if world == 1 and not args.no_extras:
        # (a) the reference's own marginalisation route on the GPU (eigendecompositions of Amm and A_r, marginalization_factor.cpp:270-294)
        #     -- the like-for-like number next to the default information-form prior
        r = bn.timed_run(False, cfg=bn.cfg(marg_mode=1, marg_amm_eig=1), profile=True)
        extras["marg_reference_route"] = {"value": args.steps * B / (r["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": r["ms"] / args.steps,
                                          "marg_kernel_ms": r["prof"].get("marg_kernel", (1, float("nan")))[1] / max(r["prof"].get("marg_kernel", (1, 0))[0], 1),
                                          "what": "vio_config marg_mode=1, marg_amm_eig=1: Amm^+ and the new prior through symmetric eigendecompositions as the reference does"}
        # (b) CLAHE pre-processing switched on (ViewController.mm:438-441 runs it before every readImage)
        r = bn.timed_run(False, clahe=True, profile=True)
        extras["clahe_on"] = {"value": args.steps * B / (r["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": r["ms"] / args.steps,
                              "kernels_ms": {k: v[1] / v[0] for k, v in r["prof"].items() if k.startswith("clahe")}}
        # (c) the motion-only PnP tracker behind FeatureTracker::solveVinsPnP (vins_pnp.cpp), B windows in lock-step, host buffers in
        try:
            extras["pnp_tracker"] = pnp_bench(bn, args)
        except Exception as e:
            extras["pnp_tracker"] = {"error": repr(e)}
        # (d) the visual-inertial alignment of the initialisation (initial_aligment.cpp), B streams x 11 frames
        try:
            extras["visual_imu_align"] = align_bench(bn, args)
        except Exception as e:
            extras["visual_imu_align"] = {"error": repr(e)}
    single = None
    configs_extra = {}
    if world == 1 and rank == 0 and not args.no_extras and args.config == "c2":
        # BASELINE.json configs[1]: ONE stream (latency-bound: one CTA per back-end kernel), same frames as stream 0
        b1 = Bench.__new__(Bench)
        b1.__dict__.update(bn.__dict__)
        b1.conf = dict(CONFIGS["c1"]); b1.B = 1
        b1.frames = bn.frames[:, :1].contiguous()
        b1.dt_d, b1.acc_d, b1.gyr_d = (torch.as_tensor(np.ascontiguousarray(x[:, :, :1]), device=bn.dev) for x in (bn.dt, bn.acc, bn.gyr))
        b1.dt_p, b1.acc_p, b1.gyr_p = (np.ascontiguousarray(x[:, :, :1]) for x in (bn.dt, bn.acc, bn.gyr))
        b1.gt = bn.gt[:1]
        r1 = b1.timed_run(False, profile=False)
        single = {"workload": CONFIGS["c1"]["name"] + ", 1 GPU", "frames_per_s": args.steps / (r1["ms"] * 1e-3), "ms_per_frame": r1["ms"] / args.steps,
                  "real_time_factor_at_30fps": args.steps / (r1["ms"] * 1e-3) / 30.0, "gpu_launches": int(r1["launches"])}
        configs_extra["c1"] = single
        del b1
        # BASELINE.json configs[4] at one GPU's share (32 of the 256 streams): short run, device-resident + end to end
        try:
            a4 = argparse.Namespace(**vars(args)); a4.steps = min(args.steps, 12); a4.warmup = 3
            del bn.frames
            torch.cuda.empty_cache()
            b4 = Bench(a4, dict(CONFIGS["c4"]), rank, local, world)
            r4 = b4.timed_run(False)
            r4e = b4.timed_run(True, profile=False)
            k4, _ = _kernel_table(r4["prof"], b4.B, algo_bytes(720, 1280, 300))
            fl4 = sum(_solve_flops(i, b4.W) for i in r4["info"])
            configs_extra["c4"] = {"workload": CONFIGS["c4"]["name"], "value": a4.steps * b4.B / (r4["ms"] * 1e-3), "unit": "frames/s",
                                   "ms_per_step": r4["ms"] / a4.steps, "steps": a4.steps, "e2e": a4.steps * b4.B / (r4e["ms"] * 1e-3),
                                   "kernels": k4, "solve_TFLOPs": fl4 / (k4["solve_kernel"]["ms_per_launch"] * 1e-3) / 1e12 if "solve_kernel" in k4 else None,
                                   "solve_info_batch": {k: [min(i[k] for i in r4["info"]), max(i[k] for i in r4["info"])] for k in ("iters", "n_feat", "n_proj", "prior_n", "err")}}
            del b4
        except Exception as e:          # the extra line must never cost the headline
            configs_extra["c4"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms, launches, prof, info, clk = main["ms"], main["launches"], main["prof"], main["info"], main["clocks"]
    total_frames = args.steps * B * world
    value = total_frames / (ms * 1e-3)
    e2e_v = total_frames / (e2e["ms"] * 1e-3)
    n_kf_steps = len([i for i in range(args.steps) if (bn.prologue + args.warmup + i) % FREQ == 0])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    kern, cand = _kernel_table(prof, B, ab)
    traffic, traffic_src = {}, None
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))          # newest committed ncu capture (names sort by round)
    if tps and args.config == "c2":
        tj = json.load(open(tps[-1]))
        traffic, traffic_src = tj["bytes_per_launch"], {"file": os.path.relpath(tps[-1], ROOT), "captured_at_commit": tj.get("commit"), "note": tj.get("note")}
    # (1) the dominant kernel of the step: solve_kernel -- FP64 ALU / tensor pipe (DMMA), no HBM roofline (Jacobians are never
    #     materialised).  Algorithmic flops per launch = SURVEY.md section 8(d) flop model with every stream's own P, L, iterations.
    fp64 = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json"))) if os.path.exists(os.path.join(ROOT, "profiles", "fp64_peak.json")) else \
        {"fma_per_clk_per_sm": 64.0, "sms": 148}
    sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    fp64_peak = 2.0 * fp64["fma_per_clk_per_sm"] * fp64["sms"] * sm_mhz * 1e6 / 1e12
    roof = None
    if "solve_kernel" in kern and info:
        fl = sum(_solve_flops(i, W) for i in info)
        ach = fl / (kern["solve_kernel"]["ms_per_launch"] * 1e-3) / 1e12
        roof = {"kernel": "solve_kernel", "bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "traffic": traffic.get("solve_kernel"), "traffic_source": traffic_src,
                "peak_source": "FP64 (the path computes in f64; MEASURED_PEAKS.json has no FP64 figure): profiles/fp64_peak.json = 64 FMA/clk/SM "
                               "measured with tools/ubench/dmma_rate.cu (DFMA and DMMA alike) x 148 SMs x SM clock under load",
                "algorithmic_flops_per_launch": fl, "flop_model": "SURVEY.md section 8(d): per iteration 1500 P + 45000 Wn + 2 n_prior^2 (linearise) + "
                "2 L (6 (P/L+1))^2 (Schur) + n_r^3/3 + 2 n_r^2 (Cholesky) + cost-only evaluation, summed over the batch with each stream's P, L, iterations"}
    # (2) the dominant HBM-class kernel of the front end (per launch = one batch of B images)
    dom = max((k for k in cand if k in kern), key=lambda k: kern[k]["ms_per_launch"] * kern[k]["launches"], default=None)     # time per keyframe period
    roof_fe = None
    if dom:
        bytes_per_launch = cand[dom] * B
        ach = bytes_per_launch / (kern[dom]["ms_per_launch"] * 1e-3) / 1e9
        roof_fe = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic.get(dom),
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                   "algorithmic_bytes_per_launch": bytes_per_launch,
                   "note": "kernel timed alone (single-stream profile pass after the timed region)"}
    cpu = parity = None
    if world == 1 and not args.no_cpu and args.config == "c2":
        cpu, parity = cpu_baselines(args, bn)
    line = {
        "metric": "VIO frames/sec (640x480, 150 feats, 10-KF window)", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
        "config": {"workload": conf["name"], "batch_per_gpu": B, "streams": B * world, "freq": FREQ,
                   "keyframe_steps_timed": n_kf_steps, "l2": "inputs change every step and the working set exceeds L2; no explicit flush",
                   "prologue_frames_untimed": bn.prologue, "data_gen_s": round(bn.t_data, 1),
                   "streams_overlap": "front end and back end on two CUDA streams, event-ordered hand-over; NCCL gather on a third" if not args.no_overlap else "single stream"},
        "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": int(B * bn.cam.rows * bn.cam.cols + B * IMU_PER_KF * 7 * 8 / FREQ),
                "d2h_bytes_per_step": int(B * (W + 1) * 16 * 8 / FREQ), "ms_per_step": e2e["ms"] / args.steps},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_frontend": roof_fe, "kernels": kern, "cpu_baseline": cpu,
        "parity_in_bench": parity, "variants": extras or None, "configs": configs_extra or None, "single_stream": single,
        "backend_phase_us_max_over_streams_at_1p9GHz": main["phases"] if main["phases"] and any(main["phases"].values()) else None,      # debug library only (VIO_LIB_NAME)
        "solve_info_stream0": info[0] if info else None,
        "solve_info_batch": {k: [min(i[k] for i in info), max(i[k] for i in info)] for k in ("iters", "n_feat", "n_proj", "prior_n", "marg_fast", "marg_sweeps", "marg_m", "chol_retry", "err")} if info else None,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- reference arm
class Quiet:
    """fd-level silencer for the reference's printf chatter (marginalization_factor.cpp prints on every call)."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def _cpu_worker(payload):
    """One stream through the CPU reference path: cv2 (OpenCV binary) KLT / RANSAC-F / goodFeaturesToTrack driven by the restated
    readImage, then the reference's factors + vendored Ceres 1.12 driven by the restated estimator loop.  Returns the wall time of the
    timed frames, its split per stage, and the final tracker / window state (for parity_in_bench)."""
    frames, dt, acc, gyr, gt, prologue, n_time, cv_threads, ref_lib = payload
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    if ref_lib:
        os.environ["VINS_REF_LIB"] = ref_lib
    import cv2
    cv2.setNumThreads(cv_threads)
    import frontend_oracle as fo
    import backend_oracle as bo
    abi = importlib.import_module("vins-mobile_b200.abi")
    cfg = abi.default_config(batch=1, max_cnt=150)
    W = cfg.window_size
    tr = fo.FeatureTrackerOracle(max_cnt=150, backend="cv2")
    est = bo.RefEstimator(cfg)
    kf = 0
    t_start = None
    t_fe = 0.0
    # the reference prints from destructors too ("release marginlizationinfo"): this worker's stdout stays on /dev/null for good,
    # results travel back through the pool's pipe
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    with Quiet():
        for i in range(prologue + n_time):
            if i == prologue:
                t_start = time.perf_counter()
                est.stage_seconds(reset=True)
                t_fe = 0.0
            t0 = time.perf_counter()
            _, _, pub = tr.read_image(frames[i])
            t_fe += time.perf_counter() - t0
            if pub:
                if kf > 0:
                    for j in range(IMU_PER_KF):
                        est.process_imu(dt[kf - 1][j], acc[kf - 1][j], gyr[kf - 1][j])
                if kf == W:
                    est.set_init_window(gt[0][:W + 1], gt[1][:W + 1], gt[2][:W + 1], np.zeros(3), np.zeros(3))
                ids = np.array(sorted(tr.image_msg.keys()), np.int32)
                xyz = np.array([tr.image_msg[k] for k in ids])
                est.process_image(ids, xyz, i / 30.0)
                kf += 1
    total = time.perf_counter() - t_start
    st = est.stage_seconds()
    s = est.state()
    packed = np.concatenate([s["P"], s["Q"], s["V"], s["Ba"], s["Bg"]], 1)
    return dict(total=total, front_end=t_fe, ceres_solve=st["ceres_solve"], marginalise=st["marginalise"], process_imu=st["process_imu"],
                process_image=st["process_image"], state=packed, ids=tr.ids.copy(), pts=tr.cur_pts.copy())


def _cpu_pool(payloads):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(payloads)) as pool:
        return pool.map_async(_cpu_worker, payloads).get(timeout=900)       # a crashed worker must not hang the bench


def _cpu_make_data(synth, n_streams, n_frames, dev):
    return make_data(synth, n_streams, n_frames, 0, dev)


def cpu_baselines(args, bn=None, n_time=None):
    """BASELINE.md section 3, both baselines, on the host cores of this box, >= 300 timed frames per stream:
      (2) throughput-fair (the headline `value`): independent single-threaded streams, one process per core, aggregate frames/s.  The
          reference's marginalisation spawns 4 pthreads per stream (marginalization_factor.hpp:18), so `cores` streams oversubscribe;
          the arm is therefore run with cores, cores/2 and cores/4 streams and the BEST aggregate is reported;
      (1) reference-faithful single stream: cv2 with its default thread pool, Ceres num_threads 1, 4 marginalisation pthreads.
    Each with the split front end / ceres::Solve / marginalisation, for the default -O2 oracle build and for -O3 -march=x86-64-v3.
    Also returns parity_in_bench: the GPU pipeline's window states / tracker ids against these CPU runs on the same frames."""
    import torch
    synth = importlib.import_module("vins-mobile_b200.synth")
    cores = os.cpu_count() or 1
    prologue = FREQ * 11
    n_time = n_time or max(args.cpu_frames, FREQ)
    n_time -= n_time % FREQ
    n_frames = prologue + n_time + FREQ
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    frames, dt, acc, gyr, gt, _ = _cpu_make_data(synth, cores, n_frames, dev)
    fr = frames.cpu().numpy()
    del frames

    def payloads(n, threads, lib):
        return [(fr[:, b], dt[:, :, b], acc[:, :, b], gyr[:, :, b], gt[b], prologue, n_time, threads, lib) for b in range(n)]

    def summarise(res, n):
        tmax = max(r["total"] for r in res)
        tot = sum(r["total"] for r in res)
        return {"value": n * n_time / tmax, "streams": n, "per_stream_frames_per_s": n_time * n / tot,
                "split_fraction": {k: sum(r[k] for r in res) / tot for k in ("front_end", "ceres_solve", "marginalise", "process_imu")},
                "ms_per_frame_per_stream": {k: 1e3 * sum(r[k] for r in res) / (n * n_time) for k in ("front_end", "ceres_solve", "marginalise")}}

    out = {}
    best = None
    res_full = None
    for lib, tag in ((None, "O2"), ("libvins_ref_o3.so", "O3_x86-64-v3")):
        if lib and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", lib)):
            continue
        runs = []
        for n in sorted({cores, max(cores // 2, 1), max(cores // 4, 1)}, reverse=True):
            res = _cpu_pool(payloads(n, 1, lib))
            if lib is None and n == cores:
                res_full = res
            runs.append(summarise(res, n))
        out[f"throughput_fair_{tag}"] = {"runs": runs, "best": max(runs, key=lambda r: r["value"])}
        single = _cpu_pool(payloads(1, -1, lib))       # cv2.setNumThreads(-1) restores OpenCV's default pool
        out[f"single_stream_faithful_{tag}"] = summarise(single, 1)
        cand = out[f"throughput_fair_{tag}"]["best"]
        if best is None or cand["value"] > best[1]["value"]:
            best = (tag, cand)
    cpu = {"value": best[1]["value"], "unit": "frames/s", "cores": cores, "kind": "reference", "build": best[0], "streams": best[1]["streams"],
           "sample": f"{n_time} timed frames per stream after {prologue} untimed (window fill + first solve), same frames / IMU as GPU streams 0..{cores - 1}; "
                     "front end = OpenCV 4.13 binary (the reference's OpenCV fork is not vendored) driven by the restated readImage, back end = reference "
                     "factor code + vendored Ceres 1.12 (oracle/_ref); value = best aggregate over {cores, cores/2, cores/4} concurrent streams and both builds",
           "per_stream_frames_per_s": best[1]["per_stream_frames_per_s"], "detail": out}
    parity = None
    if bn is not None and res_full is not None:
        # the GPU pipeline on the very same frames (frames are a deterministic function of (stream id, frame index))
        last = prologue + n_time - 1
        nb = min(cores, bn.B)
        if last < bn.n_frames_parity_cap():
            st, trk = bn.parity_states_on(fr[:last + 1, :nb], dt, acc, gyr, gt, nb, last)
            ids_equal, pts_equal, worst = True, True, 0.0
            for b in range(nb):
                ids_equal &= bool(np.array_equal(trk[b]["ids"], res_full[b]["ids"]))
                pts_equal &= bool(np.array_equal(trk[b]["ids"], res_full[b]["ids"]) and
                                  np.array_equal(trk[b]["pts"].view(np.uint32), res_full[b]["pts"].view(np.uint32)))
                g, r = st[b], res_full[b]["state"]
                for sl in (slice(0, 3), slice(7, 10)):
                    worst = max(worst, float(np.abs(g[:, sl] - r[:, sl]).max() / max(np.abs(r[:, sl]).max(), 1e-12)))
                q = np.minimum(np.abs(g[:, 3:7] - r[:, 3:7]).max(), np.abs(g[:, 3:7] + r[:, 3:7]).max())
                worst = max(worst, float(q))
            parity = {"streams": nb, "frames": last + 1, "ids_equal": ids_equal, "points_bitwise_equal": pts_equal, "max_rel_err": worst,
                      "what": "GPU pipeline (device path) vs the CPU reference arm (-O2 build) on the same frames: tracker ids / positions after the last "
                              "frame, window P, V (relative to the largest entry) and Q (absolute) after the last keyframe"}
    return cpu, parity


def run_reference(args):
    rank, local, world = _rank_world()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    args.cpu_frames = min(args.cpu_frames, max(FREQ, (args.steps + args.warmup) * 10))
    cpu, _ = cpu_baselines(args, None)
    line = {"impl": "reference", "metric": "VIO frames/sec (640x480, 150 feats, 10-KF window)", "value": cpu["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cpu["streams"] / cpu["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32 front end, f64 back end", "data": "synthetic",
            "config": {"workload": "independent streams, 640x480@30fps + 200 Hz IMU, 150 feats, 10-KF window, FREQ=3 "
                                   "(one stream per host core or fewer, best aggregate)", "streams": cpu["streams"], "freq": FREQ, "host_cores": cores},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default c2 = the headline workload)")
    ap.add_argument("--batch", type=int, default=0, help="streams per GPU (default: the configuration's)")
    ap.add_argument("--cpu-frames", type=int, default=300, help="timed frames per stream of the cpu_baseline legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the variant / extra-configuration lines (profiling runs)")
    ap.add_argument("--no-overlap", action="store_true", help="front end and back end on ONE CUDA stream (no overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
