"""Synthetic camera + IMU streams -- DATA GENERATION ONLY.

One stream = an analytic 6-DoF trajectory over a textured, NON-planar surface (a paraboloid
"bowl": a planar scene makes the fundamental matrix degenerate, which would make RANSAC-F
inlier sets implementation-defined), rendered by exact ray/quadric intersection, plus a 200 Hz
IMU sampled from the analytic derivatives of the same trajectory.

Conventions follow the reference: image is ROW x COL = 640 x 480 u8 (feature_tracker.hpp:26-27),
intrinsics/extrinsics of the iPhone7P entry (global_param.cpp:26-39, global_param.hpp:23-25),
world z up with g = (0,0,9.805) subtracted in the body-to-world mid-point propagation
(VINS.cpp:361-366), so the accelerometer reads R^T (a_w + g).

Everything is written with torch ops so the same code renders on the CPU (tests, here) and on
the GPU (bench).  Nothing in here is timed and nothing in here is part of the hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

GRAVITY = 9.805


@dataclass
class Camera:
    rows: int = 640
    cols: int = 480
    fx: float = 526.600
    fy: float = 526.678
    cx: float = 243.481
    cy: float = 315.280
    tic: tuple = (0.0, 0.092, 0.01)
    # ric = ypr2R(0,0,180deg) = Rx(pi)  (global_param.hpp:23-25, VINS.cpp:55)
    ric: tuple = (1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, -1.0)

    def scaled(self, rows, cols):
        s = cols / self.cols
        return Camera(rows, cols, self.fx * s, self.fy * s, self.cx * s, self.cy * (rows / self.rows),
                      self.tic, self.ric)


def _rot_zyx(y, p, r):
    """Rz(y) Ry(p) Rx(r) as in Utility::ypr2R (utility.hpp:96-121), radians, batched."""
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    R = np.empty(np.shape(y) + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def rot_to_quat_xyzw(R):
    """Rotation matrix -> unit quaternion (x,y,z,w), w >= 0."""
    R = np.asarray(R, np.float64)
    out = np.empty(R.shape[:-2] + (4,))
    flat_R = R.reshape(-1, 3, 3)
    flat_o = out.reshape(-1, 4)
    for i, m in enumerate(flat_R):
        t = np.trace(m)
        if t > 0:
            s = math.sqrt(t + 1.0) * 2
            w = 0.25 * s
            x = (m[2, 1] - m[1, 2]) / s
            y = (m[0, 2] - m[2, 0]) / s
            z = (m[1, 0] - m[0, 1]) / s
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            w = (m[2, 1] - m[1, 2]) / s
            x = 0.25 * s
            y = (m[0, 1] + m[1, 0]) / s
            z = (m[0, 2] + m[2, 0]) / s
        elif m[1, 1] > m[2, 2]:
            s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            w = (m[0, 2] - m[2, 0]) / s
            x = (m[0, 1] + m[1, 0]) / s
            y = 0.25 * s
            z = (m[1, 2] + m[2, 1]) / s
        else:
            s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            w = (m[1, 0] - m[0, 1]) / s
            x = (m[0, 2] + m[2, 0]) / s
            y = (m[1, 2] + m[2, 1]) / s
            z = 0.25 * s
        q = np.array([x, y, z, w])
        if w < 0:
            q = -q
        flat_o[i] = q / np.linalg.norm(q)
    return out


class Trajectory:
    """position = sum of 3 sinusoids per axis; yaw/pitch/roll sinusoids <= ~10 deg (seed 2000+s)."""

    def __init__(self, stream_id: int, speed: float = 1.0):
        r = np.random.default_rng(2000 + stream_id)
        self.amp = r.uniform(0.08, 0.30, (3, 3)) * np.array([[1.0], [1.0], [0.5]])
        self.frq = r.uniform(0.2, 0.6, (3, 3)) * speed
        self.phs = r.uniform(0, 2 * np.pi, (3, 3))
        self.aamp = np.deg2rad(r.uniform(3.0, 9.0, 3))
        self.afrq = r.uniform(0.15, 0.45, 3) * speed
        self.aphs = r.uniform(0, 2 * np.pi, 3)

    def pos(self, t, d=0):
        t = np.asarray(t, np.float64)[..., None, None]
        w = 2 * np.pi * self.frq
        ph = w * t + self.phs + d * np.pi / 2
        return (self.amp * w ** d * np.sin(ph)).sum(-1)

    def ypr(self, t, d=0):
        t = np.asarray(t, np.float64)[..., None]
        w = 2 * np.pi * self.afrq
        return self.aamp * w ** d * np.sin(w * t + self.aphs + d * np.pi / 2)

    def R(self, t):
        a = self.ypr(t)
        return _rot_zyx(a[..., 0], a[..., 1], a[..., 2])

    def omega_body(self, t):
        a = self.ypr(t)
        da = self.ypr(t, 1)
        p, r = a[..., 1], a[..., 2]
        yd, pd, rd = da[..., 0], da[..., 1], da[..., 2]
        wx = rd - yd * np.sin(p)
        wy = pd * np.cos(r) + yd * np.sin(r) * np.cos(p)
        wz = -pd * np.sin(r) + yd * np.cos(r) * np.cos(p)
        return np.stack([wx, wy, wz], -1)


class Surface:
    """z = z0 + k (x^2 + y^2), textured over [-ext, ext]^2 (seed 1000+s)."""

    def __init__(self, stream_id: int, z0=-3.2, k=0.10, ext=5.0, tex_res=2048, device="cpu"):
        self.z0, self.k, self.ext = z0, k, ext
        g = torch.Generator().manual_seed(1000 + stream_id)
        cells = 200
        small = torch.randint(0, 256, (1, 1, cells, cells), generator=g).float()
        tex = torch.nn.functional.interpolate(small, size=(tex_res, tex_res), mode="bicubic", align_corners=False)
        ks = 9
        ax = torch.arange(ks) - ks // 2
        gk = torch.exp(-(ax.float() ** 2) / (2 * 2.0 ** 2))
        gk = gk / gk.sum()
        tex = torch.nn.functional.conv2d(torch.nn.functional.pad(tex, (ks // 2,) * 4, mode="reflect"),
                                         gk.view(1, 1, 1, ks))
        tex = torch.nn.functional.conv2d(tex, gk.view(1, 1, ks, 1))
        # stretch contrast back to the full u8 range
        lo, hi = tex.quantile(0.01), tex.quantile(0.99)
        tex = ((tex - lo) / (hi - lo)).clamp(0, 1) * 255.0
        self.tex = tex.to(device)
        self.device = device

    def render(self, cam: Camera, R_wc: np.ndarray, c_w: np.ndarray) -> torch.Tensor:
        """Render u8 (T, rows, cols) images for camera-to-world rotations (T,3,3) and centres (T,3)."""
        dev = self.device
        T = R_wc.shape[0]
        v, u = torch.meshgrid(torch.arange(cam.rows, device=dev, dtype=torch.float64),
                              torch.arange(cam.cols, device=dev, dtype=torch.float64), indexing="ij")
        dc = torch.stack([(u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, torch.ones_like(u)], -1)   # (H,W,3)
        Rt = torch.as_tensor(R_wc, dtype=torch.float64, device=dev)
        c = torch.as_tensor(c_w, dtype=torch.float64, device=dev)
        out = torch.empty((T, cam.rows, cam.cols), dtype=torch.uint8, device=dev)
        for i in range(T):
            d = dc @ Rt[i].T                                   # world ray directions
            cx_, cy_, cz_ = c[i, 0], c[i, 1], c[i, 2]
            A = self.k * (d[..., 0] ** 2 + d[..., 1] ** 2)
            B = 2 * self.k * (cx_ * d[..., 0] + cy_ * d[..., 1]) - d[..., 2]
            C = self.k * (cx_ ** 2 + cy_ ** 2) + self.z0 - cz_
            disc = (B * B - 4 * A * C).clamp_min(0)
            s = -2 * C / (B + torch.sqrt(disc))
            px = cx_ + s * d[..., 0]
            py = cy_ + s * d[..., 1]
            grid = torch.stack([px / self.ext, py / self.ext], -1).float().unsqueeze(0)
            img = torch.nn.functional.grid_sample(self.tex, grid, mode="bilinear", padding_mode="reflection",
                                                  align_corners=False)
            out[i] = img[0, 0].round().clamp(0, 255).to(torch.uint8)
        return out

    def depth_along(self, c_w, d_w):
        """Ray parameter s of the hit (numpy, for ground-truth landmarks in tests)."""
        A = self.k * (d_w[..., 0] ** 2 + d_w[..., 1] ** 2)
        B = 2 * self.k * (c_w[..., 0] * d_w[..., 0] + c_w[..., 1] * d_w[..., 1]) - d_w[..., 2]
        C = self.k * (c_w[..., 0] ** 2 + c_w[..., 1] ** 2) + self.z0 - c_w[..., 2]
        return -2 * C / (B + np.sqrt(np.maximum(B * B - 4 * A * C, 0)))


@dataclass
class Stream:
    images: torch.Tensor        # (T, rows, cols) u8
    frame_t: np.ndarray         # (T,)
    imu_t: np.ndarray           # (M,)  t = (j+1)/imu_hz
    acc: np.ndarray             # (M,3)
    gyr: np.ndarray             # (M,3)
    P: np.ndarray               # (T,3)  ground-truth IMU position in world
    R: np.ndarray               # (T,3,3)
    V: np.ndarray               # (T,3)
    cam: Camera


def make_stream(stream_id: int, n_frames: int, cam: Camera | None = None, fps: float = 30.0, imu_hz: float = 200.0,
                device: str = "cpu", render: bool = True, speed: float = 1.0,
                gyr_noise: float = 0.002, acc_noise: float = 0.05, surface: Surface | None = None) -> Stream:
    cam = cam or Camera()
    traj = Trajectory(stream_id, speed)
    ft = np.arange(n_frames) / fps
    P = traj.pos(ft)
    V = traj.pos(ft, 1)
    R = traj.R(ft)
    n_imu = int(round(ft[-1] * imu_hz)) if n_frames > 1 else 0
    it = (np.arange(n_imu) + 1) / imu_hz
    rn = np.random.default_rng(3000 + stream_id)
    Ri = traj.R(it)
    a_w = traj.pos(it, 2) + np.array([0.0, 0.0, GRAVITY])
    acc = np.einsum("nji,nj->ni", Ri, a_w) + rn.normal(0, acc_noise, (n_imu, 3))
    gyr = traj.omega_body(it) + rn.normal(0, gyr_noise, (n_imu, 3))
    ric = np.array(cam.ric).reshape(3, 3)
    tic = np.array(cam.tic)
    if render:
        surf = surface or Surface(stream_id, device=device)
        R_wc = R @ ric
        c_w = P + R @ tic
        images = surf.render(cam, R_wc, c_w)
    else:
        images = torch.zeros((0, cam.rows, cam.cols), dtype=torch.uint8)
    return Stream(images, ft, it, acc, gyr, P, R, V, cam)


def make_tracks(stream_id: int, n_kf: int, max_cnt: int = 150, cam: Camera | None = None, kf_dt: float = 0.1,
                imu_hz: float = 200.0, px_noise: float = 0.1, speed: float = 1.0,
                gyr_noise: float = 0.002, acc_noise: float = 0.05):
    """Feature tracks without rendering (for back-end tests): landmarks on the Surface are projected into every keyframe,
    leave when they exit the image (1-px border, like inBorder) and are replaced by fresh ids, exactly the id discipline of
    FeatureTracker (monotone ids, kept points first).  Returns dict with per-keyframe (ids int32, xyz float64 (n,3)),
    the IMU samples between keyframes, and ground truth P/R/V at the keyframes."""
    cam = cam or Camera()
    traj = Trajectory(stream_id, speed)
    surf_z0, surf_k = -3.2, 0.10
    rng = np.random.default_rng(4000 + stream_id)
    ric = np.array(cam.ric).reshape(3, 3)
    tic = np.array(cam.tic)
    t_kf = np.arange(n_kf) * kf_dt
    P = traj.pos(t_kf); V = traj.pos(t_kf, 1); R = traj.R(t_kf)
    per = int(round(kf_dt * imu_hz))
    it = (np.arange((n_kf - 1) * per) + 1) / imu_hz
    Ri = traj.R(it)
    a_w = traj.pos(it, 2) + np.array([0.0, 0.0, GRAVITY])
    acc = np.einsum("nji,nj->ni", Ri, a_w) + rng.normal(0, acc_noise, (len(it), 3))
    gyr = traj.omega_body(it) + rng.normal(0, gyr_noise, (len(it), 3))

    def raycast(c, d):
        A = surf_k * (d[..., 0] ** 2 + d[..., 1] ** 2)
        B = 2 * surf_k * (c[0] * d[..., 0] + c[1] * d[..., 1]) - d[..., 2]
        C = surf_k * (c[0] ** 2 + c[1] ** 2) + surf_z0 - c[2]
        s = -2 * C / (B + np.sqrt(np.maximum(B * B - 4 * A * C, 0)))
        return c + s[..., None] * d

    ids_live = np.zeros(0, np.int64)
    X_live = np.zeros((0, 3))
    next_id = 0
    frames = []
    for k in range(n_kf):
        R_wc = R[k] @ ric
        c_w = P[k] + R[k] @ tic
        if len(X_live):
            pc = (X_live - c_w) @ R_wc
            u = cam.fx * pc[:, 0] / pc[:, 2] + cam.cx
            v = cam.fy * pc[:, 1] / pc[:, 2] + cam.cy
            ok = (pc[:, 2] > 0.1) & (np.rint(u) >= 1) & (np.rint(u) < cam.cols - 1) & (np.rint(v) >= 1) & (np.rint(v) < cam.rows - 1)
            ids_live, X_live, u, v = ids_live[ok], X_live[ok], u[ok], v[ok]
        else:
            u = v = np.zeros(0)
        n_new = max_cnt - len(ids_live)
        if n_new > 0:
            un = rng.uniform(2, cam.cols - 3, n_new)
            vn = rng.uniform(2, cam.rows - 3, n_new)
            d = np.stack([(un - cam.cx) / cam.fx, (vn - cam.cy) / cam.fy, np.ones(n_new)], -1) @ R_wc.T
            Xn = raycast(c_w, d)
            ids_live = np.concatenate([ids_live, next_id + np.arange(n_new)])
            next_id += n_new
            X_live = np.concatenate([X_live, Xn])
            u = np.concatenate([u, un]); v = np.concatenate([v, vn])
        un_ = u + rng.normal(0, px_noise, len(u))
        vn_ = v + rng.normal(0, px_noise, len(v))
        xyz = np.stack([(un_ - cam.cx) / cam.fx, (vn_ - cam.cy) / cam.fy, np.ones(len(u))], -1)
        frames.append((ids_live.astype(np.int32).copy(), xyz))
    return dict(frames=frames, t_kf=t_kf, imu_t=it, acc=acc, gyr=gyr, per=per, P=P, R=R, V=V, cam=cam)


def make_pnp_sequence(stream_id: int, n_frames: int = 16, n_lm: int = 60, cam: Camera | None = None, fps: float = 30.0, imu_hz: float = 200.0,
                      px_noise: float = 0.1, lag: int = 3):
    """Input of the motion-only PnP tracker (FeatureTracker::solveVinsPnP, feature_tracker.cpp:107-160): a fixed set of landmarks
    with known world positions seen in every camera frame (id, normalised observation, position, track_num), the IMU samples between
    frames, and the estimator results (`solved_vins`) that reach the tracker `lag` frames late.  Ground truth P/R/V per frame."""
    cam = cam or Camera()
    traj = Trajectory(stream_id)
    rng = np.random.default_rng(5000 + stream_id)
    t = np.arange(n_frames) / fps
    P, V, R = traj.pos(t), traj.pos(t, 1), traj.R(t)
    ric = np.array(cam.ric).reshape(3, 3)
    tic = np.array(cam.tic)
    c0, Rwc0 = P[0] + R[0] @ tic, R[0] @ ric
    u = rng.uniform(0.25 * cam.cols, 0.75 * cam.cols, n_lm)
    v = rng.uniform(0.25 * cam.rows, 0.75 * cam.rows, n_lm)
    d = np.stack([(u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, np.ones(n_lm)], -1) @ Rwc0.T
    X = c0 + d * rng.uniform(2.5, 4.5, (n_lm, 1))
    imu_t = (np.arange(int(round((n_frames - 1) / fps * imu_hz))) + 1) / imu_hz
    acc = np.einsum("nji,nj->ni", traj.R(imu_t), traj.pos(imu_t, 2) + np.array([0.0, 0.0, GRAVITY]))
    gyr = traj.omega_body(imu_t)
    obs = []
    for k in range(n_frames):
        pc = (X - (P[k] + R[k] @ tic)) @ (R[k] @ ric)
        obs.append(pc[:, :2] / pc[:, 2:3] + rng.normal(0, px_noise / cam.fx, (n_lm, 2)))
    return dict(t=t, P=P, V=V, R=R, X=X, ids=np.arange(n_lm, dtype=np.int32), track_num=np.full(n_lm, 10, np.int32), obs=np.array(obs),
                imu_t=imu_t, acc=acc, gyr=gyr, lag=lag)


def make_align_case(sid, n, scale=2.5, gyro_bias=(0.01, -0.02, 0.015), rot_noise=2e-3, pos_noise=2e-3, max_frames=None, max_imu=None):
    """Inputs of VisualIMUAlignment (initial_aligment.cpp:222-229) for one synthetic stream: n camera frames as VINS::solveInitial
    leaves them after the global SfM (VINS.cpp:889-905: ImageFrame::R = body attitude in the SfM frame c0, ImageFrame::T = camera
    position in c0, up to the unknown scale), the raw IMU samples of every frame interval (constant gyroscope bias added), and the
    ground truth (scale, gravity in c0).  c0 is rotated at random against the world, so gravity is not axis aligned."""
    tr = make_tracks(sid, n + 1, max_cnt=8)
    cfg_tic = np.array([0.0, 0.092, 0.01])
    rng = np.random.default_rng(1000 + sid)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    Rc0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    per = tr["per"]
    F = max_frames or n
    M = max_imu or per
    R = np.zeros((F, 3, 3)); T = np.zeros((F, 3)); counts = np.zeros(F, np.int32); imu0 = np.zeros((F, 6)); imu = np.zeros((F, M, 7))
    R[:] = np.eye(3)
    Pc0 = tr["P"][0] + tr["R"][0] @ cfg_tic
    for k in range(n):
        th = rng.normal(0, rot_noise, 3)
        dR = np.eye(3) + np.array([[0, -th[2], th[1]], [th[2], 0, -th[0]], [-th[1], th[0], 0]])
        u, _, vt = np.linalg.svd(dR)
        R[k] = Rc0 @ tr["R"][k] @ (u @ vt)
        T[k] = Rc0 @ (tr["P"][k] + tr["R"][k] @ cfg_tic - Pc0) / scale + rng.normal(0, pos_noise, 3)
        if k > 0:
            sl = slice((k - 1) * per, k * per)
            dts = np.diff(np.concatenate([[tr["t_kf"][k - 1]], tr["imu_t"][sl]]))
            counts[k] = per
            imu[k, :per, 0] = dts
            imu[k, :per, 1:4] = tr["acc"][sl]
            imu[k, :per, 4:7] = tr["gyr"][sl] + np.asarray(gyro_bias)
            j = max(0, (k - 1) * per - 1)
            imu0[k, :3] = tr["acc"][j]; imu0[k, 3:] = tr["gyr"][j] + np.asarray(gyro_bias)
    g_c0 = Rc0 @ np.array([0.0, 0.0, 9.805])
    return dict(n=n, R=R, T=T, counts=counts, imu0=imu0, imu=imu, tic=cfg_tic, scale=scale, g=g_c0, gyro_bias=np.asarray(gyro_bias))
