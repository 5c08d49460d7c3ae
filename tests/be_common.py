"""Shared drivers for the back-end parity tests: feed identical IMU samples / image_msg / initial window to the reference
estimator (oracle/_ref/libvins_ref.so) and to the CUDA back end (C-ABI)."""
import importlib
import os
import sys

import numpy as np

synth = importlib.import_module("vins-mobile_b200.synth")


class Quiet:
    """Silence the reference's printf chatter (marginalization_factor.cpp prints on every call) at the fd level."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def drive(est, tr, k, W, single=True):
    """Advance estimator `est` (RefEstimator or api.BackEnd batch 1) by keyframe k of track set `tr`."""
    per = tr["per"]
    if k > 0:
        sl = slice((k - 1) * per, k * per)
        dts = np.diff(np.concatenate([[tr["t_kf"][k - 1]], tr["imu_t"][sl]]))
        if hasattr(est, "B"):
            est.process_imu(dts[:, None], tr["acc"][sl][:, None, :], tr["gyr"][sl][:, None, :])
        else:
            for d, a, g in zip(dts, tr["acc"][sl], tr["gyr"][sl]):
                est.process_imu(d, a, g)
    if k == W:
        fr = list(range(W + 1))
        Q = synth.rot_to_quat_xyzw(tr["R"][fr])
        if hasattr(est, "B"):
            est.set_init_window(tr["P"][fr][None], Q[None], tr["V"][fr][None], np.zeros((1, 3)), np.zeros((1, 3)))
        else:
            est.set_init_window(tr["P"][fr], Q, tr["V"][fr], np.zeros(3), np.zeros(3))
    ids, xyz = tr["frames"][k]
    if hasattr(est, "B"):
        est.process_image_single(ids, xyz, tr["t_kf"][k])
    else:
        est.process_image(ids, xyz, tr["t_kf"][k])


def outlier_tracks(sid, n_kf, max_cnt=150, n_out=6, seed=0, mag=0.15):
    """synth.make_tracks with `n_out` tracks turned into gross outliers (the observation slides across the image, inconsistent with any
    static point): their inverse depth goes negative in a solve, removeFailures() (feature_manager.cpp:289-298) erases them while the
    front end keeps publishing the id, and the id is appended again BEHIND larger ids -- the reference's feature list stops being
    sorted by id (from keyframe 16 on for stream 0).  Returns (tracks, outlier ids)."""
    tr = synth.make_tracks(sid, n_kf, max_cnt=max_cnt)
    rng = np.random.default_rng(seed)
    born = {}
    for k, (ids, _xyz) in enumerate(tr["frames"]):
        for i in ids:
            born.setdefault(int(i), k)
    cand = [i for i, k in born.items() if 12 <= k <= 16]
    pick = rng.choice(cand, n_out, replace=False)
    drift = {int(i): rng.normal(0, mag, 2) for i in pick}
    for k, (ids, xyz) in enumerate(tr["frames"]):
        for j, i in enumerate(ids):
            i = int(i)
            if i in drift and k > born[i]:
                xyz[j, :2] += drift[i] * (k - born[i])
    return tr, pick


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def quat_err(qa, qb):
    """max over frames of min(|qa-qb|, |qa+qb|) (q and -q are the same rotation)."""
    qa, qb = np.asarray(qa), np.asarray(qb)
    return float(np.minimum(np.abs(qa - qb).max(-1), np.abs(qa + qb).max(-1)).max())


def drive_pnp(pnp, seq, k, last_t):
    """One camera frame of FeatureTracker::solveVinsPnP (feature_tracker.cpp:107-160): the estimator result of frame k - lag (if any),
    the IMU samples since the previous frame, then the frame's features.  Returns the new `last_t`."""
    import numpy as np
    if k >= seq["lag"]:
        j = k - seq["lag"]
        pnp.set_init(seq["t"][j], seq["P"][j], seq["R"][j], seq["V"][j], np.zeros(3), np.zeros(3))
    sel = (seq["imu_t"] > last_t + 1e-9) & (seq["imu_t"] <= seq["t"][k] + 1e-9)
    tt = last_t
    for ti, a, g in zip(seq["imu_t"][sel], seq["acc"][sel], seq["gyr"][sel]):
        pnp.process_imu(ti - tt, a, g)
        tt = ti
    pnp.process_image(seq["ids"], seq["obs"][k], seq["X"], seq["track_num"], seq["t"][k], True)
    return seq["t"][k]


def align_case(sid, n, **kw):
    """synth.make_align_case: inputs of VisualIMUAlignment for one synthetic stream (+ ground truth)."""
    return synth.make_align_case(sid, n, **kw)


def sfm_window(tr, k, W, scale=2.5, seed=5, mirrored=False):
    """What VINS::solveInitial holds after the global SfM (VINS.cpp:889-905) for the window ending at keyframe k: ImageFrame::R (body
    attitude in the SfM frame c0) and ImageFrame::T (camera position in c0, unknown scale) of keyframes k-W..k, from the synthetic ground
    truth; c0 is rotated at random against the world.  mirrored=True negates T (a reflected reconstruction: the alignment must reject it)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    Rc0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    tic = np.array([0.0, 0.092, 0.01])
    fr = list(range(k - W, k + 1))
    R = np.stack([Rc0 @ tr["R"][i] for i in fr])
    Pc = np.stack([tr["P"][i] + tr["R"][i] @ tic for i in fr])
    T = (Pc - Pc[0]) @ Rc0.T / scale
    return R, (-T if mirrored else T)


def sfm_frames(tr, headers, scale=2.5, seed=5):
    """sfm_window for an arbitrary list of frame timestamps (every frame of all_image_frame, keyframe or not): ImageFrame::R / T from the
    ground truth at those times."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    Rc0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    tic = np.array([0.0, 0.092, 0.01])
    fr = [int(np.argmin(np.abs(tr["t_kf"] - h))) for h in headers]
    R = np.stack([Rc0 @ tr["R"][i] for i in fr])
    Pc = np.stack([tr["P"][i] + tr["R"][i] @ tic for i in fr])
    return R, (Pc - Pc[0]) @ Rc0.T / scale


def drive_sfm(est, tr, k, W, sfm=None):
    """Like drive(), but the stream initialises from SfM poses (set_init_sfm) instead of a supplied window; one IMU sample precedes the
    first image (as on a device: the reference creates pre_integrations[0] from it, VINS.cpp:340-346).  sfm = (R, T) or None."""
    per = tr["per"]
    batched = hasattr(est, "B")
    if k == 0:
        if batched:
            est.process_imu(np.array([[0.005]]), tr["acc"][:1][:, None, :], tr["gyr"][:1][:, None, :])
        else:
            est.process_imu(0.005, tr["acc"][0], tr["gyr"][0])
    else:
        sl = slice((k - 1) * per, k * per)
        dts = np.diff(np.concatenate([[tr["t_kf"][k - 1]], tr["imu_t"][sl]]))
        if batched:
            est.process_imu(dts[:, None], tr["acc"][sl][:, None, :], tr["gyr"][sl][:, None, :])
        else:
            for d, a, g in zip(dts, tr["acc"][sl], tr["gyr"][sl]):
                est.process_imu(d, a, g)
    if sfm is not None:
        R, T = sfm
        if batched:
            est.set_init_sfm(R[None], T[None])
        else:
            est.set_init_sfm(R, T)
    ids, xyz = tr["frames"][k]
    if batched:
        est.process_image_single(ids, xyz, tr["t_kf"][k])
    else:
        est.process_image(ids, xyz, tr["t_kf"][k])


def init_scenario(est, cfg, name):
    """Two scripted initialisations from SfM poses, identical for the reference estimator and the CUDA back end (batch 1):
    'rejected_then_accepted' (stream 0): mirrored SfM at keyframe W (VisualIMUAlignment rejects it), correct SfM at W + 1;
    'non_keyframe' (stream 1): no SfM at W, frame W + 1 repeats keyframe W's measurements so that frame W + 2 triggers MARGIN_SECOND_NEW
    (the repeated frame leaves the window but stays in all_image_frame), SfM for all 12 frames of the map at W + 3."""
    W = cfg.window_size
    batched = hasattr(est, "B")
    if name == "rejected_then_accepted":
        tr = synth.make_tracks(0, 20, max_cnt=cfg.max_cnt)
        for k in range(W + 2):
            sfm = sfm_window(tr, k, W, mirrored=True) if k == W else (sfm_window(tr, k, W) if k == W + 1 else None)
            drive_sfm(est, tr, k, W, sfm)
        return tr
    assert name == "non_keyframe"
    tr = synth.make_tracks(1, 20, max_cnt=cfg.max_cnt)
    per = tr["per"]

    def imu_of(k):
        sl = slice((k - 1) * per, k * per)
        dts = np.diff(np.concatenate([[tr["t_kf"][k - 1]], tr["imu_t"][sl]]))
        if batched:
            est.process_imu(dts[:, None], tr["acc"][sl][:, None, :], tr["gyr"][sl][:, None, :])
        else:
            for d, a, g in zip(dts, tr["acc"][sl], tr["gyr"][sl]):
                est.process_imu(d, a, g)

    def image(k, src):
        ids, xyz = tr["frames"][src]
        if batched:
            est.process_image_single(ids, xyz, tr["t_kf"][k])
        else:
            est.process_image(ids, xyz, tr["t_kf"][k])

    for k in range(W + 1):
        drive_sfm(est, tr, k, W, None)
    imu_of(W + 1); image(W + 1, W)
    imu_of(W + 2); image(W + 2, W + 2)
    hd = est.init_frames()
    R, T = sfm_frames(tr, list(hd) + [tr["t_kf"][W + 3]])
    imu_of(W + 3)
    if batched:
        est.set_init_sfm_frames([len(R)], R[None], T[None])
    else:
        est.set_init_sfm_frames(R, T)
    image(W + 3, W + 3)
    return tr
