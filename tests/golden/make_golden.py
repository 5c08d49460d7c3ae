"""Generates the committed golden vectors (run HERE, where /root/reference and cv2 are available):

  frontend_golden.npz : small image pair + what the OpenCV 4.13 binary returns for the four front-end primitives
                        (pyrDown chain, goodFeaturesToTrack with a disc mask, calcOpticalFlowPyrLK, findFundamentalMat/RANSAC)
  backend_golden.npz  : inputs and outputs of the reference's own factor code + vendored Ceres 1.12 (oracle/_ref):
                        pre-integration, IMUFactor, ProjectionFactor, and the window state after every keyframe of a 16-keyframe
                        synthetic track sequence (synth.make_tracks(seed 3), regenerated deterministically by the tests).

  clahe_golden.npz    : cv2 CLAHE (clip 3, 8x8 tiles) of a 160x128 image
  pnp_golden.npz      : the reference's motion-only PnP tracker (oracle/pnp_ref.cpp) on a synthetic sequence

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cv2
import backend_oracle as bo
import frontend_oracle as fo
from be_common import Quiet, drive, drive_pnp
from conftest import texture_pair, two_view_points

synth = importlib.import_module("vins-mobile_b200.synth")
abi = importlib.import_module("vins-mobile_b200.abi")


def frontend():
    img0, img1 = texture_pair(seed=11, rows=160, cols=128)
    lv = [img0]
    for _ in range(3):
        lv.append(cv2.pyrDown(lv[-1]))
    kept = np.array([[30.2, 40.7], [90.5, 100.5], [64.0, 20.0]], np.float32)
    mask = np.full(img0.shape, 255, np.uint8)
    for c in kept:
        cv2.circle(mask, (int(np.rint(c[0])), int(np.rint(c[1]))), 30, 0, -1)
    corners = fo.cv2_good_features(img0, mask, 20)
    pts = fo.cv2_good_features(img0, None, 40, min_dist=10.0)
    nxt, st = fo.cv2_lk_track(img0, img1, pts)
    x1, x2 = two_view_points(77, n=60, nout=6)
    rmask = fo.cv2_find_fundamental(x1, x2)
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), img0=img0, img1=img1, l1=lv[1], l2=lv[2], l3=lv[3], kept=kept, corners=corners,
                        eig=cv2.cornerMinEigenVal(img0, 3, ksize=3), lk_pts=pts, lk_next=nxt, lk_status=st, f_x1=x1, f_x2=x2, f_mask=rmask,
                        cv2_version=np.array(cv2.__version__))


def backend():
    cfg = abi.default_config(batch=1, max_cnt=80)
    r = np.random.default_rng(5)
    n = 20
    dt = np.full(n, 0.005)
    acc = np.array([0.3, -0.2, 9.8]) + r.normal(0, 0.5, (n, 3))
    gyr = np.array([0.05, -0.1, 0.2]) + r.normal(0, 0.1, (n, 3))
    ba, bg = np.array([0.02, -0.01, 0.03]), np.array([0.001, 0.002, -0.003])
    pqv, jac, cov, sdt = bo.preintegrate(dt, acc, gyr, acc[0], gyr[0], ba, bg)

    def pose():
        q = np.array([0, 0, 0, 1.0]) + 0.2 * r.normal(0, 1, 4)
        return np.concatenate([r.normal(0, 0.3, 3), q / np.linalg.norm(q)])
    pi, pj = pose(), pose()
    sbi = np.concatenate([r.normal(0, 0.5, 3), r.normal(0, 0.02, 3), r.normal(0, 0.002, 3)])
    sbj = np.concatenate([r.normal(0, 0.5, 3), r.normal(0, 0.02, 3), r.normal(0, 0.002, 3)])
    imu_r, imu_J = bo.imu_factor(pqv, jac, cov, sdt, ba, bg, pi, sbi, pj, sbj)
    pts_i = np.array([0.11, -0.23, 1.0]); pts_j = np.array([0.07, -0.19, 1.0])
    pr, pJ = bo.projection_factor(cfg.fx, np.array(cfg.tic[:]), np.array(cfg.ric[:]), pts_i, pts_j, pi, pj, 0.31)
    tr = synth.make_tracks(3, 16, max_cnt=80)
    est = bo.RefEstimator(cfg)
    states, infos = [], []
    for k in range(16):
        with Quiet():
            drive(est, tr, k, cfg.window_size)
        s = est.state()
        states.append(np.concatenate([s["P"], s["Q"], s["V"], s["Ba"], s["Bg"]], 1))
        i = est.info()
        infos.append([i["marg_flag"], i["n_feat"], i["n_proj"], i["iters"], i["prior_n"]])
    np.savez_compressed(os.path.join(HERE, "backend_golden.npz"), dt=dt, acc=acc, gyr=gyr, ba=ba, bg=bg, pqv=pqv, jac=jac, cov=cov, sdt=sdt, pi=pi, pj=pj,
                        sbi=sbi, sbj=sbj, imu_r=imu_r, imu_J=imu_J, pts_i=pts_i, pts_j=pts_j, inv_dep=0.31, proj_r=pr, proj_J=pJ,
                        states=np.array(states), infos=np.array(infos), track_seed=3, n_kf=16, max_cnt=80)


def clahe():
    """clahe_golden.npz: what the OpenCV 4.13 binary returns for createCLAHE(clipLimit=3, tileGridSize=(8,8)).apply on a 160x128 image."""
    img0, _ = texture_pair(seed=11, rows=160, cols=128)
    np.savez_compressed(os.path.join(HERE, "clahe_golden.npz"), img=img0, out=fo.cv2_clahe(img0, 3.0, (8, 8)), cv2_version=np.array(cv2.__version__))


def pnp():
    """pnp_golden.npz: the reference's motion-only PnP tracker (vins_pnp.cpp via oracle/pnp_ref.cpp) on synth.make_pnp_sequence(seed 3):
    window state after every camera frame, plus one PerspectiveFactor evaluation (perspective_factor.cpp:16-67)."""
    cfg = abi.default_config(batch=1, max_cnt=150)
    seq = synth.make_pnp_sequence(3, 16)
    with Quiet():
        h = bo.RefPnP(cfg)
    states, last_t = [], 0.0
    for k in range(16):
        with Quiet():
            last_t = drive_pnp(h, seq, k, last_t)
            s = h.state()
        states.append(np.concatenate([s["P"], s["R"].reshape(-1, 9), s["V"], s["headers"][:, None], s["find_solved"][:, None].astype(float)], 1))
    r = np.random.default_rng(9)
    q = np.array([0.1, -0.2, 0.05, 1.0]); q /= np.linalg.norm(q)
    pose = np.concatenate([r.normal(0, 0.3, 3), q])
    qe = np.array([1.0, 0.0, 0.0, 0.0])                       # ric = ypr2R(0, 0, 180 deg): quaternion (x, y, z, w) = (1, 0, 0, 0)
    ex = np.concatenate([np.array(cfg.tic[:]), qe])
    pos = pose[:3] + np.array([0.4, -0.3, -3.0]); ob = np.array([0.12, -0.07])
    pr, pJp, pJe = bo.perspective_factor(ob, pos, 7, cfg.fx, pose, ex)
    np.savez_compressed(os.path.join(HERE, "pnp_golden.npz"), states=np.array(states), seed=3, n_frames=16, pf_pose=pose, pf_ex=ex, pf_pos=pos, pf_obs=ob,
                        pf_track=7, pf_r=pr, pf_Jp=pJp, pf_Je=pJe)


if __name__ == "__main__":
    frontend()
    backend()
    clahe()
    pnp()
    print("golden vectors written to", HERE)
