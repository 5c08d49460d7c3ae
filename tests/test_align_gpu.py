"""Visual-inertial alignment of the initialisation (linear half): vio_visual_imu_align (C-ABI, CUDA)
against the reference's UNMODIFIED VisualIMUAlignment (initial_aligment.cpp:222-229, compiled into oracle/_ref by oracle/Makefile)."""
import importlib

import numpy as np
import pytest

import be_common

api = importlib.import_module("vins-mobile_b200.api")
abi = importlib.import_module("vins-mobile_b200.abi")

TOL = 1e-10     # relative, max-norm: both sides solve the same normal equations in f64 (pivoted LDLT); only the summation order differs


def _oracle():
    from oracle import backend_oracle as bo
    if not bo.available() or not hasattr(bo.lib(), "vref_visual_imu_align"):
        pytest.skip("oracle/_ref not built")
    return bo


def _batch(cases, F, M):
    B = len(cases)
    n = np.array([c["n"] for c in cases], np.int32)
    R = np.stack([c["R"] for c in cases]); T = np.stack([c["T"] for c in cases])
    counts = np.stack([c["counts"] for c in cases]); imu0 = np.stack([c["imu0"] for c in cases]); imu = np.stack([c["imu"] for c in cases])
    assert R.shape == (B, F, 3, 3) and imu.shape == (B, F, M, 7)
    return n, R, T, counts, imu0, imu


@pytest.mark.gpu
def test_visual_imu_align_matches_reference():
    bo = _oracle()
    cfg = abi.default_config()
    F, M = 36, 24
    spec = [(0, 11), (1, 20), (2, 35), (3, 5), (4, 17)]
    cases = [be_common.align_case(sid, n, max_frames=F, max_imu=M) for sid, n in spec]
    # stream 5: the SfM came out mirrored (negative scale) -> SolveScale rejects, VisualIMUAlignment returns false
    bad = be_common.align_case(5, 14, max_frames=F, max_imu=M)
    bad["T"] = -bad["T"]
    cases.append(bad)
    bg0 = np.zeros((len(cases), 3)); bg0[1] = [0.002, -0.001, 0.0005]
    bgs, g, x, ok = api.visual_imu_align(cfg, *_batch(cases, F, M), bg0)
    worst = 0.0
    for b, c in enumerate(cases):
        n = c["n"]
        rb, rg, rx, rok = bo.visual_imu_align(n, c["R"][:n], c["T"][:n], c["counts"][:n], c["imu0"][:n], c["imu"][:n], bg0[b], c["tic"])
        assert int(ok[b]) == rok, (b, ok[b], rok)
        e = max(be_common.rel_err(bgs[b], rb), be_common.rel_err(g[b], rg), be_common.rel_err(x[b, :3 * n + 3], rx))
        worst = max(worst, e)
        print(f"[align] stream {b} n={n} ok={rok} rel err {e:.3e}")
        assert e < TOL, (b, n, e)
    assert ok[:5].all() and ok[5] == 0
    # and the answer is the physical one: scale, gravity direction and gyroscope bias of the synthetic streams (noise-limited)
    for b, c in enumerate(cases[:3]):
        n = c["n"]
        assert abs(x[b, 3 * n + 2] / c["scale"] - 1) < 0.25
        assert np.degrees(np.arccos(np.clip(g[b] @ c["g"] / 9.805 ** 2, -1, 1))) < 6.0
        assert np.abs(bgs[b] - c["gyro_bias"]).max() < 3e-3
    print(f"\n[align] worst relative |gpu - reference| over {len(cases)} streams: {worst:.3e}")


@pytest.mark.gpu
def test_visual_imu_align_matches_golden():
    """Against the committed outputs of the reference (tests/golden/align_golden.npz, made by tests/golden/make_align_golden.py)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.npz"))
    cfg = abi.default_config()
    F = max(z["R0"].shape[0], z["R1"].shape[0]); M = z["imu0"].shape[1]
    pad = lambda a, i: np.concatenate([a, np.zeros((F - a.shape[0],) + a.shape[1:], a.dtype)]) if a.shape[0] < F else a
    arrs = [np.stack([pad(z[f"{k}{i}"], i) for i in range(2)]) for k in ("R", "T", "counts", "imu0", "imu")]
    n = np.array([z["R0"].shape[0], z["R1"].shape[0]], np.int32)
    bgs, g, x, ok = api.visual_imu_align(cfg, n, *arrs, np.stack([z["bg00"], z["bg01"]]))
    for i in range(2):
        assert int(ok[i]) == int(z[f"ok{i}"])
        e = max(be_common.rel_err(bgs[i], z[f"bgs{i}"]), be_common.rel_err(g[i], z[f"g{i}"]), be_common.rel_err(x[i, :3 * n[i] + 3], z[f"x{i}"]))
        assert e < TOL, (i, e)


@pytest.mark.gpu
def test_visual_imu_align_capacity_and_arguments():
    cfg = abi.default_config()
    F, M = 12, 24
    c = be_common.align_case(0, 11, max_frames=F, max_imu=M)
    n, R, T, counts, imu0, imu = _batch([c], F, M)
    with pytest.raises(api.VioError) as e:
        api.visual_imu_align(cfg, np.array([F + 1], np.int32), R, T, counts, imu0, imu, np.zeros((1, 3)))
    assert "VIO_ERR_CAPACITY" in str(e.value)
    counts2 = counts.copy(); counts2[0, 3] = M + 1
    with pytest.raises(api.VioError) as e:
        api.visual_imu_align(cfg, n, R, T, counts2, imu0, imu, np.zeros((1, 3)))
    assert "VIO_ERR_CAPACITY" in str(e.value)
