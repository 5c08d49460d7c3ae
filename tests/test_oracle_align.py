"""CPU: the compiled reference VisualIMUAlignment (oracle/_ref) reproduces the committed golden vectors and recovers the truth of the
synthetic streams (scale, gravity, gyroscope bias) -- the pin of the oracle the GPU parity test compares against."""
import os

import numpy as np
import pytest

import be_common

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.npz")


def _oracle():
    from oracle import backend_oracle as bo
    if not bo.available() or not hasattr(bo.lib(), "vref_visual_imu_align"):
        pytest.skip("oracle/_ref not built")
    return bo


def test_reference_alignment_reproduces_golden():
    bo = _oracle()
    z = np.load(GOLD)
    for i in range(2):
        n = z[f"R{i}"].shape[0]
        bgs, g, x, ok = bo.visual_imu_align(n, z[f"R{i}"], z[f"T{i}"], z[f"counts{i}"], z[f"imu0{i}"], z[f"imu{i}"], z[f"bg0{i}"], z[f"tic{i}"])
        assert ok == int(z[f"ok{i}"])
        # same binary, same inputs: bitwise on the build machine; 1e-12 leaves room for another libm / compiler on the GPU box
        assert be_common.rel_err(bgs, z[f"bgs{i}"]) < 1e-12 and be_common.rel_err(g, z[f"g{i}"]) < 1e-12 and be_common.rel_err(x, z[f"x{i}"]) < 1e-12


def test_reference_alignment_recovers_synthetic_truth():
    bo = _oracle()
    for sid, n in ((0, 11), (1, 20), (2, 35)):
        c = be_common.align_case(sid, n)
        bgs, g, x, ok = bo.visual_imu_align(n, c["R"], c["T"], c["counts"], c["imu0"], c["imu"], np.zeros(3), c["tic"])
        assert ok == 1
        assert abs(x[3 * n + 2] / c["scale"] - 1) < 0.25
        assert np.degrees(np.arccos(np.clip(g @ c["g"] / 9.805 ** 2, -1, 1))) < 6.0
        assert abs(np.linalg.norm(g) - 9.805) < 1e-9
        assert np.abs(bgs - c["gyro_bias"]).max() < 3e-3


INIT_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_sfm_golden.npz")


@pytest.mark.parametrize("name", ["rejected_then_accepted", "non_keyframe"])
def test_reference_initialisation_from_sfm_reproduces_golden_and_is_metric(name):
    """The oracle's visualInitialAlign (restated around the reference's own VisualIMUAlignment) + first solve: same window as the committed
    golden vectors, and the physical one -- metric displacement over the window against the synthetic ground truth, gravity on +z."""
    import importlib
    bo = _oracle()
    if not hasattr(bo.lib(), "vref_set_init_sfm_frames"):
        pytest.skip("oracle/_ref predates the initialisation entry points")
    abi = importlib.import_module("vins-mobile_b200.abi")
    cfg = abi.default_config(batch=1, max_cnt=150)
    z = np.load(INIT_GOLD)
    ref = bo.RefEstimator(cfg)
    try:
        with be_common.Quiet():
            tr = be_common.init_scenario(ref, cfg, name)
        st = ref.state()
        ok, g, sc = ref.init_result()
        assert ok == 1 and ref.info()["solver_flag"] == 1
        for k in ("P", "Q", "V"):
            assert be_common.rel_err(st[k], z[f"{name}_{k}"]) < 1e-9, k            # same binary: bitwise here; room for another libm elsewhere
        assert abs(sc / float(z[f"{name}_scale"]) - 1) < 1e-12
        W = cfg.window_size
        kf = [int(np.argmin(np.abs(tr["t_kf"] - h))) for h in st["headers"]]
        d_est = np.linalg.norm(st["P"][W] - st["P"][0]); d_true = np.linalg.norm(tr["P"][kf[W]] - tr["P"][kf[0]])
        assert abs(d_est / d_true - 1) < 0.05, (d_est, d_true)
        assert abs(sc / 2.5 - 1) < 0.05 and np.abs(g - [0, 0, 9.805]).max() < 1e-9
    finally:
        ref.close()
