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
