import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("vins-mobile_b200")


@pytest.fixture(scope="session")
def abi():
    return importlib.import_module("vins-mobile_b200.abi")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("vins-mobile_b200.synth")


@pytest.fixture(scope="session")
def api():
    return importlib.import_module("vins-mobile_b200.api")


_STREAMS = {}


@pytest.fixture(scope="session")
def get_stream(synth):
    """Cached synthetic streams: get_stream(stream_id, n_frames) -> synth.Stream (CPU render)."""
    def _get(sid, n):
        key = (sid, n)
        if key not in _STREAMS:
            _STREAMS[key] = synth.make_stream(sid, n)
        return _STREAMS[key]
    return _get


def texture_pair(seed=5, rows=640, cols=480):
    """Two views of a random smooth texture related by a small similarity (the survey's LK test pair)."""
    import cv2
    r = np.random.default_rng(seed)
    small = r.integers(0, 256, (rows // 8 + 4, cols // 8 + 4)).astype(np.float32)
    img = cv2.resize(small, ((cols // 8 + 4) * 8, (rows // 8 + 4) * 8), interpolation=cv2.INTER_CUBIC)
    img = cv2.GaussianBlur(img, (0, 0), 1.5)
    big = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    img0 = big[16:16 + rows, 16:16 + cols].copy()
    M = cv2.getRotationMatrix2D((cols / 2 + 16, rows / 2 + 16), 1.5, 1.01)
    M[0, 2] += 3.3
    M[1, 2] -= 2.2
    img1 = cv2.warpAffine(big, M, (big.shape[1], big.shape[0]), flags=cv2.INTER_LINEAR)[16:16 + rows, 16:16 + cols].copy()
    return img0, img1


def two_view_points(seed, n=150, nout=15, noise=0.05):
    """Random two-view correspondences with outliers for the RANSAC-F tests."""
    import cv2
    r = np.random.default_rng(seed)
    K = np.array([[526.6, 0, 243.5], [0, 526.7, 315.3], [0, 0, 1]])
    X = np.stack([r.uniform(-2, 2, n), r.uniform(-3, 3, n), r.uniform(2, 6, n)], 1)
    R, _ = cv2.Rodrigues(r.normal(0, 0.02, 3))
    t = r.normal(0, 0.05, 3)
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    X2 = (R @ X.T).T + t
    x2 = (K @ X2.T).T
    x2 = x2[:, :2] / x2[:, 2:]
    x1 = x1 + r.normal(0, noise, x1.shape)
    x2 = x2 + r.normal(0, noise, x2.shape)
    o = r.choice(n, nout, replace=False)
    x2[o] += r.uniform(-8, 8, (nout, 2))
    return x1.astype(np.float32), x2.astype(np.float32)
