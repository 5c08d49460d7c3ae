"""GPU parity tests for the front end, through the C-ABI (libvio_b200.so):
CUDA kernels vs the restated oracle (bit-exact) and vs cv2 itself (exact where cv2 is exact)."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import frontend_oracle as fo
from conftest import texture_pair, two_view_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg(abi):
    return abi.default_config(batch=1, max_cnt=200)


@pytest.fixture(scope="module")
def pair():
    return texture_pair()


def test_pyramid_bit_exact(api, cfg, pair):
    img0, _ = pair
    outs = api.prim_pyramid(cfg, img0)
    ref = fo.r_build_pyramid(img0)[1:]
    for a, b in zip(outs, ref):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0], cv2.pyrDown(img0))


def test_pyramid_odd_sizes(api, abi):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (101, 77)).astype(np.uint8)
    c = abi.default_config(rows=101, cols=77, max_cnt=50)
    outs = api.prim_pyramid(c, img)
    ref = fo.r_build_pyramid(img)[1:]
    for a, b in zip(outs, ref):
        assert np.array_equal(a, b)


def test_lk_bit_exact_vs_oracle_and_status_vs_cv2(api, cfg, pair):
    img0, img1 = pair
    pts = fo.cv2_good_features(img0, None, 150)
    border = np.array([[2.5, 3.5], [477.2, 5.1], [1.0, 638.0], [478.9, 638.9], [240.3, 0.4], [0.2, 320.7], [479.4, 300.0], [100.5, 639.3]],
                      np.float32)
    pts = np.concatenate([pts, border])
    n_g, s_g = api.prim_lk(cfg, img0, img1, pts)
    n_r, s_r = fo.r_lk_track(fo.r_build_pyramid(img0), fo.r_build_pyramid(img1), pts)
    assert np.array_equal(s_g, s_r)
    ok = s_r == 1
    assert np.array_equal(n_g[ok].view(np.uint32), n_r[ok].view(np.uint32)), "LK positions must be bit-identical to the oracle"
    n_c, s_c = fo.cv2_lk_track(img0, img1, pts)
    assert np.array_equal(s_g, s_c)
    assert np.array_equal(n_g[ok].view(np.uint32), n_c[ok].view(np.uint32)), "LK positions must be bit-identical to cv2.calcOpticalFlowPyrLK"


def test_lk_negative_fourth_bilinear_weight(api, cfg, pair):
    """The three rounded 14-bit bilinear weights can add up to 2^14 + 1, which makes the fourth one -1 (OpenCV keeps it as a signed
    short).  Regression: the J-window sampler packed the weights as unsigned 16-bit halves (one gross LK error in ~20 000 tracks)."""
    img0, _ = pair
    rng = np.random.default_rng(1)
    fr = []
    while len(fr) < 40:
        a, b = rng.uniform(0, 0.02, 2).astype(np.float32)
        w = fo._weights(np.array([a]), np.array([b]))
        if int(w[3][0]) < 0:
            fr.append((a, b))
    base = fo.cv2_good_features(img0, None, 40)
    pts = (base + np.array(fr, np.float32)[:len(base)]).astype(np.float32)
    n_g, s_g = api.prim_lk(cfg, img0, img0, pts)          # identical images: the first J window sits exactly on the template position
    n_c, s_c = fo.cv2_lk_track(img0, img0, pts)
    assert np.array_equal(s_g, s_c) and s_c.all()
    assert np.array_equal(n_g.view(np.uint32), n_c.view(np.uint32))


def test_good_features_identical(api, cfg, pair):
    img0, _ = pair
    rng = np.random.default_rng(0)
    kept = rng.uniform([5, 5], [475, 635], (40, 2)).astype(np.float32)
    mask = np.full(img0.shape, 255, np.uint8)
    for c in kept:
        cv2.circle(mask, (int(np.rint(c[0])), int(np.rint(c[1]))), 30, 0, -1)
    for k in (150, 40, 7):
        g, maxv = api.prim_good_features(cfg, img0, kept, k)
        r = fo.r_good_features(img0, mask, k)
        assert np.array_equal(g, r)
        assert np.array_equal(g, fo.cv2_good_features(img0, mask, k))
        eig = fo.r_min_eig_map(img0)
        assert np.float32(maxv) == eig[mask != 0].max()
    g, _ = api.prim_good_features(cfg, img0, np.zeros((0, 2), np.float32), 150)
    assert np.array_equal(g, fo.cv2_good_features(img0, None, 150))


@pytest.mark.parametrize("seed", range(12))
def test_ransac_mask_identical(api, cfg, seed):
    n = [150, 150, 100, 40, 15, 20][seed % 6]
    x1, x2 = two_view_points(seed, n=n, nout=max(1, n // 10))
    m, iters = api.prim_ransac_f(cfg, x1, x2)
    assert m is not None
    assert np.array_equal(m, fo.r_find_fundamental(x1, x2))
    assert np.array_equal(m, fo.cv2_find_fundamental(x1, x2))
    assert iters >= 1


def test_ransac_lmeds_branch(api, cfg):
    hits = 0
    for seed in range(100, 120):
        x1, x2 = two_view_points(seed, n=14, nout=1)
        m, _ = api.prim_ransac_f(cfg, x1, x2)
        mr = fo.r_find_fundamental(x1, x2)
        hits += (m is None and mr is None) or (m is not None and mr is not None and np.array_equal(m, mr))
    assert hits >= 18


def _run_stream(api, abi, fo_backend, s, n_frames, max_cnt=150):
    c = abi.default_config(batch=1, max_cnt=max_cnt)
    fe = api.FrontEnd(c)
    tr = fo.FeatureTrackerOracle(max_cnt=max_cnt, backend=fo_backend)
    out = []
    for k in range(n_frames):
        im = s.images[k].numpy()
        pub = fe.read_images(im[None])
        _, _, pub_o = tr.read_image(im)
        assert pub == pub_o
        g = fe.stream(0)
        out.append((g, tr.ids.copy(), tr.cur_pts.copy(), tr.track_cnt.copy(), dict(tr.image_msg), fe.stats(0), dict(tr.stats)))
    fe.close()
    return out


def test_stream_ids_bit_exact_vs_restated_oracle(api, abi, get_stream):
    """The headline front-end parity claim: per-frame tracked ids, positions (bitwise), track counts and image_msg identical to
    the restated FeatureTracker oracle over a synthetic stream."""
    s = get_stream(0, 31)
    for k, (g, ids, pts, cnt, msg, st, st_o) in enumerate(_run_stream(api, abi, "restated", s, 31)):
        assert np.array_equal(g["ids"], ids), f"frame {k}: ids differ {st} {st_o}"
        assert np.array_equal(g["pts"].view(np.uint32), pts.view(np.uint32)), f"frame {k}: points differ"
        assert np.array_equal(g["track_cnt"], cnt)
        if k % 3 == 0:
            assert {int(i) for i in g["ids"]} == set(msg.keys())
            for i, xyz in zip(g["ids"], g["norm_xyz"]):
                assert tuple(xyz) == msg[int(i)]


def test_stream_vs_cv2_tracker(api, abi, synth):
    """The headline front-end parity claim against the OpenCV binary itself (the stand-in for the reference's arithmetic, SURVEY
    section 8(c)): 8 streams x 300 frames (the stream length of SURVEY section 8(d)) through one batch-8 handle; after EVERY frame the
    tracked ids, the positions (bitwise), the track counts and -- on publishing frames -- image_msg equal those of a
    FeatureTracker restatement whose KLT / RANSAC-F / goodFeaturesToTrack calls go to cv2 (feature_tracker.cpp:162-321)."""
    nb, nf = 8, 300
    streams = [synth.make_stream(i, nf, device="cuda") for i in range(nb)]
    c = abi.default_config(batch=nb, max_cnt=150)
    fe = api.FrontEnd(c)
    trs = [fo.FeatureTrackerOracle(max_cnt=150, backend="cv2") for _ in range(nb)]
    for k in range(nf):
        ims = np.stack([s.images[k].cpu().numpy() for s in streams])
        pub = fe.read_images(ims)
        for b, tr in enumerate(trs):
            _, _, pub_o = tr.read_image(ims[b])
            assert pub == pub_o
            g = fe.stream(b)
            assert np.array_equal(g["ids"], tr.ids), f"stream {b} frame {k}: ids differ"
            assert np.array_equal(g["pts"].view(np.uint32), tr.cur_pts.view(np.uint32)), f"stream {b} frame {k}: points differ"
            assert np.array_equal(g["track_cnt"], tr.track_cnt), f"stream {b} frame {k}"
            if pub:
                assert {int(i) for i in g["ids"]} == set(tr.image_msg.keys())
                for i, xyz in zip(g["ids"], g["norm_xyz"]):
                    assert tuple(xyz) == tr.image_msg[int(i)]
    fe.close()
    assert all(len(tr.ids) > 100 for tr in trs)


def test_batch_equals_single(api, abi, get_stream):
    """Batched trackers are independent: stream b of a batch-3 handle == a batch-1 handle fed the same frames."""
    streams = [get_stream(i, 10) for i in range(3)]
    c3 = abi.default_config(batch=3, max_cnt=150)
    fe3 = api.FrontEnd(c3)
    singles = [api.FrontEnd(abi.default_config(batch=1, max_cnt=150)) for _ in range(3)]
    for k in range(10):
        imgs = np.stack([s.images[k].numpy() for s in streams])
        fe3.read_images(imgs)
        for b in range(3):
            singles[b].read_images(imgs[b:b + 1])
            a, d = fe3.stream(b), singles[b].stream(0)
            assert np.array_equal(a["ids"], d["ids"])
            assert np.array_equal(a["pts"].view(np.uint32), d["pts"].view(np.uint32))
    assert fe3.launch_count() > 0
    fe3.close()
    for f in singles:
        f.close()


def test_ui_outputs(api, abi, get_stream):
    s = get_stream(0, 7)
    c = abi.default_config(batch=1, max_cnt=150)
    fe = api.FrontEnd(c)
    tr = fo.FeatureTrackerOracle(max_cnt=150, backend="restated")
    for k in range(7):
        im = s.images[k].numpy()
        fe.read_images(im[None])
        good, tl, _ = tr.read_image(im)
        g, t = fe.ui(0)
        assert len(g) == len(good)
        if len(good):
            assert np.array_equal(g, np.array(good, np.float32))
            assert np.allclose(t, np.array(tl), rtol=0, atol=1e-15)
    fe.close()


def test_frontend_golden_vectors(api, abi):
    """CUDA primitives vs the committed cv2 outputs (tests/golden/frontend_golden.npz, 160x128 image pair)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_golden.npz"))
    img0, img1 = g["img0"], g["img1"]
    c = abi.default_config(rows=img0.shape[0], cols=img0.shape[1], max_cnt=64)
    l1, l2, l3 = api.prim_pyramid(c, img0)
    assert np.array_equal(l1, g["l1"]) and np.array_equal(l2, g["l2"]) and np.array_equal(l3, g["l3"])
    corners, _ = api.prim_good_features(c, img0, g["kept"], 20)
    assert np.array_equal(corners, g["corners"])
    nxt, st = api.prim_lk(c, img0, img1, g["lk_pts"])
    assert np.array_equal(st, g["lk_status"])
    assert np.array_equal(nxt[st == 1].view(np.uint32), g["lk_next"][st == 1].view(np.uint32))      # bitwise the cv2 output
    m, _ = api.prim_ransac_f(c, g["f_x1"], g["f_x2"])
    assert np.array_equal(m, g["f_mask"])


def test_config_c4_1280x720_300_features(api, abi, synth):
    """BASELINE.json configs[4] shape: 1280x720, 300 features.  Tens of thousands of corner candidates per frame exercise the banded
    selection; ids / points / counts must stay bit-identical to the restated oracle."""
    cam = synth.Camera().scaled(720, 1280)
    s = synth.make_stream(7, 5, cam=cam)
    c = abi.default_config(batch=1, max_cnt=300, rows=720, cols=1280)
    fe = api.FrontEnd(c)
    tr = fo.FeatureTrackerOracle(rows=720, cols=1280, max_cnt=300, fx=c.fx, fy=c.fy, cx=c.cx, cy=c.cy, backend="restated")
    for k in range(5):
        im = s.images[k].numpy()
        fe.read_images(im[None])
        tr.read_image(im)
        g = fe.stream(0)
        assert np.array_equal(g["ids"], tr.ids), f"frame {k}: {fe.stats(0)} {tr.stats}"
        assert np.array_equal(g["pts"].view(np.uint32), tr.cur_pts.view(np.uint32))
    assert len(tr.ids) >= 250
    fe.close()


def test_clahe_bit_exact_and_tracker_on_equalised_frames(api, abi, get_stream):
    """K0: the CUDA CLAHE equals the restated OpenCV CLAHE byte for byte (640x480 and 1280x720, clip 3, 8x8 tiles), and a tracker with
    vio_frontend_set_clahe() enabled publishes exactly what the oracle tracker publishes on frames equalised by the oracle."""
    from conftest import texture_pair
    c = abi.default_config(batch=1, max_cnt=150)
    s = get_stream(0, 7)
    for im in (s.images[0].numpy(), texture_pair(11)[0], np.random.default_rng(1).integers(0, 256, (640, 480)).astype(np.uint8)):
        assert np.array_equal(api.prim_clahe(c, im), fo.r_clahe(im))
    big = texture_pair(2, rows=720, cols=1280)[0]
    assert np.array_equal(api.prim_clahe(abi.default_config(batch=1, max_cnt=150, rows=720, cols=1280), big), fo.r_clahe(big))
    with pytest.raises(api.VioError):
        api.prim_clahe(c, s.images[0].numpy(), 3.0, 7, 8)                  # 480 % 7 != 0: OpenCV would pad
    fe = api.FrontEnd(c)
    fe.set_clahe(True, 3.0, 8, 8)
    tr = fo.FeatureTrackerOracle(max_cnt=150, backend="restated")
    for k in range(7):
        im = s.images[k].numpy()
        fe.read_images(im[None])
        tr.read_image(fo.r_clahe(im))
        g = fe.stream(0)
        assert np.array_equal(g["ids"], tr.ids), f"frame {k}"
        assert np.array_equal(g["pts"].view(np.uint32), tr.cur_pts.view(np.uint32)), f"frame {k}"
    fe.close()
