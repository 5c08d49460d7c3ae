"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, refuses to compute without a device, the oracle
reproduces the committed golden vectors, and the N>1 host logic (stream sharding + state gather) works over gloo world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(api):
    lib = api.lib()
    names = api.exported_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_struct_matches_ctypes(abi):
    """sizeof(vio_config) as gcc lays it out == the ctypes mirror."""
    src = '#include "vio_b200.h"\n#include <stdio.h>\nint main(){printf("%zu\\n", sizeof(vio_config));return 0;}\n'
    exe = "/tmp/_vio_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    n = int(subprocess.run([exe], capture_output=True, check=True).stdout)
    assert n == C.sizeof(abi.VioConfig)


def test_config_default_matches_reference_constants(api, abi):
    c = abi.VioConfig()
    api.lib().vio_config_default(C.byref(c))
    # global_param.cpp:29-39 (iPhone7P), global_param.hpp:28,37,42-46, feature_tracker.hpp:25-28, feature_manager.hpp:24-25
    assert (c.rows, c.cols) == (640, 480)
    assert (c.fx, c.fy, c.cx, c.cy) == (526.600, 526.678, 243.481, 315.280)
    assert list(c.tic) == [0.0, 0.092, 0.01]
    assert (c.min_dist, c.f_threshold, c.freq, c.window_size, c.num_of_f) == (30, 1.0, 3, 10, 1000)
    assert (c.acc_n, c.acc_w, c.gyr_n, c.gyr_w, c.gravity) == (0.5, 0.002, 0.2, 4.0e-5, 9.805)
    assert c.max_iters == 10 and c.min_parallax == 10.0 / 549.0 and c.init_depth == 5.0
    d = abi.default_config()
    for f, _ in abi.VioConfig._fields_:
        a, b = getattr(c, f), getattr(d, f)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), f


def test_no_cpu_fallback(api, abi):
    """Without a CUDA device every compute entry point must fail loudly (never fall back to the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.VioError):
        api.FrontEnd(abi.default_config())
    with pytest.raises(api.VioError):
        api.BackEnd(abi.default_config())
    with pytest.raises(api.VioError):
        api.prim_pyramid(abi.default_config(), np.zeros((640, 480), np.uint8))
    with pytest.raises(api.VioError):
        api.PnP(abi.default_config())
    with pytest.raises(api.VioError):
        api.prim_clahe(abi.default_config(), np.zeros((640, 480), np.uint8))


def test_product_path_does_not_touch_oracle():
    for f in ("api.py", "abi.py", "shard.py", "__init__.py"):
        txt = open(os.path.join(ROOT, "vins-mobile_b200", f)).read()
        assert not re.search(r"^\s*(import|from)\s+.*(oracle|cv2)", txt, re.M), f
    for f in os.listdir(os.path.join(ROOT, "vins-mobile_b200", "csrc")):
        txt = open(os.path.join(ROOT, "vins-mobile_b200", "csrc", f)).read()
        assert "#include \"../../oracle" not in txt and "libvins_ref" not in txt


def test_argument_validation(api, abi):
    h = C.c_void_p()
    bad = abi.default_config()
    bad.max_cnt = 100000
    assert api.lib().vio_frontend_create(C.byref(bad), C.byref(h)) == 1          # VIO_ERR_ARG
    bad = abi.default_config()
    bad.window_size = 100
    assert api.lib().vio_backend_create(C.byref(bad), C.byref(h)) == 1
    bad = abi.default_config()
    bad.max_cnt = 100000
    assert api.lib().vio_pnp_create(C.byref(bad), C.byref(h)) == 1
    assert api.lib().vio_frontend_set_clahe(None, 1, 3.0, 8, 8) == 1
    assert api.lib().vio_pnp_process_imu(None, 1, None, None, None) == 1


# ------------------------------------------------------------------------------- golden vectors vs the oracle
def test_frontend_oracle_reproduces_golden():
    import frontend_oracle as fo
    g = np.load(os.path.join(GOLD, "frontend_golden.npz"))
    pyr = fo.r_build_pyramid(g["img0"])
    for a, k in zip(pyr[1:], ("l1", "l2", "l3")):
        assert np.array_equal(a, g[k])
    assert np.array_equal(fo.r_min_eig_map(g["img0"]), g["eig"])
    mask = np.full(g["img0"].shape, 255, np.uint8)
    yy, xx = np.mgrid[0:mask.shape[0], 0:mask.shape[1]]
    for c in g["kept"]:
        mask[(xx - int(np.rint(c[0]))) ** 2 + (yy - int(np.rint(c[1]))) ** 2 <= 900] = 0
    assert np.array_equal(fo.r_good_features(g["img0"], mask, 20), g["corners"])
    nxt, st = fo.r_lk_track(fo.r_build_pyramid(g["img0"]), fo.r_build_pyramid(g["img1"]), g["lk_pts"])
    assert np.array_equal(st, g["lk_status"])
    assert np.array_equal(nxt[st == 1].view(np.uint32), g["lk_next"][st == 1].view(np.uint32))      # bitwise the cv2 output
    assert np.array_equal(fo.r_find_fundamental(g["f_x1"], g["f_x2"]), g["f_mask"])


def test_clahe_oracle_reproduces_golden():
    import frontend_oracle as fo
    g = np.load(os.path.join(GOLD, "clahe_golden.npz"))
    assert np.array_equal(fo.r_clahe(g["img"], 3.0, (8, 8)), g["out"])


def test_backend_oracle_reproduces_golden(abi, synth):
    import backend_oracle as bo
    if not bo.available():
        pytest.skip("oracle/_ref not built")
    from be_common import Quiet, drive
    g = np.load(os.path.join(GOLD, "backend_golden.npz"))
    pqv, jac, cov, sdt = bo.preintegrate(g["dt"], g["acc"], g["gyr"], g["acc"][0], g["gyr"][0], g["ba"], g["bg"])
    assert np.array_equal(pqv, g["pqv"]) and np.array_equal(jac, g["jac"]) and np.array_equal(cov, g["cov"])
    r, J = bo.imu_factor(pqv, jac, cov, sdt, g["ba"], g["bg"], g["pi"], g["sbi"], g["pj"], g["sbj"])
    assert np.allclose(r, g["imu_r"], rtol=1e-12, atol=0) and np.allclose(J, g["imu_J"], rtol=1e-12, atol=1e-300)
    cfg = abi.default_config(batch=1, max_cnt=int(g["max_cnt"]))
    tr = synth.make_tracks(int(g["track_seed"]), int(g["n_kf"]), max_cnt=int(g["max_cnt"]))
    est = bo.RefEstimator(cfg)
    for k in range(int(g["n_kf"])):
        with Quiet():
            drive(est, tr, k, cfg.window_size)
        s = est.state()
        got = np.concatenate([s["P"], s["Q"], s["V"], s["Ba"], s["Bg"]], 1)
        # The reference is NOT bit-reproducible across processes once marginalisation has run: MarginalizationInfo orders its blocks by
        # unordered_map iteration over heap addresses (quirk Q10), so round-off differs run to run and is amplified by the
        # not-yet-converged 10-iteration solves (measured here: 5e-9 after the first prior, 3e-6 after five).  Exact before that.
        if k <= cfg.window_size:
            assert np.array_equal(got, g["states"][k]), f"keyframe {k}"
        else:
            assert np.allclose(got, g["states"][k], rtol=0, atol=5e-5), f"keyframe {k}"
        i = est.info()
        assert [i["marg_flag"], i["n_feat"], i["n_proj"], i["prior_n"]] == g["infos"][k][[0, 1, 2, 4]].tolist()


def test_pnp_oracle_reproduces_golden_and_tracks_ground_truth(abi, synth):
    """SURVEY section 8(f) rank 3 (motion-only PnP tracker): the reference's vins_pnp.cpp / perspective_factor.cpp / imu_factor_pnp.h,
    compiled unmodified into oracle/_ref, reproduce the committed golden vectors (the solve has no wall-time cap in the oracle, so it is
    deterministic) and follow the synthetic ground truth to better than 1 cm once the 7-frame window is full."""
    import backend_oracle as bo
    if not bo.available() or not hasattr(bo.lib(), "vpnp_create"):
        pytest.skip("oracle/_ref not built")
    from be_common import Quiet, drive_pnp
    g = np.load(os.path.join(GOLD, "pnp_golden.npz"))
    r, Jp, Je = bo.perspective_factor(g["pf_obs"], g["pf_pos"], int(g["pf_track"]), abi.default_config().fx, g["pf_pose"], g["pf_ex"])
    assert np.allclose(r, g["pf_r"], rtol=1e-13, atol=0) and np.allclose(Jp, g["pf_Jp"], rtol=1e-13, atol=1e-300) and np.allclose(Je, g["pf_Je"], rtol=1e-13, atol=1e-300)
    # residual = sqrt_info * (proj - obs) * track_num / 10 and its pose Jacobian against central differences (tangent-space perturbation)
    def res_at(dp):
        pose = g["pf_pose"].copy()
        pose[:3] += dp[:3]
        x, y, z, w = pose[3:]
        dq = np.array([0.5 * dp[3], 0.5 * dp[4], 0.5 * dp[5], 1.0])
        q = np.array([w * dq[0] + x * dq[3] + y * dq[2] - z * dq[1], w * dq[1] - x * dq[2] + y * dq[3] + z * dq[0],
                      w * dq[2] + x * dq[1] - y * dq[0] + z * dq[3], w * dq[3] - x * dq[0] - y * dq[1] - z * dq[2]])
        pose[3:] = q / np.linalg.norm(q)
        return bo.perspective_factor(g["pf_obs"], g["pf_pos"], int(g["pf_track"]), abi.default_config().fx, pose, g["pf_ex"])[0]
    num = np.stack([(res_at(np.eye(6)[i] * 1e-6) - res_at(-np.eye(6)[i] * 1e-6)) / 2e-6 for i in range(6)], 1)
    assert np.allclose(num, Jp, rtol=1e-5, atol=1e-4)
    cfg = abi.default_config(batch=1, max_cnt=150)
    seq = synth.make_pnp_sequence(int(g["seed"]), int(g["n_frames"]))
    with Quiet():
        h = bo.RefPnP(cfg)
    last_t = 0.0
    for k in range(int(g["n_frames"])):
        with Quiet():
            last_t = drive_pnp(h, seq, k, last_t)
            s = h.state()
        got = np.concatenate([s["P"], s["R"].reshape(-1, 9), s["V"], s["headers"][:, None], s["find_solved"][:, None].astype(float)], 1)
        assert np.allclose(got, g["states"][k], rtol=0, atol=1e-9), f"frame {k}"
        if k >= 7:
            i = h.n - 2                                          # FeatureTracker reads vins_pnp.Ps[PNP_SIZE - 1] (feature_tracker.cpp:156)
            j = int(round(s["headers"][i] * 30.0))
            assert np.linalg.norm(s["P"][i] - seq["P"][j]) < 0.01, f"frame {k}"
            assert np.abs(s["R"][i] - seq["R"][j]).max() < 5e-3
    h.close()


# ------------------------------------------------------------------------------- N > 1 host logic over gloo
_WORKER = r'''
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
shard = importlib.import_module("vins-mobile_b200.shard")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, B, NF = dist.get_rank(), 3, 11
ids = shard.stream_ids_for_rank(rank, 2, B)
local = torch.stack([torch.full((NF, 16), float(i), dtype=torch.float64) + torch.arange(16, dtype=torch.float64) / 100 for i in ids])
allst = shard.gather_states(local)
assert allst.shape == (2 * B, NF, 16)
for sid in range(2 * B):
    assert shard.rank_of_stream(sid, B) == sid // B
    assert torch.equal(allst[sid], torch.full((NF, 16), float(sid), dtype=torch.float64) + torch.arange(16, dtype=torch.float64) / 100)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_sharding_and_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    """vins-mobile_b200/host/vio_host.hpp (FeatureTracker / VINS with the reference's member names) builds with g++ against the
    C-ABI; without a GPU the example loop must fail loudly with VIO_ERR_CUDA (exit code 2), never compute on the CPU."""
    exe = str(tmp_path / "example_loop")
    pkg = os.path.join(ROOT, "vins-mobile_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(pkg, "host", "example_loop.cpp"), "-L", pkg, "-lvio_b200",
                    f"-Wl,-rpath,{pkg}"], check=True)
    import torch
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "frame_count" in r.stdout
    else:
        assert r.returncode == 2 and "error" in r.stdout
