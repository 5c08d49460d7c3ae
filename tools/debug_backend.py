"""Print per-keyframe differences between the reference estimator and the CUDA back end (run on the GPU box)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import backend_oracle as bo
from be_common import Quiet, drive, quat_err, rel_err
synth = importlib.import_module("vins-mobile_b200.synth"); abi = importlib.import_module("vins-mobile_b200.abi"); api = importlib.import_module("vins-mobile_b200.api")

n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = abi.default_config(batch=1, max_cnt=150)
W = cfg.window_size
tr = synth.make_tracks(sid, n_kf, max_cnt=150)
ref = bo.RefEstimator(cfg); gpu = api.BackEnd(cfg)
np.set_printoptions(precision=5, linewidth=200, suppress=False)
for k in range(n_kf):
    with Quiet():
        drive(ref, tr, k, W)
    try:
        drive(gpu, tr, k, W)
        gs, gi, gf = gpu.state(), gpu.info(), gpu.features()
    except Exception as e:
        print("kf", k, "GPU EXCEPTION", e); break
    rs, ri, rf = ref.state(), ref.info(), ref.features()
    line = f"kf {k:2d} ref[m{ri['marg_flag']} nf{ri['n_feat']} np{ri['n_proj']} it{ri['iters']} c0 {ri['cost0']:.6f} c1 {ri['cost1']:.6f} pn{ri['prior_n']} F{ri['failure']} fc{ri['frame_count']} ltn{ri['last_track_num']}] "
    line += f"gpu[m{gi['marg_flag']} nf{gi['n_feat']} np{gi['n_proj']} it{gi['iters']} c0 {gi['cost0']:.6f} c1 {gi['cost1']:.6f} pn{gi['prior_n']} F{gi['failure']} fc{gi['frame_count']} ltn{gi['last_track_num']} err{gi['err']}] "
    line += "dP %.2e dQ %.2e dV %.2e dBa %.2e dBg %.2e" % (rel_err(gs['P'], rs['P']), quat_err(gs['Q'], rs['Q']), rel_err(gs['V'], rs['V']), rel_err(gs['Ba'], rs['Ba']) if np.abs(rs['Ba']).max() > 0 else 0, rel_err(gs['Bg'], rs['Bg']) if np.abs(rs['Bg']).max() > 0 else 0)
    same_feat = len(rf['ids']) == len(gf['ids']) and np.array_equal(rf['ids'], gf['ids']) and np.array_equal(rf['start'], gf['start']) and np.array_equal(rf['n_obs'], gf['n_obs'])
    line += f" feat[{len(rf['ids'])},{len(gf['ids'])} same={same_feat}]"
    if same_feat and k >= W:
        m = (rf['depth'] > 0) & (gf['depth'] > 0)
        line += " ddepth %.2e flag_same=%s" % (np.abs(gf['depth'][m] / rf['depth'][m] - 1).max() if m.any() else 0, np.array_equal(rf['solve_flag'], gf['solve_flag']))
    if k >= W:
        rps, gps = ref.post_solve(), gpu.post_solve()
        line += " post dP %.2e dV %.2e" % (rel_err(gps[:, :3], rps[:, :3]), rel_err(gps[:, 7:10], rps[:, 7:10]))
        rp, gp = ref.prior(), gpu.prior()
        if rp is not None and gp is not None:
            line += " prior dH %.2e db %.2e c0 %.4e/%.4e pres_same=%s" % (rel_err(gp['H'], rp['H']), rel_err(gp['b'], rp['b']), gp['c0'], rp['c0'], np.array_equal(rp['present'], gp['present']))
        else:
            line += f" prior ref={rp is not None} gpu={gp is not None}"
    print(line, flush=True)
