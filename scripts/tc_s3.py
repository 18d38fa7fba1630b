"""Dev: accuracy/time of 3 vs 4 digit planes at C3 against the fp64 path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200
N, M = 500, 100000
res = {}
for tag, prec, S in (("fp64", 0, None), ("S4", 1, "4"), ("S3", 1, "3"), ("S2", 1, "2")):
    if S: os.environ["BALM_TC_SLICES"] = S
    c = balm_b200.Context(N, 0, prec)
    gt, init = c.synth_virtual(M, seed=10)
    c.evaluate(init, want_H=False)
    c.reset_counters()
    H, g, r = c.evaluate(init)
    tm = c.timings()
    dx, q1, bad = c.solve(0.01)
    # second point: near convergence
    poses, tr, _ = c.damping_iter(init, max_iter=3, gauge_mode=2, min_planes_per_pose=0)
    H2, g2, r2 = c.evaluate(poses)
    dx2, _, _ = c.solve(tr[-1]["u"])
    res[tag] = (H, dx, H2, dx2, poses)
    print(tag, "syrk ms", round(tm["ms_syrk"], 3), "obs ms", round(tm["ms_obs"], 3), flush=True)
    c.close()
H0, dx0, H20, dx20, p0 = res["fp64"]
for tag in ("S4", "S3", "S2"):
    H, dx, H2, dx2, p = res[tag]
    print(tag, "relH %.2e |ddx| %.2e (|dx| %.2e) ; near-conv relH %.2e |ddx| %.2e (|dx| %.2e); dpose after 3 it %.2e" % (
        np.abs(H - H0).max() / np.abs(H0).max(), np.abs(dx - dx0).max(), np.abs(dx0).max(),
        np.abs(H2 - H20).max() / np.abs(H20).max(), np.abs(dx2 - dx20).max(), np.abs(dx20).max(), np.abs(p - p0).max()))
