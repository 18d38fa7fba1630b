"""Dev check of the tcgen05 path against the fp64 path / oracle (run on the GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import balm_b200, scenes

def rel(a, b): return np.abs(a - b).max() / np.abs(b).max()

for (N, M, drop) in [(6, 40, 0.0), (50, 300, 0.0), (37, 64, 0.0), (20, 200, 0.4), (200, 2000, 0.0)]:
    sc = scenes.make_scene(n_poses=N, n_planes=M, seed=21, drop=drop, pts_size=10)
    res = {}
    for mode in (0, 1):
        c = balm_b200.Context(N, 0, mode)
        c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
        t = time.time()
        H, g, r = c.evaluate(sc["poses_init"])
        res[mode] = (H, g, r, time.time() - t)
        c.close()
    H0, g0, r0, _ = res[0]; H1, g1, r1, _ = res[1]
    dscale = np.sqrt(np.abs(np.diag(H0)))
    print(f"N={N} M={M} drop={drop}: relH {rel(H1, H0):.3e} rel-to-diag {np.abs((H1-H0)/np.outer(dscale,dscale)).max():.3e} "
          f"relg {rel(g1, g0):.3e} r {abs(r1-r0)/abs(r0):.1e} sym {np.array_equal(H1, H1.T)}", flush=True)

# C3-size: tensor vs fp64 on device
N, M = 500, 100000
out = {}
for mode in (0, 1):
    c = balm_b200.Context(N, 0, mode)
    gt, init = c.synth_virtual(M, seed=10)
    c.evaluate(init, want_H=False)
    c.reset_counters()
    H, g, r = c.evaluate(init)
    tm = c.timings()
    out[mode] = (H, g, r)
    print("mode", mode, {k: round(v, 3) for k, v in tm.items() if v}, flush=True)
    if mode == 1:
        dx1, q1, bad = c.solve(0.01)
    else:
        dx0, q0, bad = c.solve(0.01)
    c.close()
H0, g0, r0 = out[0]; H1, g1, r1 = out[1]
print(f"C3: relH {rel(H1, H0):.3e} relg {rel(g1, g0):.3e} |dx1-dx0| {np.abs(dx1-dx0).max():.3e} |dx| {np.abs(dx0).max():.3e}")
