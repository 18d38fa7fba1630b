"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): both precision modes, dense + sparse; the persistent
tile-DAG factorisation with several 64-column steps (n = 240: chain + near/far workers), the multi-kernel fallback, the
sliding-window marginalisation, the pose-covariance propagation, a multi-batch sparse evaluation, and the device
association chain balm_cut_voxels -> balm_marginalize -> balm_append_scan (octree keys + layers)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import balm_b200, scenes
for prec in (0, 1):
    for drop in (0.0, 0.4):
        sc = scenes.make_scene(n_poses=40, n_planes=90, seed=3, drop=drop, pts_size=6, with_fix=True)
        c = balm_b200.Context(40, 0, prec)
        c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
        H, g, r = c.evaluate(sc["poses_init"])
        poses, tr, _ = c.damping_iter(sc["poses_init"], max_iter=2, min_planes_per_pose=0, hess_includes_fix=True)
        print("prec", prec, "drop", drop, "r", r, "iters", len(tr), flush=True)
        if prec == 0:
            raw, cov = c.pose_covariance(poses, point_noise=0.01, include_fix=True)
            print("  covariance trace", np.trace(cov), flush=True)
            M2, K2 = c.marginalize(3, poses, min_ps=5)
            shifted = np.vstack([poses[3:], np.tile(poses[-1], (3, 1))])
            print("  marginalised ->", M2, K2, "residual", c.residual(shifted), flush=True)
        c.close()
os.environ["BALM_G_BUDGET_MB"] = "1"
sc = scenes.make_scene(n_poses=20, n_planes=120, seed=4, drop=0.5, pts_size=6)
c = balm_b200.Context(20, 0, 1)
c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
print("batched sparse r", c.evaluate(sc["poses_init"])[2], flush=True)
c.close()
import assoc_ref
N, mg = 6, 2
pts, frs, poses = assoc_ref.synthetic_scans(n_poses=N + 1, pts_per_scan=1500, seed=13)
p12 = scenes.pack_poses([r for r, _ in poses], [p for _, p in poses])
c = balm_b200.Context(N, 0, 0)
win = frs < N
M0, K0 = c.cut_voxels(pts[win].astype(np.float32), frs[win], p12[:N], voxel_size=2.0, layer_limit=2, min_ps=15,
                      eigen_value_array=(1 / 16, 1 / 16, 1 / 16))
M1, K1 = c.marginalize(mg, p12[:N], min_ps=15)
shifted = np.vstack([p12[mg:N], p12[N:N + 1], p12[N:N + 1]])
M2, K2, matched = c.append_scan(pts[frs == N].astype(np.float32), shifted, N - mg)
keys, layers = c.download_keys(with_layers=True)
print("assoc chain", (M0, K0), (M1, K1), (M2, K2), "matched", matched, "layers", np.bincount(layers, minlength=3), flush=True)
c.close()
print("SANITIZE_DONE")
