"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): both precision modes, dense + sparse."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import balm_b200, scenes
for prec in (0, 1):
    for drop in (0.0, 0.4):
        sc = scenes.make_scene(n_poses=40, n_planes=90, seed=3, drop=drop, pts_size=6)
        c = balm_b200.Context(40, 0, prec)
        c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
        H, g, r = c.evaluate(sc["poses_init"])
        poses, tr, _ = c.damping_iter(sc["poses_init"], max_iter=2, min_planes_per_pose=0)
        print("prec", prec, "drop", drop, "r", r, "iters", len(tr), flush=True)
        c.close()
print("SANITIZE_DONE")
