"""Dev: time the GPU association on a realworld-sized input (177 scans x 75k points = 13.3M points)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import balm_b200, assoc_ref, scenes
n_poses, pps = int(os.environ.get("NP", 177)), int(os.environ.get("PPS", 75000))
pts, frs, poses = assoc_ref.synthetic_scans(n_poses=n_poses, pts_per_scan=pps, seed=1, room=12.0)
poses12 = scenes.pack_poses([r for r, _ in poses], [p for _, p in poses])
xyz = pts.astype(np.float32)
c = balm_b200.Context(n_poses, 0, 1)
for rep in range(3):
    c.sync(); t0 = time.perf_counter()
    M, K = c.cut_voxels(xyz, frs, poses12)
    c.sync(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {len(frs)} points -> {M} plane voxels, {K} observations in {dt*1e3:.1f} ms ({len(frs)/dt/1e6:.1f} Mpoints/s incl. H2D of {xyz.nbytes/1e6:.0f} MB)")
t0 = time.perf_counter()
poses_out, tr, _ = c.damping_iter(poses12, min_planes_per_pose=0)
print(f"BA on them: {len(tr)} LM iterations in {(time.perf_counter()-t0)*1e3:.1f} ms, cost {tr[0]['r1']:.4f} -> {tr[-1]['r2']:.4f}")
if os.environ.get("CHECK"):
    t0 = time.perf_counter()
    ref = assoc_ref.cut_voxels(pts, frs, poses)
    print(f"numpy restatement: {time.perf_counter()-t0:.1f} s, voxels {len(ref[3])}")
