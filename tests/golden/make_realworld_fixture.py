#!/usr/bin/env python
"""Builds tests/golden/realworld_voxels.npz from the reference's own real-world dataset
(/root/reference/datas/benchmark_realworld: binary PCD scans + alidarPose.csv). Run in the build container only:
/root/reference does not exist on the GPU box, so the tests read the committed .npz, never the dataset.

The plane voxels are produced by a numpy restatement of the reference's association step (the code that FEEDS the
hot path):
  read_pose            src/benchmark/benchmark_realworld.cpp:31-73   (4 csv lines per pose = rows of [R|p; 0 0 0 t])
  PCD reader           :75-106 (x y z intensity ... float32, DATA binary)
  re-anchor to pose 0  :163-168
  cut_voxel            src/benchmark/bavoxel.hpp:1170-1223 (float32 voxel index, "-1 for negatives", trunc to int64)
  OCTO_TREE_NODE::recut / judge_eigen / cut_func   bavoxel.hpp:654-776  (layer_limit=2, min_ps=15, layer_size=30,
                       eigen_value_array={1/16,1/16,1/9} as set at benchmark_realworld.cpp:183-185, voxel_size=2)
  tras_opt -> VOX_HESS::push_voxel                 bavoxel.hpp:908-929, 30-51
To keep the fixture small only the first N_POSES scans are used (win_size = N_POSES) and points are decimated.
Outputs: CSR arrays of body-frame point clusters (exactly what push_voxel registers), initial poses, and -- as golden
outputs -- the CPU oracle's LM trace and refined poses on this input.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DATA = "/root/reference/datas/benchmark_realworld"
N_POSES = int(os.environ.get("N_POSES", "48"))
DECIMATE = int(os.environ.get("DECIMATE", "2"))
OUT = os.environ.get("OUT", "realworld_voxels.npz")  # N_POSES=177 DECIMATE=12 OUT=realworld_c5_177.npz: every scan of the dataset
VOXEL_SIZE = 2.0
LAYER_LIMIT = 2
MIN_PS = 15
LAYER_SIZE = [30, 30, 30, 30]
EIGEN_VALUE_ARRAY = [1.0 / 16, 1.0 / 16, 1.0 / 9, 1.0 / 16]


def read_poses(path):
    rows = [list(map(float, ln.strip().rstrip(",").split(","))) for ln in open(path) if ln.strip()]
    poses = []
    for i in range(0, len(rows) - 3, 4):
        m = np.array(rows[i:i + 4])
        poses.append((m[:3, :3].copy(), m[:3, 3].copy()))
    return poses


def read_pcd_xyz(path):
    with open(path, "rb") as f:
        n = 0
        fields = []
        while True:
            ln = f.readline().decode("ascii", "replace").strip()
            if ln.startswith("FIELDS"):
                fields = ln.split()[1:]
            if ln.startswith("POINTS"):
                n = int(ln.split()[1])
            if ln.startswith("DATA"):
                assert ln.split()[1] == "binary"
                break
        raw = np.frombuffer(f.read(n * 4 * len(fields)), dtype=np.float32).reshape(n, len(fields))
    return raw[:, :3].astype(np.float64)


class Node:
    __slots__ = ("pts_o", "pts_t", "frame", "center", "quater", "layer", "leaves", "state", "push")

    def __init__(self, pts_o, pts_t, frame, center, quater, layer):
        self.pts_o, self.pts_t, self.frame = pts_o, pts_t, frame
        self.center, self.quater, self.layer = center, quater, layer
        self.leaves, self.state, self.push = [], 0, 0


def judge_eigen(node):
    p = node.pts_t
    n = len(p)
    c = p.sum(0) / n
    cov = p.T @ p / n - np.outer(c, c)
    lam = np.linalg.eigvalsh(cov)
    return lam[0] / lam[1] < float(np.float32(EIGEN_VALUE_ARRAY[node.layer]))  # `float eigen_value_array[]` (bavoxel.hpp:11)


def recut(node, out):
    point_size = len(node.pts_o)
    if point_size <= MIN_PS:
        return
    if judge_eigen(node):
        if point_size > LAYER_SIZE[node.layer]:
            node.state = 2
        if point_size > MIN_PS:
            node.push = 1
            out.append(node)
        return
    if node.layer == LAYER_LIMIT:
        node.state = 2
        return
    node.state = 1
    octant = (node.pts_t > node.center[None, :].astype(np.float32).astype(np.float64))
    leafnum = 4 * octant[:, 0] + 2 * octant[:, 1] + octant[:, 2]
    for lf in range(8):
        m = leafnum == lf
        if not m.any():
            continue
        xyz = np.array([(lf >> 2) & 1, (lf >> 1) & 1, lf & 1])
        child = Node(node.pts_o[m], node.pts_t[m], node.frame[m],
                     (node.center + (2 * xyz - 1) * node.quater).astype(np.float32).astype(np.float64),
                     node.quater / 2, node.layer + 1)
        recut(child, out)


def main():
    poses = read_poses(os.path.join(DATA, "alidarPose.csv"))[:N_POSES]
    R0, p0 = poses[0]
    poses = [(R0.T @ R, R0.T @ (p - p0)) for R, p in poses]  # benchmark_realworld.cpp:163-168
    all_o, all_t, all_f = [], [], []
    for i, (R, p) in enumerate(poses):
        pts = read_pcd_xyz(os.path.join(DATA, f"full{i}.pcd"))[::DECIMATE]
        all_o.append(pts)
        all_t.append(pts @ R.T + p)
        all_f.append(np.full(len(pts), i, dtype=np.int32))
    po, pt, fr = np.concatenate(all_o), np.concatenate(all_t), np.concatenate(all_f)
    loc = (pt / VOXEL_SIZE).astype(np.float32)          # float loc_xyz[3]  (bavoxel.hpp:1172,1180)
    loc = np.where(loc < 0, loc - np.float32(1.0), loc)  # :1181
    key = np.trunc(loc).astype(np.int64)                 # (int64_t) cast :1184
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    key, po, pt, fr = key[order], po[order], pt[order], fr[order]
    change = np.any(key[1:] != key[:-1], axis=1)
    starts = np.concatenate([[0], np.nonzero(change)[0] + 1, [len(key)]])
    planes = []
    for a, b in zip(starts[:-1], starts[1:]):
        center = ((0.5 + key[a]) * VOXEL_SIZE).astype(np.float32).astype(np.float64)  # :1213-1215 (float voxel_center)
        recut(Node(po[a:b], pt[a:b], fr[a:b], center, VOXEL_SIZE / 4.0, 0), planes)
    row_ptr, pose_idx, obs, coe = [0], [], [], []
    for nd in planes:  # tras_opt -> push_voxel (bavoxel.hpp:908-929, 30-51)
        frames = np.unique(nd.frame)
        if len(nd.pts_o) < MIN_PS or len(frames) < 2:
            continue
        for f in frames:
            q = nd.pts_o[nd.frame == f]
            P = q.T @ q
            v = q.sum(0)
            obs.append([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], float(len(q))])
            pose_idx.append(int(f))
        row_ptr.append(len(pose_idx))
        coe.append(float(len(nd.pts_o)))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes  # noqa: E402
    poses12 = scenes.pack_poses([R for R, _ in poses], [p for _, p in poses])
    row_ptr = np.array(row_ptr, dtype=np.int64)
    pose_idx = np.array(pose_idx, dtype=np.int32)
    obs10 = np.array(obs, dtype=np.float64)
    coe = np.array(coe, dtype=np.float64)
    planes_per_pose = np.bincount(pose_idx, minlength=N_POSES)
    print(f"poses {N_POSES}, points {len(po)}, root voxels {len(starts) - 1}, plane voxels {len(coe)}, "
          f"observations {len(pose_idx)}, planes/pose min {planes_per_pose.min()} mean {planes_per_pose.mean():.0f}")
    # golden outputs: the CPU oracle on this input (reference LM constants, bavoxel.hpp:1087,1104,1155)
    from oracle import oracle_py as orc
    o = orc.Oracle(N_POSES, row_ptr, pose_idx, obs10, coe)
    st, poses_out, tr, per = o.damping_iter(poses12, gauge_mode=0)
    assert st == 0
    H, g, r = o.evaluate_threads(poses12, threads=4)
    print("oracle LM:", [(round(t["r1"], 4), round(t["r2"], 4), t["accepted"]) for t in tr])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", OUT),
                        n_poses=N_POSES, row_ptr=row_ptr, pose_idx=pose_idx, obs10=obs10, coe=coe, poses_init=poses12,
                        oracle_poses=poses_out, oracle_r1=np.array([t["r1"] for t in tr]),
                        oracle_r2=np.array([t["r2"] for t in tr]),
                        oracle_accepted=np.array([t["accepted"] for t in tr]), oracle_residual0=r,
                        oracle_g0=g, oracle_Hdiag0=np.diag(H).copy(), oracle_per_iter=per,
                        # the whole Hessian is pinned through three probes (the GPU test also compares every entry with
                        # the oracle run on the box): H 1, H w (w = fixed pseudo-random weights), Frobenius norm
                        oracle_H0_rowsum=H @ np.ones(len(g)), oracle_H0_probe=H @ np.cos(np.arange(len(g)) * 0.7),
                        oracle_H0_fro=np.linalg.norm(H))
    print("wrote", os.path.getsize(os.path.join(ROOT, "tests", "golden", OUT)), "bytes")


if __name__ == "__main__":
    main()
