"""Synthetic plane-feature scenes of the benchmark_virtual shape, in numpy (test-side generator).

Follows /root/reference/src/benchmark/benchmark_virtual.cpp:547-606 (scene) and :491-503 (pose noise):
trajectory R_i = Exp(i/N*rotEnd), p_i = i/N*traEnd with |rotEnd| = 0.5 rad, |traEnd| = 1 m; the first three
planes axis-aligned, others Exp(U(-pi,pi)^3); centre U(-surf_range,surf_range)^3; every pose sees every plane
with pts_size points (U(-.5,.5), U(-.5,.5), N(0,point_noise)) expressed in the body frame and rounded to
float32 (pcl::PointXYZINormal stores floats, benchmark_virtual.cpp:600-602); clusters built by
PointCluster::push (tools.hpp:311-316); coe = winSize*ptsSize (:391).  numpy's RNG replaces
std::default_random_engine (the reference seeds with time(0), so no stream is reproducible anyway).
`drop` removes a random fraction of observations to exercise ragged/sparse co-visibility.
"""
import numpy as np


def exp_so3(phi):
    phi = np.asarray(phi, dtype=np.float64)
    n = np.linalg.norm(phi)
    if n < 1e-11:
        return np.eye(3)
    a = phi / n
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(n) * K + (1 - np.cos(n)) * K @ K


def pack_poses(Rs, ps):
    """list of (R 3x3, p 3) -> N x 12 (R column-major, then p)"""
    return np.stack([np.concatenate([R.T.reshape(9), p]) for R, p in zip(Rs, ps)])


def unpack_pose(p12):
    return p12[:9].reshape(3, 3).T, p12[9:12]


def make_scene(n_poses=6, n_planes=40, pts_size=40, point_noise=0.01, surf_range=2.0, seed=10, drop=0.0,
               with_fix=False, rot_noise=2 / 57.3, tra_noise=0.1):
    rng = np.random.default_rng(seed)
    rot_end = rng.normal(-1, 1, 3)
    tra_end = rng.normal(-1, 1, 3)
    rot_end = rot_end / np.linalg.norm(rot_end) * 0.5
    tra_end = tra_end / np.linalg.norm(tra_end) * 1.0
    Rs = [exp_so3(i / n_poses * rot_end) for i in range(n_poses)]
    ps = [i / n_poses * tra_end for i in range(n_poses)]
    row_ptr = [0]
    pose_idx, obs, coe, fix = [], [], [], []
    for s in range(n_planes):
        if s < 3:
            fd = np.zeros(3)
            fd[s] = np.pi / 2
            rot = exp_so3(fd)
        else:
            rot = exp_so3(rng.uniform(-np.pi, np.pi, 3))
        center = rng.uniform(-surf_range, surf_range, 3)
        seen = []
        for j in range(n_poses):
            if drop > 0 and rng.uniform() < drop:
                continue
            seen.append(j)
        if len(seen) < 2:  # push_voxel skips voxels with <2 observing poses (bavoxel.hpp:37)
            seen = [0, n_poses - 1]
        n_tot = 0
        for j in seen:
            loc = np.stack([rng.uniform(-0.5, 0.5, pts_size), rng.uniform(-0.5, 0.5, pts_size),
                            rng.normal(0, point_noise, pts_size)], axis=1)
            w = loc @ rot.T + center
            b = (w - ps[j]) @ Rs[j]  # R^T (x - p)
            b = b.astype(np.float32).astype(np.float64)
            P = b.T @ b
            v = b.sum(0)
            obs.append([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], float(pts_size)])
            pose_idx.append(j)
            n_tot += pts_size
        row_ptr.append(len(pose_idx))
        if with_fix:
            loc = np.stack([rng.uniform(-0.5, 0.5, pts_size), rng.uniform(-0.5, 0.5, pts_size),
                            rng.normal(0, point_noise, pts_size)], axis=1)
            w = loc @ rot.T + center
            P = w.T @ w
            v = w.sum(0)
            fix.append([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], float(pts_size)])
        coe.append(float(n_poses * pts_size) if drop == 0 else float(n_tot))
    poses_gt = pack_poses(Rs, ps)
    Rn, pn = [], []
    for i in range(n_poses):
        rv = rng.normal(0, rot_noise, 3) / 1.732
        tv = rng.normal(0, tra_noise, 3) / 1.732
        Rn.append(Rs[i] @ exp_so3(rv))
        pn.append(ps[i] + tv)
    poses_init = pack_poses(Rn, pn)
    return dict(n_poses=n_poses, row_ptr=np.array(row_ptr, dtype=np.int64),
                pose_idx=np.array(pose_idx, dtype=np.int32), obs10=np.array(obs, dtype=np.float64),
                coe=np.array(coe, dtype=np.float64),
                fix10=(np.array(fix, dtype=np.float64) if with_fix else None),
                poses_gt=poses_gt, poses_init=poses_init)
