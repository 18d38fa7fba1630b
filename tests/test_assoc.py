"""GPU association (balm_cut_voxels: cut_voxel hashing + PointCluster accumulation + octree recut + tras_opt/push_voxel,
bavoxel.hpp:1170-1223, 654-776, 908-929, 30-51) against the numpy restatement tests/assoc_ref.py:
identical set of plane voxels (63-bit node keys), identical observing frames, clusters equal to summation rounding,
and the BA that follows gives the same poses. Inputs: synthetic lidar-like scans, and a decimated slice of the
reference's own dataset (tests/golden/realworld_points_small.npz, made from datas/benchmark_realworld)."""
import os

import numpy as np
import pytest

import assoc_ref
import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "realworld_points_small.npz")


def test_assoc_oracle_basics():
    pts, frs, poses = assoc_ref.synthetic_scans(n_poses=6, pts_per_scan=3000)
    rp, pi, obs, coe, keys = assoc_ref.cut_voxels(pts, frs, poses)
    assert len(coe) > 50 and np.all(np.diff(keys) > 0)
    k = np.diff(rp)
    assert k.min() >= 2 and np.all(coe > 15)                      # push_voxel / recut thresholds
    assert np.allclose([obs[a:b, 9].sum() for a, b in zip(rp[:-1], rp[1:])], coe)   # coe = sum of N (bavoxel.hpp:42-44)
    for a, b in zip(rp[:-1], rp[1:]):
        assert np.all(np.diff(pi[a:b]) > 0)


def _compare(ctx, pts, frs, poses, **kw):
    poses12 = scenes.pack_poses([r for r, _ in poses], [p for _, p in poses])
    M, K = ctx.cut_voxels(pts, frs, poses12, **kw)
    rp, pi, obs, coe = ctx.download_voxels()
    rp0, pi0, obs0, coe0, keys0 = assoc_ref.cut_voxels(pts.astype(np.float64), frs, poses, **kw)
    assert (M, K) == (len(coe0), len(pi0))
    assert np.array_equal(rp, rp0) and np.array_equal(pi, pi0) and np.array_equal(coe, coe0)
    assert np.array_equal(obs[:, 9], obs0[:, 9])
    scale = np.abs(obs0).max(axis=0)
    assert np.all(np.abs(obs - obs0) <= 1e-11 * scale)
    return poses12, (rp0, pi0, obs0, coe0)


@pytest.mark.gpu
def test_gpu_association_matches_oracle_synthetic():
    import balm_b200
    pts, frs, poses = assoc_ref.synthetic_scans(n_poses=12, pts_per_scan=6000)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(frs))                 # arbitrary point order: the library sorts by frame itself
    for order in (np.arange(len(frs)), perm):
        c = balm_b200.Context(12, 0, 1)
        poses12, ref = _compare(c, pts[order].astype(np.float32), frs[order], poses)
    # the BA on GPU-associated voxels equals the BA on oracle-associated voxels
    p1, tr1, _ = c.damping_iter(poses12, min_planes_per_pose=0)
    c2 = balm_b200.Context(12, 0, 1)
    c2.set_voxels(*ref)
    p2, tr2, _ = c2.damping_iter(poses12, min_planes_per_pose=0)
    assert len(tr1) == len(tr2) and np.abs(p1 - p2).max() <= 1e-9
    assert tr1[-1]["r2"] < tr1[0]["r1"]


@pytest.mark.gpu
@pytest.mark.parametrize("layer_limit,voxel_size", [(2, 2.0), (1, 2.0), (0, 1.0), (2, 0.7)])
def test_gpu_association_options(layer_limit, voxel_size):
    import balm_b200
    pts, frs, poses = assoc_ref.synthetic_scans(n_poses=8, pts_per_scan=5000, seed=9)
    c = balm_b200.Context(8, 0, 0)
    _compare(c, pts.astype(np.float32), frs, poses, layer_limit=layer_limit, voxel_size=voxel_size)


@pytest.mark.gpu
def test_gpu_association_realworld_slice():
    import balm_b200
    d = np.load(GOLD)
    poses = [scenes.unpack_pose(p) for p in d["poses"]]
    c = balm_b200.Context(len(poses), 0, 1)
    poses12, ref = _compare(c, d["xyz"], d["frame"], poses, voxel_size=2.0, eigen_value_array=(1 / 16, 1 / 16, 1 / 9))
    r = c.residual(poses12)
    assert np.isfinite(r) and r > 0
