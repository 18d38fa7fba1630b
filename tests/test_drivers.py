"""Headless benchmark drivers (balm_b200/drivers.py) on the GPU: files on disk -> refined poses."""
import numpy as np
import pytest

import assoc_ref
import scenes
from balm_b200 import drivers, io

pytestmark = pytest.mark.gpu


def _write_dataset(tmp_path, n_poses=10, pts=5000):
    pts_b, frames, poses = assoc_ref.synthetic_scans(n_poses=n_poses, pts_per_scan=pts, seed=9)
    # an arbitrary world frame: the driver must re-anchor everything to pose 0 (benchmark_realworld.cpp:163-168)
    Rw, pw = scenes.exp_so3(np.array([0.3, -0.2, 0.5])), np.array([4.0, -2.0, 1.0])
    R = np.stack([Rw @ r for r, _ in poses])
    p = np.stack([Rw @ t + pw for _, t in poses])
    io.write_pose_csv(tmp_path / "alidarPose.csv", R, p, np.arange(n_poses) * 0.1)
    for i in range(n_poses):
        io.write_pcd(tmp_path / f"full{i}.pcd", pts_b[frames == i], binary=(i % 2 == 0))  # both encodings
    return pts_b, frames, poses


def test_benchmark_realworld_from_files(tmp_path, capsys):
    import balm_b200
    pts_b, frames, poses = _write_dataset(tmp_path)
    res = drivers.benchmark_realworld(str(tmp_path), voxel_size=2.0)
    out = capsys.readouterr().out
    assert res is not None and "The size of poses: 10" in out and "iter0: (" in out
    # same numbers as the library called directly on the re-anchored inputs
    R0, p0 = poses[0]
    anchored = scenes.pack_poses([R0.T @ r for r, _ in poses], [R0.T @ (t - p0) for _, t in poses])
    assert np.abs(res["poses_init"] - anchored).max() < 1e-12
    c = balm_b200.Context(len(poses), 0, balm_b200.PREC_TENSOR)
    M, K = c.cut_voxels(pts_b, frames, res["poses_init"], voxel_size=2.0)
    assert (M, K) == (res["n_voxels"], res["n_obs"]) and M >= 3 * len(poses)
    ref, tr, _ = c.damping_iter(res["poses_init"])
    assert np.abs(ref - res["poses"]).max() < 1e-9 and len(tr) == len(res["trace"])
    assert res["trace"][-1]["r2"] < res["trace"][0]["r1"]           # the cost went down
    assert np.abs(res["poses"][0] - np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])).max() < 1e-12  # gauge: pose 0


def test_benchmark_realworld_plane_guard(tmp_path, capsys):
    _write_dataset(tmp_path, n_poses=6, pts=300)                     # far too few points for 18 planes
    assert drivers.benchmark_realworld(str(tmp_path), voxel_size=2.0) is None
    out = capsys.readouterr().out
    assert "Initial error too large." in out and "The optimization is terminated." in out


def test_benchmark_virtual_reaches_the_noise_floor(capsys):
    res = drivers.benchmark_virtual(winSize=20, sufSize=150, ptsSize=40, point_noise=0.01, seed=3)
    out = capsys.readouterr().out
    assert "winSize: 20" in out and "RSME: " in out
    rot, tran = res["rsme"]
    rot0, tran0 = res["rsme_init"]
    assert rot < 0.05 * rot0 and tran < 0.05 * tran0 and rot * 57.3 < 0.05 and tran < 0.005
    assert np.abs(res["poses"][0] - np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])).max() == 0.0  # :472-479


def test_cli_entry_points(tmp_path, capsys):
    from balm_b200 import benchmark_realworld, benchmark_virtual
    _write_dataset(tmp_path)
    out_csv = tmp_path / "refined.csv"
    assert benchmark_realworld.main(["--file_path", str(tmp_path), "--voxel_size", "2", "--out", str(out_csv)]) == 0
    R, p, _ = io.read_pose_csv(out_csv)
    assert len(R) == 10 and np.allclose(R[0], np.eye(3)) and np.allclose(p[0], 0)
    assert benchmark_virtual.main(["--winSize", "8", "--sufSize", "60", "--point_noise", "0.01"]) == 0
    assert "RSME: " in capsys.readouterr().out
