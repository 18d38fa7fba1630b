"""Real-data golden fixture: plane voxels cut from the reference's own dataset (datas/benchmark_realworld, first 48
scans) by a numpy restatement of the reference association (tests/golden/make_realworld_fixture.py), with the CPU
oracle's outputs on it. Ragged co-visibility (30 793 observations over 1002 voxels x 48 poses), real lidar coordinates.
CPU: the oracle still reproduces the committed vectors. GPU: the CUDA path matches them through the C ABI."""
import os

import numpy as np
import pytest

import scenes
from oracle import oracle_py as orc

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# realworld_voxels.npz : first 48 scans, points decimated x2   (1002 voxels, 30 793 observations)
# realworld_c5_177.npz : BASELINE config C5 at its real window: ALL 177 scans of datas/benchmark_realworld, points
#                        decimated x12 (1606 voxels, 76 564 observations, co-visibility 2..177 poses per voxel)
FIXTURES = ["realworld_voxels.npz", "realworld_c5_177.npz"]


@pytest.fixture(scope="module", params=FIXTURES)
def gold(request):
    d = np.load(os.path.join(GOLD_DIR, request.param))
    return {k: d[k] for k in d.files}


def _pose_err(a, b):
    rot = max(np.linalg.norm(orc.log_so3(scenes.unpack_pose(x)[0].T @ scenes.unpack_pose(y)[0])) for x, y in zip(a, b))
    tra = max(np.linalg.norm(x[9:] - y[9:]) for x, y in zip(a, b))
    return rot, tra


def _check_H_probes(H, gold, tol):
    """Every entry of H enters the three committed probes (H 1, H w, ||H||_F)."""
    n = H.shape[0]
    scale = np.abs(H).max() * n
    assert np.abs(H @ np.ones(n) - gold["oracle_H0_rowsum"]).max() <= tol * scale
    assert np.abs(H @ np.cos(np.arange(n) * 0.7) - gold["oracle_H0_probe"]).max() <= tol * scale
    assert abs(np.linalg.norm(H) - gold["oracle_H0_fro"]) <= tol * gold["oracle_H0_fro"]


def test_oracle_reproduces_golden_vectors(gold):
    N = int(gold["n_poses"])
    o = orc.Oracle(N, gold["row_ptr"], gold["pose_idx"], gold["obs10"], gold["coe"])
    H, g, r = o.evaluate_threads(gold["poses_init"], threads=4)
    assert abs(r - gold["oracle_residual0"]) <= 1e-12 * abs(r)
    assert np.abs(g - gold["oracle_g0"]).max() <= 1e-11 * np.abs(g).max()
    assert np.abs(np.diag(H) - gold["oracle_Hdiag0"]).max() <= 1e-11 * np.abs(np.diag(H)).max()
    _check_H_probes(H, gold, 1e-11)
    st, poses, tr, per = o.damping_iter(gold["poses_init"], gauge_mode=0)
    assert st == 0 and [t["accepted"] for t in tr] == list(gold["oracle_accepted"].astype(bool))
    assert np.allclose([t["r2"] for t in tr], gold["oracle_r2"], rtol=1e-10)
    assert max(_pose_err(poses, gold["oracle_poses"])) <= 1e-9
    # the LM run on real data decreases the cost monotonically (benchmark_realworld has no recorded expectation)
    assert all(b <= a for a, b in zip(gold["oracle_r2"], gold["oracle_r2"][1:]))
    # every pose is seen by >= 20 planes (the reference's precheck, bavoxel.hpp:1079)
    assert np.bincount(gold["pose_idx"], minlength=N).min() >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("prec", [0, 1])
def test_gpu_matches_golden_realworld(gold, prec):
    import balm_b200
    N = int(gold["n_poses"])
    c = balm_b200.Context(N, 0, prec)
    c.set_voxels(gold["row_ptr"], gold["pose_idx"], gold["obs10"], gold["coe"])
    H, g, r = c.evaluate(gold["poses_init"])
    assert abs(r - gold["oracle_residual0"]) <= 1e-11 * abs(r)
    assert np.abs(g - gold["oracle_g0"]).max() <= 1e-10 * np.abs(g).max()
    tolH = 1e-9 if prec == 0 else 1e-8
    assert np.abs(np.diag(H) - gold["oracle_Hdiag0"]).max() <= tolH * np.abs(np.diag(H)).max()
    # the WHOLE Hessian: against the committed probes and, entry by entry, against the oracle run on this box
    _check_H_probes(H, gold, tolH)
    Ho, go, ro = orc.Oracle(N, gold["row_ptr"], gold["pose_idx"], gold["obs10"], gold["coe"]).evaluate_threads(
        gold["poses_init"], threads=4)
    assert np.abs(H - Ho).max() <= tolH * np.abs(Ho).max() and np.array_equal(H, H.T)
    poses, tr, per = c.damping_iter(gold["poses_init"], gauge_mode=2, want_per_iter=True)
    assert [t["accepted"] for t in tr] == list(gold["oracle_accepted"].astype(bool))
    for it in range(len(tr)):
        rot, tra = _pose_err(per[it], gold["oracle_per_iter"][it])
        assert rot <= 1e-6 and tra <= 1e-6, (it, rot, tra)
    final, _, _ = c.damping_iter(gold["poses_init"], gauge_mode=0)
    assert max(_pose_err(final, gold["oracle_poses"])) <= 1e-6
