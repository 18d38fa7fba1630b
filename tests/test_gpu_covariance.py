"""GPU: balm_pose_covariance against the numpy restatement of left_jacobian_point / multi_second /
H^-1 Rcov H^-T (tests/numpy_cov.py; BAs_left.hpp:342-473, 995-1023, 1089-1096), through the C ABI."""
import numpy as np
import pytest

import numpy_cov as nc
import scenes
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("n_poses,n_planes,drop,with_fix", [
    (6, 30, 0.0, True),      # dense, fix clusters (the consistency experiment's sliding window has them)
    (14, 70, 0.4, True),     # ragged co-visibility
    (40, 90, 0.0, True),     # several pose tiles, n = 240 (two 128-column blocks)
])
def test_pose_covariance_matches_reference_restatement(n_poses, n_planes, drop, with_fix, prec):
    import balm_b200
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=91, drop=drop, with_fix=with_fix, pts_size=20)
    c = balm_b200.Context(n_poses, 0, prec)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    poses, tr, _ = c.damping_iter(sc["poses_init"], gauge_mode=2, hess_includes_fix=True)   # covariance at the optimum
    pn = 0.01
    raw, cov = c.pose_covariance(poses, point_noise=pn, include_fix=True)
    ref = nc.left_jacobian_point(n_poses, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], poses, sc["fix10"], pnoise=pn)
    assert np.abs(raw - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.array_equal(raw, raw.T)
    o = orc.Oracle(n_poses, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    H = o.evaluate(poses, include_fix=True)[0]
    ref_cov = nc.pose_covariance(H, ref)
    assert np.abs(cov - ref_cov).max() <= 1e-7 * np.abs(ref_cov).max()
    assert np.abs(cov - cov.T).max() <= 1e-9 * np.abs(cov).max()
    assert np.linalg.eigvalsh(0.5 * (cov + cov.T)).min() >= -1e-9 * np.abs(cov).max()
    # explicit per-observation covariances (PointCluster::c_cov, toolss.hpp:288): same answer when they are the isotropic ones
    cc = np.stack([nc.cluster_cov_isotropic(o10, pn) for o10 in sc["obs10"]])
    raw2, _ = c.pose_covariance(poses, c_cov=cc, include_fix=True, want_cov=False)
    assert np.abs(raw2 - raw).max() <= 1e-12 * np.abs(raw).max()


def test_pose_covariance_batches_and_singular_hessian(monkeypatch):
    import balm_b200
    sc = scenes.make_scene(n_poses=10, n_planes=120, seed=92, drop=0.3, with_fix=True, pts_size=12)
    c = balm_b200.Context(10, 0, 0)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    raw1, _ = c.pose_covariance(sc["poses_gt"], point_noise=0.02, include_fix=True, want_cov=False)
    monkeypatch.setenv("BALM_COV_BUDGET_MB", "1")     # 1 MiB row buffer -> many voxel batches
    raw2, _ = c.pose_covariance(sc["poses_gt"], point_noise=0.02, include_fix=True, want_cov=False)
    assert np.abs(raw1 - raw2).max() <= 1e-12 * np.abs(raw1).max()
    # without fix clusters in the Hessian the gauge is free: H is singular and the propagation is refused, not garbage
    with pytest.raises(balm_b200.BalmError) as e:
        c.pose_covariance(sc["poses_gt"], point_noise=0.02, include_fix=False)
    assert e.value.status == 3
