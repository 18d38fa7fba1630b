"""CPU checks of the covariance-propagation oracle (tests/numpy_cov.py, restating BAs_left.hpp:342-473, 1089-1096):
Ls must be the derivative of the left-update gradient with respect to the 9 parameters of one cluster -- verified by
central differences of the C oracle's gradient -- and the isotropic c_cov must equal what PointCluster::push accumulates
point by point (toolss.hpp:311-343)."""
import numpy as np
import pytest

import numpy_cov as nc
import scenes
from oracle import oracle_py as orc


def test_isotropic_cluster_cov_equals_pointwise_accumulation():
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(17, 3)) * [2.0, 0.5, 1.0] + [1.0, -2.0, 0.3]
    pn = 0.05
    acc = np.zeros((9, 9))
    for x, y, z in pts:                                  # toolss.hpp:317-341
        Bi = np.array([[2 * x, 0, 0], [y, x, 0], [z, 0, x], [0, 2 * y, 0], [0, z, y], [0, 0, 2 * z]])
        Bf = np.vstack([Bi, np.eye(3)])
        acc += Bf @ (np.eye(3) * pn * pn) @ Bf.T
    P = pts.T @ pts
    v = pts.sum(0)
    o10 = np.array([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], len(pts)])
    assert np.abs(nc.cluster_cov_isotropic(o10, pn) - acc).max() <= 1e-12 * np.abs(acc).max()


@pytest.mark.parametrize("drop,with_fix", [(0.0, False), (0.4, True)])
def test_Ls_is_the_gradient_jacobian_wrt_cluster_parameters(drop, with_fix):
    sc = scenes.make_scene(n_poses=6, n_planes=12, seed=9, drop=drop, with_fix=with_fix, pts_size=15)
    x = sc["poses_init"]
    rng = np.random.default_rng(0)
    coe = sc["coe"] * rng.uniform(0.5, 1.5, len(sc["coe"]))       # Ls must carry the voxel weight
    Rcov, Ls = nc.left_jacobian_point(6, sc["row_ptr"], sc["pose_idx"], sc["obs10"], coe, x, sc["fix10"], pnoise=0.01,
                                      return_Ls=True)
    assert np.abs(Rcov - Rcov.T).max() <= 1e-12 * np.abs(Rcov).max()
    assert np.linalg.eigvalsh(Rcov).min() >= -1e-9 * np.abs(Rcov).max()

    def grad(obs):
        o = orc.Oracle(6, sc["row_ptr"], sc["pose_idx"], obs, coe, sc["fix10"])
        return o.evaluate(x, include_fix=with_fix)[1]

    for (a, s) in list(Ls)[::7]:
        for q in (0, 1, 4, 6, 8):                        # P00, P01, P12, v0, v2
            h = 1e-6 * max(1.0, abs(sc["obs10"][s, q]))
            op, om = sc["obs10"].copy(), sc["obs10"].copy()
            op[s, q] += h
            om[s, q] -= h
            fd = (grad(op) - grad(om)) / (2 * h)
            assert np.abs(fd - Ls[(a, s)][:, q]).max() <= 2e-6 * max(1e-3, np.abs(Ls[(a, s)][:, q]).max()), (a, s, q)


def test_voxel_range_split_adds_up():
    sc = scenes.make_scene(n_poses=5, n_planes=9, seed=4, drop=0.3, pts_size=10)
    args = (5, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["poses_init"])
    full = nc.left_jacobian_point(*args, pnoise=0.02)
    parts = sum(nc.left_jacobian_point(*args, pnoise=0.02, beg=b, end=e) for b, e in ((0, 2), (2, 4), (4, 6), (6, 9)))
    assert np.abs(full - parts).max() <= 1e-13 * np.abs(full).max()   # multi_second: 4 ranges summed (BAs_left.hpp:995-1023)
