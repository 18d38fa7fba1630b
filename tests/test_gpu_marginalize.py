"""GPU: balm_marginalize (to_margi + marginalize + the re-registration of tras_opt/push_voxel on the
voxel set held in HBM) against the numpy restatement tests/assoc_ref.py::marginalize_ref, and the BA that follows."""
import numpy as np
import pytest

import assoc_ref
import scenes
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_fix,drop,mg", [(False, 0.0, 3), (True, 0.5, 4), (False, 0.6, 1)])
def test_marginalize_matches_reference_restatement(with_fix, drop, mg):
    import balm_b200
    N = 12
    sc = scenes.make_scene(n_poses=N, n_planes=150, seed=95, drop=drop, with_fix=with_fix, pts_size=9)
    if with_fix:                      # some fix clusters already hold >= 50 points: they must NOT absorb (bavoxel.hpp:790)
        sc["fix10"][::3] *= 8.0
    c = balm_b200.Context(N, 0, 0)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    poses, tr, _ = c.damping_iter(sc["poses_init"], gauge_mode=2, min_planes_per_pose=0)
    M2, K2 = c.marginalize(mg, poses, min_ps=15)
    rp, pi, ob, fx, co = assoc_ref.marginalize_ref(N, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["fix10"], poses, mg, 15)
    assert M2 == len(co) and K2 == len(pi) and 0 < M2 < 150 + 1
    rp_g, pi_g, ob_g, co_g = c.download_voxels()
    assert np.array_equal(rp_g, rp) and np.array_equal(pi_g, pi) and np.array_equal(co_g, co)
    assert np.array_equal(ob_g, ob)                                  # the window clusters are moved, not recomputed
    fx_g = c.download_fix()
    assert np.abs(fx_g - fx).max() <= 1e-12 * np.abs(fx).max()
    # the shifted window evaluates like a freshly registered problem with these fix clusters (poses shifted the same way)
    # (evaluated at the perturbed start, where the gradient is not a difference of cancelling terms)
    shifted = np.vstack([sc["poses_init"][mg:], np.tile(sc["poses_init"][-1], (mg, 1))])
    o = orc.Oracle(N, rp, pi, ob, co, fx)
    H, g, r = c.evaluate(shifted, include_fix=True)
    Ho, go, ro = o.evaluate(shifted, include_fix=True)
    assert abs(r - ro) <= 1e-12 * abs(ro) and np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max()
    assert abs(c.residual(shifted) - o.residual(shifted)) <= 1e-12 * abs(ro)
