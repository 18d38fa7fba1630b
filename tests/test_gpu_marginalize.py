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


def test_sliding_window_on_the_device_matches_the_restatement():
    """The whole sliding-window step with nothing leaving HBM: associate a window of scans
    (balm_cut_voxels), optimise, retire the two oldest scans into the fix clusters (balm_marginalize), associate a NEW
    scan with the voxels already there (balm_append_scan: leaf lookup through the octree keys + recut's re-judgement),
    optimise again -- against assoc_ref.cut_voxels / marginalize_ref / append_scan_ref step by step."""
    import balm_b200
    N, mg = 10, 2
    pts, frs, poses = assoc_ref.synthetic_scans(n_poses=N + mg, pts_per_scan=5000, seed=13)
    poses12_all = scenes.pack_poses([r for r, _ in poses], [p for _, p in poses])
    win = frs < N                                               # the first N scans fill the window
    c = balm_b200.Context(N, 0, 0)
    kw = dict(voxel_size=2.0, layer_limit=2, min_ps=15, eigen_value_array=(1 / 16, 1 / 16, 1 / 16))
    M0, K0 = c.cut_voxels(pts[win].astype(np.float32), frs[win], poses12_all[:N], **kw)
    rp, pi, ob, co, keys, lays = assoc_ref.cut_voxels(pts[win], frs[win], poses[:N], with_layers=True, **kw)
    kg, lg = c.download_keys(with_layers=True)
    assert np.array_equal(kg.astype(np.int64), keys) and np.array_equal(lg, lays) and M0 == len(co)
    # the case the layers exist for: a layer-1 leaf in octant 7 carries the same key digits as a layer-0 / layer-2 pattern
    assert ((keys & 63) == 63).sum() > (lays == 0).sum() or ((keys & 7) == 7).sum() > (lays <= 1).sum()
    # window BA (fix clusters empty so far)
    p1, tr1, _ = c.damping_iter(poses12_all[:N], gauge_mode=2, min_planes_per_pose=0)
    # retire the two oldest scans
    M1, K1 = c.marginalize(mg, p1, min_ps=15)
    rp1, pi1, ob1, fx1, co1 = assoc_ref.marginalize_ref(N, rp, pi, ob, None, p1, mg, 15)
    keep = _kept_mask(rp, pi, ob, mg, 15)
    keys1, lays1 = keys[keep], lays[keep]
    kg, lg = c.download_keys(with_layers=True)
    assert M1 == len(co1) and np.array_equal(kg.astype(np.int64), keys1) and np.array_equal(lg, lays1)
    # the window shifts: slots 0..N-mg-1 hold the remaining scans, the two new scans take slots N-mg and N-mg+1
    shifted = np.vstack([p1[mg:], poses12_all[N:N + mg]])
    kcur, lcur, rcur, picur, obcur, fxcur, cocur = keys1, lays1, rp1, pi1, ob1, fx1, co1
    for j in range(mg):
        slot = N - mg + j
        new = pts[frs == N + j]
        M2, K2, matched = c.append_scan(new.astype(np.float32), shifted, slot)
        kcur, lcur, rcur, picur, obcur, fxcur, cocur, matched_ref = assoc_ref.append_scan_ref(
            kcur, lcur, rcur, picur, obcur, fxcur, cocur, new, shifted, slot, **kw)
        assert matched == matched_ref and matched > 0.3 * len(new)
        assert M2 == len(cocur) and K2 == len(picur)
        kg, lg = c.download_keys(with_layers=True)
        assert np.array_equal(kg.astype(np.int64), kcur) and np.array_equal(lg, lcur)
        rp_g, pi_g, ob_g, co_g = c.download_voxels()
        assert np.array_equal(rp_g, rcur) and np.array_equal(pi_g, picur) and np.array_equal(co_g, cocur)
        assert np.all(np.abs(ob_g - obcur) <= 1e-11 * np.abs(obcur).max(axis=0))
        assert np.abs(c.download_fix() - fxcur).max() <= 1e-12 * np.abs(fxcur).max()
        assert (pi_g == slot).sum() > 20                                  # the new scan really joined many voxels
    # the window is full again: the next BA runs on it (the fix clusters carry the retired scans' information)
    p2, tr2, _ = c.damping_iter(shifted, gauge_mode=2, min_planes_per_pose=0, hess_includes_fix=True)
    assert np.isfinite(tr2[-1]["r2"]) and tr2[-1]["r2"] <= tr2[0]["r1"]
    o = orc.Oracle(N, rcur, picur, obcur, cocur, fxcur)
    st, p2o, tr2o, _ = o.damping_iter(shifted, gauge_mode=2, min_planes_per_pose=0, hess_includes_fix=True)
    assert [t["accepted"] for t in tr2] == [t["accepted"] for t in tr2o]
    assert np.abs(p2 - p2o).max() <= 1e-6


def _kept_mask(row_ptr, pose_idx, obs10, mg, min_ps):
    keep = []
    for a in range(len(row_ptr) - 1):
        ss = [s for s in range(row_ptr[a], row_ptr[a + 1]) if pose_idx[s] >= mg]
        keep.append(len(ss) >= 2 and int(sum(obs10[s][9] for s in ss)) >= min_ps)
    return np.array(keep, dtype=bool)
