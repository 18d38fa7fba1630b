"""GPU parity at the BENCHMARKED sizes and configurations ("next round" item 1):

  * BASELINE C3 (500 poses x 100 000 voxels) and C2 (200 poses x 20 000 voxels): a voxel sub-range of the FULL problem,
    downloaded from HBM (balm_download_voxel_range), against the CPU oracle on exactly those voxels -- values, not only
    properties -- in both precisions, for balm_evaluate(head, end) and balm_residual;
  * the benchmark_virtual twin configuration (benchmark_virtual.cpp:375-482: u0 = 0.1, <= 20 iterations, fix cluster in
    the Hessian, pose 0 forced to the identity) per LM iteration against the oracle with the same options;
  * an INDEFINITE damped system (large pose perturbation, tiny damping): balm_solve against the oracle's diagonally
    pivoted LDL^T (the stand-in for Eigen's, bavoxel.hpp:1114);
  * the right-update evaluator acc_evaluate2 (bavoxel.hpp:53-158) as a second reference-side pin of r and g.
Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest

import scenes
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu

PRECS = [pytest.param(0, id="fp64"), pytest.param(1, id="tensor")]
TOLH = {0: 1e-9, 1: 1e-8}


def _pose_err(a, b):
    rot = max(np.linalg.norm(orc.log_so3(scenes.unpack_pose(x)[0].T @ scenes.unpack_pose(y)[0])) for x, y in zip(a, b))
    tra = max(np.linalg.norm(x[9:] - y[9:]) for x, y in zip(a, b))
    return rot, tra


def _subrange_vs_oracle(c, N, init, head, end, prec):
    import balm_b200
    rp, pi, ob, co = c.download_voxel_range(head, end)
    assert rp[0] == 0 and rp[-1] == len(pi) == len(ob) and len(co) == end - head
    o = orc.Oracle(N, rp, pi, ob, co)
    H, g, r = c.evaluate(init, head, end)
    Ho, go, ro = o.evaluate_threads(init, threads=4)
    assert abs(r - ro) <= 1e-12 * abs(ro), (r, ro)
    assert np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(H - Ho).max() <= TOLH[prec] * np.abs(Ho).max()
    assert np.array_equal(H, H.T)
    # balm_residual on the same voxels registered as a problem of their own (host-buffer path)
    c2 = balm_b200.Context(N, 0, prec)
    c2.set_voxels(rp, pi, ob, co)
    r2 = c2.residual(init)
    assert abs(r2 - o.residual(init)) <= 1e-12 * abs(ro)
    H2, g2, _ = c2.evaluate(init)
    # the sub-range of the big problem and the small problem are the same sums: identical to rounding of the
    # partial-sum order (fp64) / to the fixed-point grid of the digit planes (tensor: column scales differ)
    assert np.abs(H2 - H).max() <= (1e-12 if prec == 0 else 2e-8) * np.abs(H).max()
    assert np.abs(g2 - g).max() <= 1e-12 * np.abs(g).max()
    c2.close()
    return o


@pytest.mark.parametrize("prec", PRECS)
def test_c3_subrange_values_vs_oracle(prec):
    """BASELINE C3 at its full size in HBM; 256 voxels out of the middle of it vs the oracle (N = 500: the oracle's
    O(k^2) pair loop takes a few seconds for 256 voxels)."""
    import balm_b200
    N, M = 500, 100000
    c = balm_b200.Context(N, 0, prec)
    gt, init = c.synth_virtual(M, seed=10)
    _subrange_vs_oracle(c, N, init, 61_440, 61_440 + 256, prec)
    c.close()


@pytest.mark.parametrize("prec", PRECS)
def test_c2_subrange_values_and_lm_vs_oracle(prec):
    """BASELINE C2 (200 poses x 20 000 voxels) at full size: 2 048 voxels of it vs the oracle, then the LM loop of that
    2 048-voxel problem per iteration (1e-6 rad / 1e-6 m, same accept/reject sequence)."""
    import balm_b200
    N, M = 200, 20000
    c = balm_b200.Context(N, 0, prec)
    gt, init = c.synth_virtual(M, seed=11)
    head = 7_000
    o = _subrange_vs_oracle(c, N, init, head, head + 2048, prec)
    c.close()
    rp, pi, ob, co = o.row_ptr, o.pose_idx, o.obs10, o.coe
    c = balm_b200.Context(N, 0, prec)
    c.set_voxels(rp, pi, ob, co)
    poses, tr, per = c.damping_iter(init, max_iter=3, want_per_iter=True, gauge_mode=2)
    st, poses_o, tr_o, per_o = o.damping_iter(init, max_iter=3, gauge_mode=2)
    assert st == 0 and [t["accepted"] for t in tr] == [t["accepted"] for t in tr_o]
    for it in range(len(tr)):
        rot, tra = _pose_err(per[it], per_o[it])
        assert rot <= 1e-6 and tra <= 1e-6, (it, rot, tra)
        assert abs(tr[it]["r2"] - tr_o[it]["r2"]) <= (1e-9 if prec == 0 else 1e-7) * abs(tr_o[it]["r2"])


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_poses,n_planes,with_fix", [(20, 150, True), (50, 300, False)])
def test_benchmark_virtual_twin_per_iteration(n_poses, n_planes, with_fix, prec):
    """The single-thread twin of config C1 (benchmark_virtual.cpp:375-482): u0 = 0.1, <= 20 iterations, the fix cluster
    inside the Hessian (:241-243), pose 0 := identity at the end (:472-479) -- per iteration against the oracle run with
    exactly these options (threads = 1)."""
    import balm_b200
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=61, with_fix=with_fix)
    c = balm_b200.Context(n_poses, 0, prec)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    o = orc.Oracle(n_poses, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    kw = dict(max_iter=20, u0=0.1, hess_includes_fix=True)
    poses, tr, per = c.damping_iter(sc["poses_init"], want_per_iter=True, gauge_mode=1, **kw)
    st, poses_o, tr_o, per_o = o.damping_iter(sc["poses_init"], gauge_mode=1, threads=1, **kw)
    assert st == 0 and len(tr) == len(tr_o) >= 3
    assert [t["accepted"] for t in tr] == [t["accepted"] for t in tr_o]
    for it in range(len(tr)):
        rot, tra = _pose_err(per[it], per_o[it])
        assert rot <= 1e-6 and tra <= 1e-6, (it, rot, tra)
        assert abs(tr[it]["r1"] - tr_o[it]["r1"]) <= (1e-9 if prec == 0 else 1e-7) * abs(tr_o[it]["r1"])
        assert abs(tr[it]["u"] - tr_o[it]["u"]) <= 1e-4 * tr_o[it]["u"]
    assert max(_pose_err(poses, poses_o)) <= 1e-6                       # after the gauge step
    assert np.array_equal(poses[0], np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]))
    rot, tra = orc.rmse(poses, poses_o)
    assert rot <= 1e-6 and tra <= 1e-6


def test_indefinite_damped_system_vs_pivoted_ldlt():
    """SURVEY 7 hard part 3: at a strongly perturbed start H has negative eigenvalues, and with tiny damping H + uD is
    INDEFINITE. Eigen's LDLT (bavoxel.hpp:1114) pivots on the diagonal; the library factors without pivoting, checks the
    residual of the solve in fp64 and refines it. Contract: either the solution agrees with the pivoted one, or the
    step is flagged (not_pd) -- it is never silently inaccurate."""
    import balm_b200
    sc = scenes.make_scene(n_poses=12, n_planes=60, seed=71, rot_noise=12 / 57.3, tra_noise=0.6)
    c = balm_b200.Context(12, 0, 0)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
    H, g, r = c.evaluate(sc["poses_init"])
    n = len(g)
    seen_indefinite = 0
    for u in (1e-4, 1e-3, 1e-2):
        A = H + u * np.diag(np.diag(H))
        lam = np.linalg.eigvalsh(A)
        seen_indefinite += lam[0] < 0
        cond = np.abs(lam).max() / np.abs(lam).min()
        dxo, zp = orc.ldlt_solve(A, -g)
        dx, q1, bad = c.solve(u)
        relres = np.linalg.norm(A @ dx + g) / (np.linalg.norm(A, 2) * np.linalg.norm(dx) + np.linalg.norm(g))
        relres_o = np.linalg.norm(A @ dxo + g) / (np.linalg.norm(A, 2) * np.linalg.norm(dxo) + np.linalg.norm(g))
        if not bad:
            # backward error at rounding level -- as good as the pivoted factorisation's ...
            assert relres <= max(1e-14, 10 * relres_o), (u, relres, relres_o)
            # ... hence the two solutions agree to what the conditioning of H + uD allows (two backward-stable solves of
            # a system with condition number `cond` differ by O(cond * eps); here cond ~ 1e8: the gauge directions are
            # held by the tiny damping only)
            assert np.abs(dx - dxo).max() <= 20 * cond * 2.3e-16 * np.abs(dxo).max(), (u, cond, np.abs(dx - dxo).max())
            assert abs(q1 - 0.5 * dx @ (u * np.diag(H) * dx - g)) <= 1e-9 * abs(q1)
    assert seen_indefinite >= 1, "the scene was meant to produce an indefinite damped matrix"
    # and inside the LM loop: the same accept/reject sequence as the oracle from this start (rejections included)
    o = orc.Oracle(12, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
    kw = dict(max_iter=8, u0=1e-4, min_planes_per_pose=0, gauge_mode=2)
    poses, tr, per = c.damping_iter(sc["poses_init"], want_per_iter=True, **kw)
    st, poses_o, tr_o, per_o = o.damping_iter(sc["poses_init"], **kw)
    assert [t["accepted"] for t in tr] == [t["accepted"] for t in tr_o]
    for it in range(len(tr)):
        assert max(_pose_err(per[it], per_o[it])) <= 1e-6, it


@pytest.mark.parametrize("prec", PRECS)
def test_gradient_against_right_update_reference(prec):
    """acc_evaluate2 (bavoxel.hpp:53-158; numpy restatement in tests/numpy_acc2.py): same residual, and a gradient tied
    to the left one by the per-pose adjoint map -- a reference-side check of r and g that shares no code with the
    left-update oracle."""
    import balm_b200
    import numpy_acc2 as a2
    sc = scenes.make_scene(n_poses=10, n_planes=80, seed=81, drop=0.3, with_fix=True, pts_size=16)
    c = balm_b200.Context(10, 0, prec)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    x = sc["poses_init"]
    gR, rR = a2.acc_evaluate2(10, sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], x, sc["fix10"])
    H, g, r = c.evaluate(x, include_fix=True)
    assert abs(r - rR) <= 1e-11 * abs(rR) and abs(c.residual(x) - rR) <= 1e-11 * abs(rR)
    assert np.abs(a2.left_to_right_gradient(g, x) - gR).max() <= 1e-9 * np.abs(gR).max()


def test_empty_shard_is_registered_and_contributes_zero():
    """A rank whose shard holds no voxels (fewer voxels than ranks) registers an empty problem: H = g = r = 0."""
    import balm_b200
    c = balm_b200.Context(5, 0, 0)
    c.set_voxels(np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros((0, 10)), np.zeros(0))
    H, g, r = c.evaluate(np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (5, 1)))
    assert r == 0 and not H.any() and not g.any()
    assert c.residual(np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (5, 1))) == 0


@pytest.mark.parametrize("prec", PRECS)
def test_not_pd_pivot_is_reported_and_rejects_the_step(prec):
    """A zero pivot (u = -1 makes every diagonal entry of H + u diag(H) exactly zero) must raise not_pd in balm_solve, and
    inside the LM loop count as a rejected step (u *= v, poses untouched) -- the reference never checks its LDLT
    (bavoxel.hpp:1114); a step that cannot be computed is the one thing the library refuses instead of following."""
    import balm_b200
    sc = scenes.make_scene(n_poses=9, n_planes=60, seed=72)
    c = balm_b200.Context(9, 0, prec)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
    c.evaluate(sc["poses_init"])
    dx, q1, bad = c.solve(-1.0)
    assert bad
    dx, q1, bad = c.solve(0.01)
    assert not bad and np.isfinite(dx).all()
    poses, tr, _ = c.damping_iter(sc["poses_init"], max_iter=1, u0=-1.0, min_planes_per_pose=0, gauge_mode=2)
    assert len(tr) == 1 and tr[0]["not_pd"] and not tr[0]["accepted"] and tr[0]["u"] == -1.0
    assert np.array_equal(poses, sc["poses_init"])                   # a rejected step leaves the poses alone (:1144-1149)


@pytest.mark.parametrize("n_poses,n_planes", [(3, 12), (11, 40), (22, 60), (130, 60), (320, 40)])
def test_solve_edge_sizes_vs_numpy(n_poses, n_planes):
    """The persistent tile-DAG factorisation at sizes around its tile boundaries: n = 18 (one partial tile), 66 (one full
    tile + 2 columns), 132, 780 (12 tiles + 12 columns: near and far worker groups, partial last block), 1920 (30 tiles)."""
    import balm_b200
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=73, pts_size=8)
    c = balm_b200.Context(n_poses, 0, 0)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
    H, g, r = c.evaluate(sc["poses_init"])
    for u in (0.01, 1.0):
        dx, q1, bad = c.solve(u)
        A = H + u * np.diag(np.diag(H))
        ref = np.linalg.solve(A, -g)
        assert not bad
        assert np.abs(dx - ref).max() <= 1e-9 * max(1e-3, np.abs(ref).max()), (n_poses, u)
        assert abs(q1 - 0.5 * ref @ (u * np.diag(H) * ref - g)) <= 1e-9 * abs(q1)
    dx2, _, _ = c.solve(1.0)
    assert np.array_equal(dx, dx2)                                    # bit-reproducible
