"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs.  Tolerances: residual rel 1e-12, gradient rel 1e-10, Hessian rel 1e-9
(fp64 mode), pose update per LM iteration 1e-6 rad / 1e-6 m with the same accept/reject sequence.
"""
import numpy as np
import pytest

import scenes
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu


def _ctx(sc, precision=0):
    import balm_b200
    c = balm_b200.Context(sc["n_poses"], 0, precision)
    c.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])
    return c


def _oracle(sc):
    return orc.Oracle(sc["n_poses"], sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["fix10"])


def _check_eval(c, o, poses, include_fix=False, head=0, end=None, tolH=1e-9, tolr=1e-12):
    H, g, r = c.evaluate(poses, head, end, include_fix=include_fix)
    Ho, go, ro = o.evaluate(poses, head, end, include_fix=include_fix)
    assert abs(r - ro) <= tolr * abs(ro), (r, ro)
    assert np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(H - Ho).max() <= tolH * np.abs(Ho).max()
    assert np.array_equal(H, H.T)
    return H, g, r


def _pose_err(a, b):
    rot = max(np.linalg.norm(orc.log_so3(scenes.unpack_pose(x)[0].T @ scenes.unpack_pose(y)[0])) for x, y in zip(a, b))
    tra = max(np.linalg.norm(x[9:] - y[9:]) for x, y in zip(a, b))
    return rot, tra


PRECS = [pytest.param(0, id="fp64"), pytest.param(1, id="tensor")]
# Hessian tolerance: fp64 path 1e-9; tcgen05 split-integer path: G' is rounded to 30 bits below
# each column maximum before the EXACT integer accumulation -> 1e-8 of max|H| (measured 2e-10 .. 4e-9)
TOLH = {0: 1e-9, 1: 1e-8}


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_poses,n_planes,drop,with_fix", [
    (6, 40, 0.0, False),      # dense, one warp tile
    (50, 300, 0.0, False),    # dense, benchmark_virtual C1 shape (scaled down in M)
    (37, 64, 0.0, False),     # dense, N not a multiple of 32, n not a multiple of 128
    (20, 200, 0.4, False),    # ragged / sparse co-visibility (pose-major lists)
    (12, 50, 0.0, True),      # fix cluster present
    (9, 30, 0.5, True),       # sparse + fix
    (2, 5, 0.0, False),       # minimum: two poses
])
def test_evaluate_matches_oracle(n_poses, n_planes, drop, with_fix, prec):
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=21, drop=drop, with_fix=with_fix)
    c, o = _ctx(sc, prec), _oracle(sc)
    _check_eval(c, o, sc["poses_init"], tolH=TOLH[prec])
    if with_fix:
        _check_eval(c, o, sc["poses_init"], include_fix=True, tolH=TOLH[prec])
    r = c.residual(sc["poses_init"])
    ro = o.residual(sc["poses_init"])
    assert abs(r - ro) <= 1e-12 * abs(ro)
    # at the ground truth as well: lambda_min sits at the noise floor (1e-4) of a covariance whose raw second
    # moments are O(10), so ANY fp64 evaluation (the oracle included) carries eps*|P/N|/lambda_min ~ 1e-11
    _check_eval(c, o, sc["poses_gt"], tolr=1e-10, tolH=TOLH[prec])


@pytest.mark.parametrize("drop", [0.0, 0.4])
def test_tensor_single_sweep_speculation(drop, monkeypatch):
    """Tensor path inside balm_damping_iter: from the second evaluation on, the column scales come from the previous
    evaluation and ONE observation sweep writes the digit planes; a device-side check either accepts that or re-runs
    the digit sweep with fresh scales. The public evaluate never speculates (pure function of its arguments)."""
    sc = scenes.make_scene(n_poses=24, n_planes=150, seed=33, drop=drop)
    # 4 iterations: all of them are genuine descent steps. (A 5th would sit at the converged cost, where r1 - r2 is pure
    # rounding and its SIGN -- the accept flag -- is a coin flip between two arithmetically different paths.)
    kw = dict(max_iter=4, force_hess=True, rel_tol=-1.0)

    def run():
        c = _ctx(sc, 1)
        c.reset_counters()
        H0, g0, r0 = c.evaluate(sc["poses_init"])
        H1, g1, r1 = c.evaluate(sc["poses_init"])
        tm = c.timings()
        assert tm["single_sweeps"] == 0 and tm["redone_sweeps"] == 0
        assert np.array_equal(H0, H1) and np.array_equal(g0, g1) and r0 == r1
        c.reset_counters()
        poses, tr, _ = c.damping_iter(sc["poses_init"], **kw)
        return poses, [(t["r1"], t["r2"], t["accepted"]) for t in tr], c.timings()

    monkeypatch.setenv("BALM_NO_SPEC", "1")
    p_two, t_two, tm = run()                                   # every evaluation sweeps twice
    assert tm["n_eval"] == 4 and tm["single_sweeps"] == 0 and tm["redone_sweeps"] == 0
    monkeypatch.delenv("BALM_NO_SPEC")
    p_one, t_one, tm = run()
    assert tm["n_eval"] == 4 and tm["single_sweeps"] + tm["redone_sweeps"] == 3 and tm["single_sweeps"] >= 1
    assert [x[2] for x in t_one] == [x[2] for x in t_two]
    rot, tra = _pose_err(p_one, p_two)
    assert rot <= 1e-8 and tra <= 1e-8                         # scales differ by powers of two at most -> rounding only
    for skew in (4, -9):   # adopted scales 16x too large (digits overflow) / 512x too small (precision lost)
        monkeypatch.setenv("BALM_TC_SPEC_SKEW", str(skew))
        p_bad, t_bad, tm = run()
        assert tm["single_sweeps"] == 0 and tm["redone_sweeps"] == 3, (skew, tm)
        # every speculation rejected -> the digit sweep re-ran with fresh scales -> the two-sweep result (the
        # accumulators come from a different instantiation of the sweep kernel, hence "to rounding", not "same bits")
        assert np.abs(p_bad - p_two).max() <= 1e-12
        assert [x[2] for x in t_bad] == [x[2] for x in t_two]
        assert max(abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(t_bad, t_two)) <= 1e-12


@pytest.mark.parametrize("prec", PRECS)
def test_reregistration_same_shape_reuses_buffers(prec):
    """A second window of the same shape keeps the device buffers (no free/malloc) -- results must be those of a
    fresh context, and nothing speculative may leak from the previous problem."""
    import balm_b200
    from balm_b200 import _lib
    scA = scenes.make_scene(n_poses=16, n_planes=90, seed=41)
    scB = scenes.make_scene(n_poses=16, n_planes=90, seed=42)
    c = _ctx(scA, prec)
    c.evaluate(scA["poses_init"])
    c.damping_iter(scA["poses_init"], max_iter=3)
    c.set_voxels(scB["row_ptr"], scB["pose_idx"], scB["obs10"], scB["coe"], scB["fix10"])
    c.reset_counters()
    o = _oracle(scB)
    H, g, r = _check_eval(c, o, scB["poses_init"], tolH=TOLH[prec])
    assert c.timings()["single_sweeps"] == 0                      # first evaluation of a new problem sweeps twice
    fresh = _ctx(scB, prec)
    H2, g2, r2 = fresh.evaluate(scB["poses_init"])
    assert np.array_equal(H, H2) and np.array_equal(g, g2) and r == r2
    pa, ta, _ = c.damping_iter(scB["poses_init"], max_iter=4)
    pb, tb, _ = fresh.damping_iter(scB["poses_init"], max_iter=4)
    assert [t["accepted"] for t in ta] == [t["accepted"] for t in tb]
    assert np.abs(pa - pb).max() <= 1e-9
    # a malformed registration leaves the context empty, not half-registered
    bad = scB["pose_idx"].copy()
    bad[1] = bad[0]
    with pytest.raises(_lib.BalmError):
        c.set_voxels(scB["row_ptr"], bad, scB["obs10"], scB["coe"], scB["fix10"])
    with pytest.raises(_lib.BalmError):
        c.evaluate(scB["poses_init"])
    c.set_voxels(scB["row_ptr"], scB["pose_idx"], scB["obs10"], scB["coe"], scB["fix10"])
    H3, g3, r3 = c.evaluate(scB["poses_init"])
    assert np.array_equal(H3, H2)


def test_phase_pipelining_gives_identical_results(monkeypatch):
    """The LM loop enqueues evaluation, solve, update and trial residual back to back (one host wait per iteration);
    BALM_SYNC_PHASES=1 waits after every phase. Same kernels, same order -> bit-identical poses and trace."""
    sc = scenes.make_scene(n_poses=20, n_planes=120, seed=44, drop=0.3)
    res = []
    for flag in (None, "1"):
        if flag:
            monkeypatch.setenv("BALM_SYNC_PHASES", flag)
        c = _ctx(sc, 1)
        poses, tr, _ = c.damping_iter(sc["poses_init"], max_iter=5)
        tm = c.timings()
        assert tm["n_eval"] >= 1 and tm["n_solve"] == len(tr) and tm["n_residual"] == len(tr)
        assert tm["ms_solve"] > 0 and tm["ms_syrk"] > 0
        res.append((poses, [(t["r1"], t["r2"], t["q1"], t["accepted"]) for t in tr]))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


def test_voxel_range_semantics():
    sc = scenes.make_scene(n_poses=10, n_planes=60, seed=22)
    c, o = _ctx(sc), _oracle(sc)
    Ha, ga, ra = _check_eval(c, o, sc["poses_init"], head=0, end=17)
    Hb, gb, rb = _check_eval(c, o, sc["poses_init"], head=17, end=60)
    H, g, r = c.evaluate(sc["poses_init"])
    assert np.abs(Ha + Hb - H).max() <= 1e-12 * np.abs(H).max()
    assert abs(ra + rb - r) <= 1e-13 * abs(r)
    H0, g0, r0 = c.evaluate(sc["poses_init"], 5, 5)  # empty range
    assert r0 == 0 and not H0.any() and not g0.any()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("drop", [0.0, 0.4])   # 0.4: sparse co-visibility in several voxel batches (pose-major lists cut per batch)
def test_batched_evaluation_equals_single_batch(monkeypatch, prec, drop):
    sc = scenes.make_scene(n_poses=16, n_planes=300, seed=23, drop=drop)
    c1 = _ctx(sc, prec)
    H1, g1, r1 = c1.evaluate(sc["poses_init"])
    monkeypatch.setenv("BALM_G_BUDGET_MB", "1")  # 1 MiB of G' -> many voxel batches
    c2 = _ctx(sc, prec)
    H2, g2, r2 = c2.evaluate(sc["poses_init"])
    assert np.abs(H1 - H2).max() <= (1e-12 if prec == 0 else 2e-8) * np.abs(H1).max()
    assert np.abs(g1 - g2).max() <= 1e-12 * np.abs(g1).max()
    assert abs(r1 - r2) <= 1e-13 * abs(r1)


@pytest.mark.parametrize("prec", PRECS)
def test_run_to_run_deterministic(prec):
    sc = scenes.make_scene(n_poses=24, n_planes=200, seed=24)
    c = _ctx(sc, prec)
    H1, g1, r1 = c.evaluate(sc["poses_init"])
    H2, g2, r2 = c.evaluate(sc["poses_init"])
    assert np.array_equal(H1, H2) and np.array_equal(g1, g2) and r1 == r2


def test_solve_matches_oracle_ldlt():
    sc = scenes.make_scene(n_poses=30, n_planes=120, seed=25)
    c, o = _ctx(sc), _oracle(sc)
    H, g, r = c.evaluate(sc["poses_init"])
    for u in (0.01, 0.1, 3.0):
        dx, q1, bad = c.solve(u)
        dxo, trial, q1o = o.lm_step(H, g, u, sc["poses_init"])
        assert not bad
        assert np.abs(dx - dxo).max() <= 1e-9 * max(1e-3, np.abs(dxo).max())
        assert abs(q1 - q1o) <= 1e-9 * abs(q1o)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_poses,n_planes,drop", [(8, 60, 0.0), (50, 400, 0.0), (20, 300, 0.3)])
def test_damping_iter_matches_oracle_per_iteration(n_poses, n_planes, drop, prec):
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=26, drop=drop)
    c, o = _ctx(sc, prec), _oracle(sc)
    poses, tr, per = c.damping_iter(sc["poses_init"], want_per_iter=True, gauge_mode=2)
    st, poses_o, tr_o, per_o = o.damping_iter(sc["poses_init"], gauge_mode=2)
    assert st == 0 and len(tr) == len(tr_o)
    assert [t["accepted"] for t in tr] == [t["accepted"] for t in tr_o]
    for it in range(len(tr)):
        rot, tra = _pose_err(per[it], per_o[it])
        assert rot <= 1e-6 and tra <= 1e-6, (it, rot, tra)   # north_star: 1e-6 rad / 1e-6 m per iteration
        assert abs(tr[it]["r2"] - tr_o[it]["r2"]) <= (1e-9 if prec == 0 else 1e-7) * abs(tr_o[it]["r2"])
        assert abs(tr[it]["u"] - tr_o[it]["u"]) <= 1e-5 * tr_o[it]["u"]
    # final gauge step (bavoxel.hpp:1159-1164) and the benchmark_virtual variant
    p0, _, _ = c.damping_iter(sc["poses_init"], gauge_mode=0)
    _, p0o, _, _ = o.damping_iter(sc["poses_init"], gauge_mode=0)
    assert max(_pose_err(p0, p0o)) <= 1e-6
    p1, _, _ = c.damping_iter(sc["poses_init"], max_iter=20, u0=0.1, gauge_mode=1, hess_includes_fix=True)
    assert np.array_equal(p1[0], np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]))
    rot, tra = orc.rmse(p1, sc["poses_gt"])
    assert rot * 57.3 < 0.1 and tra < 0.01  # converged to the noise floor (benchmark_virtual.cpp:517-518)


def test_too_few_planes_is_reported():
    import balm_b200
    sc = scenes.make_scene(n_poses=4, n_planes=10, seed=27)
    c = _ctx(sc)
    with pytest.raises(balm_b200.BalmError) as e:
        c.damping_iter(sc["poses_init"])
    assert e.value.status == 4  # reference: printf + exit(0) (bavoxel.hpp:1079-1085)


def test_reference_call_surface_mirror():
    """VOX_HESS / BALM2 with the reference's names and argument meaning (bavoxel.hpp:30-51,304-470,1025-1166)."""
    import balm_b200
    from balm_b200 import bavoxel
    sc = scenes.make_scene(n_poses=8, n_planes=50, seed=28, drop=0.2)
    bavoxel.win_size = 8
    vox = balm_b200.VOX_HESS()
    for a in range(len(sc["row_ptr"]) - 1):
        vec = [balm_b200.PointCluster() for _ in range(8)]
        for s in range(sc["row_ptr"][a], sc["row_ptr"][a + 1]):
            o10 = sc["obs10"][s]
            pc = vec[sc["pose_idx"][s]]
            pc.P = np.array([[o10[0], o10[1], o10[2]], [o10[1], o10[3], o10[4]], [o10[2], o10[4], o10[5]]])
            pc.v = o10[6:9].copy()
            pc.N = int(o10[9])
        vox.push_voxel(vec, balm_b200.PointCluster(), 0.0, 0)
    xs = [balm_b200.IMUST(*scenes.unpack_pose(p)) for p in sc["poses_init"]]
    opt = balm_b200.BALM2()
    r, H, g = opt.divide_thread_left(xs, vox)
    o = _oracle(sc)
    Ho, go, ro = o.evaluate_threads(sc["poses_init"], threads=4)
    assert abs(r - ro) <= 1e-12 * abs(ro) and np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max()
    assert abs(opt.only_residual(xs, vox) - o.residual(sc["poses_init"])) <= 1e-12 * abs(ro)
    opt.damping_iter(xs, vox, verbose=False)
    _, po, _, _ = o.damping_iter(sc["poses_init"])
    assert max(_pose_err(bavoxel.pack_poses(xs), po)) <= 1e-6


@pytest.mark.parametrize("prec", PRECS)
def test_synth_scene_roundtrip_and_parity_c1(prec):
    """BASELINE config C1 (50 poses, 2k plane voxels) generated in HBM, downloaded, checked against the oracle."""
    import balm_b200
    c = balm_b200.Context(50, 0, prec)
    gt, init = c.synth_virtual(2000, seed=10)
    row_ptr, pose_idx, obs10, coe = c.download_voxels()
    assert row_ptr[-1] == 100000 and np.all(coe == 50 * 40) and np.all(obs10[:, 9] == 40)
    o = orc.Oracle(50, row_ptr, pose_idx, obs10, coe)
    H, g, r = c.evaluate(init)
    Ho, go, ro = o.evaluate_threads(init, threads=4)
    assert abs(r - ro) <= 1e-12 * abs(ro)
    assert np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
    assert np.abs(H - Ho).max() <= TOLH[prec] * np.abs(Ho).max()
    # the planes are planes: at ground truth the cost sits at the point-noise floor  sum coe * sigma^2
    assert 0.8 < c.residual(gt) / (coe.sum() * 1e-4) < 1.2
    # host-buffer path (balm_set_voxels) reproduces the HBM-resident one bit for bit
    c2 = balm_b200.Context(50, 0, prec)
    c2.set_voxels(row_ptr, pose_idx, obs10, coe)
    H2, g2, r2 = c2.evaluate(init)
    assert np.array_equal(H, H2) and np.array_equal(g, g2) and r == r2


@pytest.mark.parametrize("prec", PRECS)
def test_full_size_properties_c3(prec):
    """BASELINE config C3 (500 poses, 100k voxels): too big for the oracle, so size-independent properties:
    gauge null space (Supplementary eq. 170-185), symmetry, residual consistency, LM monotone decrease."""
    import balm_b200
    N, M = 500, 100000
    c = balm_b200.Context(N, 0, prec)
    gt, init = c.synth_virtual(M, seed=10)
    H, g, r = c.evaluate(init)
    assert np.array_equal(H, H.T)
    assert abs(r - c.residual(init)) <= 1e-12 * abs(r)
    rng = np.random.default_rng(0)
    dT = np.tile(rng.normal(size=6), N)
    assert abs(g @ dT) <= 1e-8 * np.abs(g).max() * np.linalg.norm(dT)
    assert abs(dT @ H @ dT) <= (1e-8 if prec == 0 else 1e-7) * np.abs(H).max() * (dT @ dT)
    # linearity in the voxel set: two halves add up
    Ha, ga, ra = c.evaluate(init, 0, M // 2)
    Hb, gb, rb = c.evaluate(init, M // 2, M)
    assert np.abs(Ha + Hb - H).max() <= (1e-11 if prec == 0 else 2e-9) * np.abs(H).max()
    poses, tr, _ = c.damping_iter(init, max_iter=6)
    costs = [t["r2"] for t in tr if t["accepted"]]
    assert len(costs) >= 3 and all(b <= a for a, b in zip(costs, costs[1:]))
    rot, tra = orc.rmse(poses, _gauge(gt))
    assert rot * 57.3 < 0.02 and tra < 0.002


def _gauge(p):
    R0, p0 = scenes.unpack_pose(p[0])
    return scenes.pack_poses([R0.T @ scenes.unpack_pose(x)[0] for x in p], [R0.T @ (x[9:] - p0) for x in p])


def test_tensor_path_matches_fp64_path_c2():
    """BASELINE config C2 shape (200 poses, 20k voxels): tcgen05 split-integer SYRK vs the fp64 DMMA path on the
    same device data, and the pose update they imply (north_star: 1e-6 rad / 1e-6 m per iteration)."""
    import balm_b200
    N, M = 200, 20000
    out = {}
    for prec in (0, 1):
        c = balm_b200.Context(N, 0, prec)
        gt, init = c.synth_virtual(M, seed=11)
        H, g, r = c.evaluate(init)
        dx, q1, bad = c.solve(0.01)
        poses, tr, per = c.damping_iter(init, want_per_iter=True, gauge_mode=2)
        out[prec] = (H, g, r, dx, per, tr)
        c.close()
    H0, g0, r0, dx0, per0, tr0 = out[0]
    H1, g1, r1, dx1, per1, tr1 = out[1]
    assert np.array_equal(g0, g1) and r0 == r1           # the O(K) passes are shared
    assert np.abs(H1 - H0).max() <= 1e-8 * np.abs(H0).max()
    assert np.abs(dx1 - dx0).max() <= 1e-7
    assert [t["accepted"] for t in tr0] == [t["accepted"] for t in tr1]
    for a, b in zip(per0, per1):
        assert max(_pose_err(a, b)) <= 1e-6


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_poses,n_planes,drop", [
    (3, 2, 0.0),       # two voxels only
    (130, 40, 0.0),    # N > 128: several pose tiles, n = 780 -> 7 column blocks with padding
    (700, 12, 0.0),    # pose table larger than the 64 KB shared-memory staging of the stats kernel
    (65, 33, 0.6),     # heavily ragged
])
def test_evaluate_edge_shapes(n_poses, n_planes, drop, prec):
    sc = scenes.make_scene(n_poses=n_poses, n_planes=n_planes, seed=41, drop=drop, pts_size=8)
    c, o = _ctx(sc, prec), _oracle(sc)
    _check_eval(c, o, sc["poses_init"], tolH=TOLH[prec])
    r = c.residual(sc["poses_init"])
    assert abs(r - o.residual(sc["poses_init"])) <= 1e-12 * abs(r)


def test_tensor_digit_planes_option(monkeypatch):
    """BALM_TC_SLICES=3 (22-bit fixed point) still meets the pose contract; 4 planes (default) is tighter."""
    sc = scenes.make_scene(n_poses=24, n_planes=300, seed=42)
    o = _oracle(sc)
    Ho, go, ro = o.evaluate(sc["poses_init"])
    errs = {}
    for S in (3, 4):
        monkeypatch.setenv("BALM_TC_SLICES", str(S))
        c = _ctx(sc, 1)
        H, g, r = c.evaluate(sc["poses_init"])
        errs[S] = np.abs(H - Ho).max() / np.abs(Ho).max()
    assert errs[4] <= 1e-8 and errs[3] <= 5e-6 and errs[4] < errs[3]


def test_evaluate_is_independent_of_history():
    """The tensor path derives its column scales from the current poses only (two sweeps), so the same inputs give
    the same bits whatever was evaluated before."""
    sc = scenes.make_scene(n_poses=16, n_planes=120, seed=43)
    c = _ctx(sc, 1)
    H1, g1, r1 = c.evaluate(sc["poses_init"])
    c.evaluate(sc["poses_gt"])
    H2, g2, r2 = c.evaluate(sc["poses_init"])
    assert np.array_equal(H1, H2) and np.array_equal(g1, g2) and r1 == r2


def test_two_gpu_shards_match_single_gpu():
    """Voxel shards on two GPUs + NCCL all-reduce of [H|g|r] == one GPU with all voxels (needs >= 2 GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "scripts", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("prec", PRECS)
def test_c2_shape_per_iteration_parity_vs_oracle(prec):
    """BASELINE config C2 shape (200 poses; voxel count scaled to what the CPU oracle finishes in seconds):
    per-iteration pose parity with the oracle, 1e-6 rad / 1e-6 m, same accept/reject sequence."""
    import balm_b200
    N, M = 200, 600
    c = balm_b200.Context(N, 0, prec)
    gt, init = c.synth_virtual(M, seed=12)
    row_ptr, pose_idx, obs10, coe = c.download_voxels()
    o = orc.Oracle(N, row_ptr, pose_idx, obs10, coe)
    poses, tr, per = c.damping_iter(init, max_iter=4, want_per_iter=True, gauge_mode=2)
    st, poses_o, tr_o, per_o = o.damping_iter(init, max_iter=4, gauge_mode=2)
    assert st == 0 and [t["accepted"] for t in tr] == [t["accepted"] for t in tr_o]
    for it in range(len(tr)):
        rot, tra = _pose_err(per[it], per_o[it])
        assert rot <= 1e-6 and tra <= 1e-6, (it, rot, tra)
