#!/usr/bin/env python
"""bench.py -- BA iterations/sec of the BALM 2.0 hot path on B200 (BASELINE.json metric).

A "step" is one Levenberg-Marquardt iteration of BALM2::damping_iter (bavoxel.hpp:1104-1157):
  1 Hessian/gradient evaluation (left_evaluate_acc2) + 1 damped LDL^T solve + SE(3) left update
  + 1 residual-only evaluation, with the Hessian re-evaluated every iteration (force_hess) so that every step
  is the same amount of work whether the trial step is accepted or not.
Workload (N=1): BASELINE config C3 -- 500 poses x 100 000 plane voxels, synthetic plane features of the
benchmark_virtual shape (benchmark_virtual.cpp:547-606), every pose sees every plane, 40 points/observation.
With --gpus G each rank holds its own 100k-voxel shard of one G*100k-voxel scene (weak scaling, BASELINE
config C4 shape); the library all-reduces [H|g|r] with NCCL once per evaluation.

  value      : shard-iterations per second with the voxels already resident in HBM (G * K / t)
  e2e        : the same through the host-buffer call (balm_set_voxels from pinned host arrays + damping_iter +
               poses back), copies inside the timed region
  roofline   : dominant kernel = the rank-3M symmetric update (SYRK, bavoxel.hpp:404-418 restated)
  cpu_baseline : the CPU oracle (port of the reference loop nest; Eigen/PCL/ROS are absent so the reference
               itself cannot be built here) timed on a bounded voxel sample of the same workload

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--poses P --voxels M --precision fp64|tensor]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PTS = 40
NOISE = 0.01
RANGE = 2.0
SEED = 10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(n_poses, voxels_total, sample_voxels, threads, obs=None):
    """Time the oracle on `sample_voxels` voxels of the workload and extrapolate the O(M) passes linearly in M
    (accumulation and residual cost are exactly linear in M); the n x n LDL^T solve is
    timed at full size. Returns (seconds per LM iteration, detail dict)."""
    from oracle import oracle_py as orc
    import balm_b200
    if obs is None:
        c = balm_b200.Context(n_poses, 0, 0)
        gt, init = c.synth_virtual(sample_voxels, 0, PTS, NOISE, RANGE, SEED)
        row_ptr, pose_idx, obs10, coe = c.download_voxels()
        c.close()
    else:
        row_ptr, pose_idx, obs10, coe, init = obs
    o = orc.Oracle(n_poses, row_ptr, pose_idx, obs10, coe)
    t0 = time.perf_counter()
    H, g, r = o.evaluate_threads(init, threads=threads)  # divide_thread_left, 4 std::threads in the reference
    t_eval = time.perf_counter() - t0
    t0 = time.perf_counter()
    dx, trial, q1 = o.lm_step(H, g, 0.01, init)          # single-thread pivoted LDL^T + update (Eigen: 1 thread)
    t_solve = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.residual(trial)                                    # single thread, as the reference
    t_res = time.perf_counter() - t0
    scale = voxels_total / sample_voxels
    t_iter = t_eval * scale + t_solve + t_res * scale
    return t_iter, dict(t_eval_sample=t_eval, t_solve=t_solve, t_residual_sample=t_res, scale=scale)


def numpy_sample(n_poses, sample_voxels):
    """benchmark_virtual-shaped sample scene generated with numpy (tests/scenes.py) -- keeps the reference arm free
    of any balm_b200 code."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    sc = scenes.make_scene(n_poses=n_poses, n_planes=sample_voxels, pts_size=PTS, point_noise=NOISE, surf_range=RANGE,
                           seed=SEED)
    return sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], sc["poses_init"]


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path. The reference needs Eigen, PCL and
    ROS, none of which exist in this image (no baseline/_ref, no oracle/_ref), so the arm times the oracle port
    of the reference loop nest (4 threads for the accumulation, 1 for residual and LDL^T, like the reference).
    Each step = one LM iteration on a bounded voxel sample, extrapolated linearly in M. No balm_b200 code runs."""
    if rank != 0:
        return
    n, m = args.poses, args.voxels
    sample = min(args.cpu_sample_voxels, 128)
    data = numpy_sample(n, sample)
    times = []
    for i in range(args.warmup + args.steps):
        t, detail = cpu_sample(n, m, sample, 4, data)
        if i >= args.warmup:
            times.append(t)
    t_iter = float(np.mean(times))
    val = 1.0 / t_iter
    line = {
        "impl": "reference", "metric": "ba_iterations_per_sec", "value": val, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_iter, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE C3: BA LM iteration, {n} poses x {m} plane voxels (benchmark_virtual shape, "
                               f"{PTS} pts/obs), CPU oracle port of the reference loop nest", "poses": n, "voxels": m},
        "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 4, "kind": "port",
                         "sample": f"{sample} of {m} voxels at {n} poses, O(M) passes scaled x{m / sample:.0f}; "
                                   f"LDL^T n={6 * n} at full size; accumulation 4 threads, rest 1 thread",
                         "detail": detail},
        "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="balm_b200")
    ap.add_argument("--poses", type=int, default=500)
    ap.add_argument("--voxels", type=int, default=100000, help="plane voxels per GPU")
    ap.add_argument("--precision", default=os.environ.get("BALM_BENCH_PRECISION", "tensor"), choices=["fp64", "tensor"])
    ap.add_argument("--cpu-sample-voxels", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import balm_b200

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, M = args.poses, args.voxels
    prec = balm_b200.PREC_TENSOR if args.precision == "tensor" else balm_b200.PREC_FP64
    ctx = balm_b200.Context(N, local_rank, prec)
    gt, init = ctx.synth_virtual(M, rank * M, PTS, NOISE, RANGE, SEED)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(balm_b200.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))

    lm = dict(u0=0.01, v0=2.0, rel_tol=-1.0, gauge_mode=2, min_planes_per_pose=0, force_hess=True)

    # ---- HBM-resident timing ----
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs ~100 ms to produce its first sample: started before the warm-up steps
    ctx.damping_iter(init, max_iter=args.warmup, **lm)
    ctx.reset_counters()
    barrier()
    ctx.timer_begin()
    poses, trace, _ = ctx.damping_iter(init, max_iter=args.steps, **lm)
    ms = ctx.timer_end()
    barrier()
    clocks = sampler.stop()
    tm = ctx.timings()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    K = args.steps
    value = world * K / (ms_max * 1e-3)

    # ---- end-to-end through the host-buffer call ----
    e2e = None
    if not args.no_e2e:
        row_ptr, pose_idx, obs10, coe = ctx.download_voxels()

        def pinned(a):
            tt = torch.from_numpy(a).pin_memory()
            return tt.numpy(), tt
        keep = []
        arrs = []
        for a in (row_ptr, pose_idx, obs10, coe):
            n_, t_ = pinned(a)
            arrs.append(n_)
            keep.append(t_)
        del row_ptr, pose_idx, obs10, coe
        for rep in range(2):  # first repetition warms the allocator / page tables
            barrier()
            t0 = time.perf_counter()
            ctx.set_voxels(arrs[0], arrs[1], arrs[2], arrs[3])
            t_set = time.perf_counter() - t0
            p2, tr2, _ = ctx.damping_iter(init, max_iter=K, **lm)
            ctx.sync()
            t1 = time.perf_counter()
        te = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = sum(a.nbytes for a in arrs) + init.nbytes
        e2e = {"value": world * K / float(te.item()), "unit": "iter/s", "h2d_bytes_per_step": int(h2d / K),
               "d2h_bytes_per_step": int(init.nbytes / K + 8 * 3),
               "set_voxels_ms": 1e3 * t_set,
               "note": "balm_set_voxels(pinned host CSR arrays) + damping_iter(K) + poses back, per call"}
        assert np.abs(p2 - poses).max() < 1e-9  # same answer through the host path

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n_eval = max(tm["n_eval"], 1)
    syrk_ms = tm["ms_syrk"] / n_eval
    k = N
    flops = 108.0 * k * (k + 1) * M  # SYRK-half count per evaluation
    obs_bytes = 80.0 * M * N
    if args.precision == "tensor":
        peak_tf, peak_note = pk["bf16_sus"], f"bf16 dense sustained, {pk['src']} (int8 tcgen05 pipe is 2x bf16)"
    else:
        peak_tf, peak_note = pk["bf16_sus"], (f"bf16 dense sustained, {pk['src']}; the fp64 path runs on the DMMA "
                                              f"pipe whose nominal peak is 40 TFLOP/s (not in MEASURED_PEAKS.json)")
    ach_tf = flops / (syrk_ms * 1e-3) / 1e12
    traffic = None  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_tensor_ncu_full_summary.json")))
        for k_ in prof:
            if args.precision == "tensor" and k_["kernel"].startswith("syrk_tc") and N == 500 and M == 100000:
                def gb(x):
                    v, u = x.split()
                    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                traffic = gb(k_["dram__bytes_read.sum"]) + gb(k_["dram__bytes_write.sum"])
    except Exception:
        traffic = None
    roof = {"kernel": "syrk_f64_kernel" if args.precision == "fp64" else "syrk_tc_2sm_kernel", "bound": "tensor",
            "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
            "peak_note": peak_note, "algorithmic_flops_per_launch": flops, "ms_per_launch": syrk_ms,
            "traffic_note": "dram__bytes_read+write per launch from profiles/r1_tensor_ncu_full_summary.json; algorithmic "
                            "bytes of this kernel = digit planes read once (3 x 3M x ldg B) + fp64 partial tiles written"}
    if args.precision == "tensor":
        planes = tm["digit_planes"]
        pairs = planes * (planes + 1) // 2
        nbk = (6 * N + 127) // 128
        tiles_exec = sum(((bj + 2) // 2) * 2 for bj in range(nbk))  # 2-SM pairs: odd columns carry one redundant tile
        int8_ops = 2.0 * pairs * tiles_exec * 128 * 128 * 3 * M
        roof["digit_planes"] = planes
        roof["executed_int8_tops"] = int8_ops / (syrk_ms * 1e-3) / 1e12
        roof["frac_of_int8_nominal_4500_tops"] = roof["executed_int8_tops"] / 4500.0
        roof["frac_of_2x_measured_bf16_burst"] = roof["executed_int8_tops"] / (2.0 * pk["bf16"])
        roof["executed_note"] = ("int8 digit-pair MMAs actually issued (S(S+1)/2 products per tile incl. the redundant "
                                 "below-diagonal tiles of the 2-SM pairing); the int8 tcgen05 pipe peaks at 2x the bf16 "
                                 "pipe, so 2x the measured cuBLAS bf16 burst is the comparable measured ceiling")
    if args.precision == "fp64":
        roof["frac_of_fp64_nominal_40tf"] = ach_tf / 40.0
    res_ms = tm["ms_residual"] / max(tm["n_residual"], 1)
    stats_ms = tm["ms_stats"] / max(tm["n_stats"], 1) if tm["n_stats"] else res_ms
    obs_ms = tm["ms_obs"] / n_eval
    if args.precision == "tensor":
        gout = 18.0 * tm["digit_planes"]
        gnote = (f"80 B/obs read (counted once) + {gout:.0f} B/obs written (18 values x {tm['digit_planes']} int8 digit "
                 f"planes); {tm['single_sweeps']} of {n_eval} evaluations needed one sweep (scales speculated from the "
                 f"previous evaluation and accepted), {tm['redone_sweeps']} re-ran the digit sweep, the rest swept twice")
    else:
        gout, gnote = 144.0, "80 B/obs read + 144 B/obs fp64 G' written"
    hbm = {
        "residual_pass": {"bound": "hbm", "achieved": obs_bytes / (res_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                          "unit": "GB/s", "frac": obs_bytes / (res_ms * 1e-3) / 1e9 / pk["hbm"], "ms": res_ms},
        "voxel_stats": {"bound": "hbm", "achieved": obs_bytes / (stats_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                        "unit": "GB/s", "frac": obs_bytes / (stats_ms * 1e-3) / 1e9 / pk["hbm"], "ms": stats_ms,
                        "note": f"ran in {tm['n_stats']} of {n_eval} evaluations; the others took the eigen data "
                                f"from the residual pass of the accepted step (same kernel, same poses)"},
        "obs_pass": {"bound": "hbm", "achieved": (obs_bytes + gout * M * N) / (obs_ms * 1e-3) / 1e9,
                     "peak": pk["hbm"], "unit": "GB/s",
                     "frac": (obs_bytes + gout * M * N) / (obs_ms * 1e-3) / 1e9 / pk["hbm"], "ms": obs_ms,
                     "bytes_note": gnote},
    }
    phases = {k_: (tm[k_] / n_eval if k_ not in ("ms_solve", "ms_residual") else
                   tm[k_] / max(tm["n_solve" if k_ == "ms_solve" else "n_residual"], 1))
              for k_ in ("ms_stats", "ms_obs", "ms_slice", "ms_syrk", "ms_assemble", "ms_allreduce", "ms_solve",
                         "ms_residual")}

    cpu = None
    if world == 1 and not args.no_cpu:
        t_iter, detail = cpu_sample(N, M, args.cpu_sample_voxels, 4)
        cpu = {"value": 1.0 / t_iter, "unit": "iter/s", "cores": 4, "kind": "port",
               "sample": f"{args.cpu_sample_voxels} of {M} voxels at {N} poses, O(M) passes scaled "
                         f"x{M / args.cpu_sample_voxels:.0f}; LDL^T n={6 * N} timed at full size; accumulation "
                         f"4 threads (bavoxel.hpp:1027), residual + LDL^T 1 thread; gcc -O3 without -march=native "
                         f"(CMakeLists.txt:9)", "detail": detail}

    line = {
        "metric": "ba_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "s8 split-integer -> f64",
        "data": "synthetic",
        "config": {"workload": f"BASELINE C3 per GPU: BA LM iteration, {N} poses x {M} plane voxels/GPU "
                               f"({world * M} total), benchmark_virtual shape, {PTS} pts/obs, every pose sees every "
                               f"plane; value counts one iteration of each rank's shard",
                   "poses": N, "voxels_per_gpu": M, "total_voxels": world * M, "precision": args.precision,
                   "parallelism": f"voxel-shard x{world}, NCCL all-reduce of [H|g|r] per evaluation",
                   "l2": "inputs (4.0 GB of observations per GPU) exceed the 126 MB L2; no explicit flush",
                   "lm": "force_hess=1, convergence exit disabled, 1 eval + 1 solve + 1 update + 1 residual per step"},
        "job_iter_per_s": K / (ms_max * 1e-3),
        "phases_ms": phases, "roofline": roof, "roofline_hbm": hbm, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": tm["launches"], "clocks": clocks,
        "sweeps": {"evaluations": tm["n_eval"], "stats_passes": tm["n_stats"], "single_sweeps": tm["single_sweeps"],
                   "redone_sweeps": tm["redone_sweeps"]},
        "lm_trace": [{"r1": t_["r1"], "r2": t_["r2"], "acc": t_["accepted"]} for t_ in trace[:4]],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
