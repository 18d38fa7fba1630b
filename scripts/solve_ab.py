#!/usr/bin/env python
"""A/B of the damped normal-equations solve (kernel K6, `(Hess + u*D).ldlt().solve(-JacT)`, bavoxel.hpp:1113-1114):
the library's blocked LDL^T (balm_solve) against cuSOLVER on the same matrix and the same B200 --
  potrf + potrs (Cholesky; valid in the accepted-step regime where H + uD is positive definite),
  sytrf + sytrs (Bunch-Kaufman LDL^T, the pivoted factorisation closest to Eigen's LDLT),
both through torch.linalg (torch is plumbing here: it only forwards to cusolverDn<t>potrf/potrs/sytrf).
SURVEY.md section 7 step 6 names cuSOLVER as the bar to beat.  Usage:  python scripts/solve_ab.py [--reps 20]
Prints one JSON line per size: n, ms per solve of each, max |dx - dx_ref|.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--voxels", type=int, default=3000)
    args = ap.parse_args()
    torch.backends.cuda.preferred_linalg_library("cusolver")
    u = 0.01
    for N in (50, 200, 500):
        c = balm_b200.Context(N, 0, balm_b200.PREC_FP64)
        gt, init = c.synth_virtual(args.voxels, seed=10)
        H, g, r = c.evaluate(init)
        n = 6 * N
        for _ in range(3):
            c.solve(u)
        c.reset_counters()
        for _ in range(args.reps):
            dx, q1, bad = c.solve(u)
        tm = c.timings()
        ms_balm = tm["ms_solve"] / tm["n_solve"]
        A = torch.from_numpy(H + u * np.diag(np.diag(H))).cuda()
        b = torch.from_numpy(-g).cuda().reshape(n, 1)

        def chol():  # cholesky_ex: no device->host info check, so nothing but cusolverDnDpotrf + potrs is timed
            L, info = torch.linalg.cholesky_ex(A)
            return torch.cholesky_solve(b, L)

        def chol_factor_only():
            return torch.linalg.cholesky_ex(A)[0]

        def sytrf():
            LD, piv = torch.linalg.ldl_factor(A)
            return torch.linalg.ldl_solve(LD, piv, b)

        ms_potrf_only, _ = timed(chol_factor_only, args.reps)
        ms_chol, x_chol = timed(chol, args.reps)
        try:
            ms_sytrf, x_sy = timed(sytrf, args.reps)
            e_sy = float(np.abs(x_sy.cpu().numpy().ravel() - dx).max())
        except Exception as e:  # noqa: BLE001
            ms_sytrf, e_sy = None, str(e)[:80]
        xr = np.linalg.solve(H + u * np.diag(np.diag(H)), -g)
        print(json.dumps({"n": n, "poses": N, "balm_solve_ms": ms_balm, "cusolver_potrf_potrs_ms": ms_chol,
                          "cusolver_potrf_only_ms": ms_potrf_only, "cusolver_sytrf_sytrs_ms": ms_sytrf,
                          "balm_vs_numpy_max_abs": float(np.abs(dx - xr).max()),
                          "potrf_vs_numpy_max_abs": float(np.abs(x_chol.cpu().numpy().ravel() - xr).max()),
                          "sytrf_vs_balm_max_abs": e_sy, "dx_max": float(np.abs(xr).max()), "not_pd": bool(bad)}),
              flush=True)
        c.close()


if __name__ == "__main__":
    main()
