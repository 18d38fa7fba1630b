"""Experiment: do the tcgen05 SYRK (tensor pipe) and the observation sweep (fp64 pipe) run faster side by side on the
same SMs than back to back? C3, one B200. Prints the five timings of balm_debug_overlap_probe."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200
from balm_b200 import _lib as L

N, M = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ctx = balm_b200.Context(N, 0, balm_b200.PREC_TENSOR)
gt, init = ctx.synth_virtual(M, 0, 40, 0.01, 2.0, 10)
lm = dict(u0=0.01, v0=2.0, rel_tol=-1.0, gauge_mode=2, min_planes_per_pose=0, force_hess=True)
ctx.damping_iter(init, max_iter=4, **lm)
out = (C.c_float * 5)()
lib = L.lib()
lib.balm_debug_overlap_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
L.check(lib.balm_debug_overlap_probe(ctx._h, 5, out))
names = ["syrk 2sm alone (168 regs)", "syrk 2sm alone (128 regs)", "sweep alone (128-thread CTAs)",
         "sweep alone (96-thread CTAs)", "syrk(128 regs) || sweep(96-thread CTAs)"]
for n, v in zip(names, out):
    print(f"{n:45s} {v:8.3f} ms")
print(f"back to back {out[0] + out[2]:.3f} ms -> side by side {out[4]:.3f} ms")
poses, tr, _ = ctx.damping_iter(init, max_iter=3, **lm)   # the context still works afterwards
print("after probe: r", [t["r2"] for t in tr])
