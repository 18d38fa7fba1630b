#!/usr/bin/env python
"""Timeline of the persistent tile-DAG factorisation (BALM_DAG_TRACE=1): per-step times of the chain CTA and the busy
span / task count of every worker CTA. Usage: BALM_DAG_TRACE=1 python scripts/dag_trace.py [N_POSES]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ["BALM_DAG_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200  # noqa: E402
from balm_b200 import _lib as L  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
c = balm_b200.Context(N, 0, 0)
gt, init = c.synth_virtual(1500, seed=10)
c.evaluate(init, want_H=False)
for _ in range(3):
    c.solve(0.01)
c.reset_counters()
c.solve(0.01)
print("ms_solve", c.timings()["ms_solve"])
nt = (6 * N + 63) // 64
buf = np.zeros(4 * nt + 8 * 1024, dtype=np.uint64)
L.lib().balm_debug_dag_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.check(L.lib().balm_debug_dag_trace(c._h, buf.ctypes.data_as(C.c_void_p), len(buf)))
ch = buf[:4 * nt].reshape(nt, 4).astype(np.int64)
t0 = ch[0, 0]
print("chain: total %.1f us" % ((ch[-1, 3] - t0) / 1e3))
print(" step  start   wait_panel  panel+wait_diag  factor   (us)")
for k in range(nt):
    s, a, b, e = ch[k]
    if k == 0:
        print(f"{k:4d} {(s - t0) / 1e3:8.1f} {'':>10} {'':>12} {(e - s) / 1e3:8.1f}")
    else:
        print(f"{k:4d} {(s - t0) / 1e3:8.1f} {(a - s) / 1e3:10.1f} {(b - a) / 1e3:12.1f} {(e - b) / 1e3:8.1f}")
w = buf[4 * nt:].reshape(-1, 8).astype(np.int64)
st = buf[4 * nt:4 * nt + 8].astype(np.int64)
if nt > 20:
    print("step 20 fine (us): mini-panel+load %.1f | sub-panels %s | X assembly %.1f | write-back %.1f | release %.1f" % (
        (st[0] - ch[20, 0]) / 1e3, np.round(np.diff(st[0:5]) / 1e3, 1).tolist(), (st[5] - st[4]) / 1e3, (st[6] - st[5]) / 1e3,
        (ch[20, 3] - st[6]) / 1e3))
w = w[1:149]
act = w[w[:, 2] > 0]
print("workers: %d active, tasks/CTA min %d mean %.1f max %d; span mean %.1f us; us/task mean %.2f" % (
    len(act), act[:, 2].min(), act[:, 2].mean(), act[:, 2].max(), (act[:, 1] - act[:, 0]).mean() / 1e3,
    ((act[:, 1] - act[:, 0]) / act[:, 2]).mean() / 1e3))
print("near group (CTA 1-8): tasks", w[:8, 2].tolist(), "end us", ((w[:8, 1] - t0) / 1e3).round(1).tolist())
far = w[8:][w[8:, 2] > 0]
print("far group end us: min %.1f max %.1f" % (((far[:, 1] - t0) / 1e3).min(), ((far[:, 1] - t0) / 1e3).max()))
print("far CTA means (us): span %.0f | update tasks: wait %.0f load %.0f mma %.0f store+release %.0f | panel tasks %.0f" % (
    (far[:, 1] - far[:, 0]).mean() / 1e3, far[:, 3].mean() / 1e3, far[:, 4].mean() / 1e3, far[:, 5].mean() / 1e3,
    far[:, 6].mean() / 1e3, far[:, 7].mean() / 1e3))
