"""Dev: time the evaluation phases at C3 for the tensor path (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200
N, M = 500, 100000
c = balm_b200.Context(N, 0, 1)
gt, init = c.synth_virtual(M, seed=10)
c.evaluate(init, want_H=False)
c.reset_counters()
for _ in range(3):
    H, g, r = c.evaluate(init, want_H=False)
tm = c.timings()
print(os.environ.get("TAG", ""), {k: round(v / 3, 3) for k, v in tm.items() if k.startswith("ms_") and v})
dx, q1, bad = c.solve(0.01)
print("dx max", np.abs(dx).max(), "q1", q1, "bad", bad, "sum|dx|", np.abs(dx).sum())
