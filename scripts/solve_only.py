import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200
N, M = 500, 20000
c = balm_b200.Context(N, 0, 1)
gt, init = c.synth_virtual(M, seed=10)
c.evaluate(init, want_H=False)
for _ in range(3):
    dx, q1, bad = c.solve(0.01)
c.reset_counters()
for _ in range(5):
    dx, q1, bad = c.solve(0.01)
print({k: v / 5 for k, v in c.timings().items() if k == "ms_solve"})
