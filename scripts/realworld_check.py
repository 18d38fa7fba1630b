"""Dev: tensor-path accuracy on the real-data golden fixture (GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import balm_b200
d = np.load(os.path.join(ROOT, "tests", "golden", "realworld_voxels.npz"))
N = int(d["n_poses"])
def run(prec, S=None):
    if S: os.environ["BALM_TC_SLICES"] = str(S)
    c = balm_b200.Context(N, 0, prec)
    c.set_voxels(d["row_ptr"], d["pose_idx"], d["obs10"], d["coe"])
    H, g, r = c.evaluate(d["poses_init"])
    dx, q1, bad = c.solve(0.01)
    return H, g, dx
H0, g0, dx0 = run(0)
dg = np.sqrt(np.abs(np.diag(H0)))
print("diag(H) range", np.abs(np.diag(H0)).min(), np.abs(np.diag(H0)).max(), "|dx|", np.abs(dx0).max())
for S in (4, 3):
    H, g, dx = run(1, S)
    E = H - H0
    print(f"S={S}: max|dH|/max|H| {np.abs(E).max()/np.abs(H0).max():.2e}  max |dH_ij|/sqrt(Hii Hjj) {np.abs(E/np.outer(dg,dg)).max():.2e} "
          f" diag rel {np.abs(np.diag(E)/np.diag(H0)).max():.2e}  |ddx| {np.abs(dx-dx0).max():.2e}")
    i = np.argmax(np.abs(np.diag(E)/np.diag(H0))); print("   worst diag idx", i, np.diag(H0)[i], np.diag(E)[i])
