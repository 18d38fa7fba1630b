"""Host-side mirror of the reference's call surface for the BA hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as /root/reference/src/benchmark/bavoxel.hpp:
  PointCluster                       include/tools.hpp:290-349
  IMUST (R, p only)                  include/tools.hpp:141-201
  VOX_HESS.push_voxel                bavoxel.hpp:30-51
  VOX_HESS.left_evaluate_acc2        bavoxel.hpp:304-426
  VOX_HESS.evaluate_only_residual    bavoxel.hpp:428-470
  BALM2.divide_thread_left           bavoxel.hpp:1025-1059
  BALM2.only_residual                bavoxel.hpp:1061-1067
  BALM2.damping_iter                 bavoxel.hpp:1069-1166
The arithmetic runs in libbalm_b200.so (CUDA, sm_100a); this module only packs the reference's
pointer-of-vectors layout into the CSR arrays once per damping_iter (the association is static during BA,
benchmark_realworld.cpp:187-218).
"""
import sys

import numpy as np

from . import _lib as L
from .context import Context

win_size = 20  # the reference's mutable global (bavoxel.hpp:17); must equal len(x_stats)


class PointCluster:
    """P = sum p p^T, v = sum p, N  (include/tools.hpp:290-349)"""

    __slots__ = ("P", "v", "N")

    def __init__(self):
        self.P = np.zeros((3, 3))
        self.v = np.zeros(3)
        self.N = 0

    def clear(self):
        self.P[:] = 0
        self.v[:] = 0
        self.N = 0

    def push(self, vec):
        vec = np.asarray(vec, dtype=np.float64)
        self.N += 1
        self.P += np.outer(vec, vec)
        self.v += vec

    def cov(self):
        c = self.v / self.N
        return self.P / self.N - np.outer(c, c)

    def __iadd__(self, o):
        self.P += o.P
        self.v += o.v
        self.N += o.N
        return self

    def transform(self, sigv, R, p):
        self.N = sigv.N
        self.v = R @ sigv.v + self.N * p
        rp = np.outer(R @ sigv.v, p)
        self.P = R @ sigv.P @ R.T + rp + rp.T + self.N * np.outer(p, p)

    def pack10(self):
        P = self.P
        return [P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], self.v[0], self.v[1], self.v[2], float(self.N)]


class IMUST:
    """Pose state; BA uses only R and p (include/tools.hpp:141-201)."""

    __slots__ = ("R", "p")

    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64)
        self.p = np.zeros(3) if p is None else np.array(p, dtype=np.float64)


def pack_poses(xs):
    return np.stack([np.concatenate([x.R.T.reshape(9), x.p]) for x in xs])


def unpack_poses(arr, xs):
    for x, a in zip(xs, arr):
        x.R = a[:9].reshape(3, 3).T.copy()
        x.p = a[9:12].copy()


class VOX_HESS:
    """Factor container (bavoxel.hpp:21-482). Holds references to the per-voxel cluster vectors like the
    reference's non-owning pointers; the device copy is made lazily and reused until a voxel is pushed."""

    def __init__(self, device=0, precision=L.PREC_FP64):
        self.sig_vecs = []
        self.plvec_voxels = []
        self.coeffs = []
        self._device = device
        self._precision = precision
        self._ctx = None

    def push_voxel(self, vec_orig, fix, feat_eigen=0.0, layer=0):
        process_size = sum(1 for i in range(win_size) if vec_orig[i].N != 0)
        if process_size < 2:  # bavoxel.hpp:37
            return
        coe = float(sum(vec_orig[j].N for j in range(win_size)))  # bavoxel.hpp:42-44
        self.plvec_voxels.append(vec_orig)
        self.sig_vecs.append(fix)
        self.coeffs.append(coe)
        self._ctx = None

    def _context(self, n_poses):
        if n_poses != win_size:
            raise ValueError("win_size must equal len(x_stats) (benchmark_realworld.cpp:170)")
        if self._ctx is None:
            row_ptr, pose_idx, obs, fix = [0], [], [], []
            any_fix = False
            for vec, fx in zip(self.plvec_voxels, self.sig_vecs):
                for i in range(win_size):
                    if vec[i].N != 0:
                        pose_idx.append(i)
                        obs.append(vec[i].pack10())
                row_ptr.append(len(pose_idx))
                fix.append(fx.pack10() if fx is not None else [0.0] * 10)
                any_fix = any_fix or (fx is not None and fx.N != 0)
            ctx = Context(n_poses, self._device, self._precision)
            ctx.set_voxels(np.array(row_ptr), np.array(pose_idx), np.array(obs), np.array(self.coeffs),
                           np.array(fix) if any_fix else None)
            self._ctx = ctx
        return self._ctx

    def left_evaluate_acc2(self, xs, head, end):
        """-> (Hess, JacT, residual) over voxels [head, end)  (bavoxel.hpp:304-426; fix ignored as at :325)"""
        ctx = self._context(len(xs))
        return ctx.evaluate(pack_poses(xs), int(head), int(end), include_fix=False)

    def evaluate_only_residual(self, xs):
        return self._context(len(xs)).residual(pack_poses(xs))


class BALM2:
    """LM optimiser (bavoxel.hpp:984-1168)."""

    def divide_thread_left(self, x_stats, voxhess, x_ab=None):
        """-> (residual, Hess, JacT). The reference splits the voxel range over 4 std::threads and sums the
        partial results (bavoxel.hpp:1044-1056); here the split is over GPUs/CTAs inside the library."""
        H, g, r = voxhess.left_evaluate_acc2(x_stats, 0, len(voxhess.plvec_voxels))
        return r, H, g

    def only_residual(self, x_stats, voxhess, x_ab=None):
        return voxhess.evaluate_only_residual(x_stats)

    def damping_iter(self, x_stats, voxhess, verbose=True):
        """In-place LM refinement of x_stats (bavoxel.hpp:1069-1166)."""
        ctx = voxhess._context(len(x_stats))
        try:
            poses, trace, _ = ctx.damping_iter(pack_poses(x_stats), max_iter=10, u0=0.01, v0=2.0, rel_tol=1e-6,
                                               hess_includes_fix=False, gauge_mode=0, min_planes_per_pose=20,
                                               verbose=verbose)
        except L.BalmError as e:
            if e.status == L.ERR_TOO_FEW_PLANES:  # bavoxel.hpp:1079-1085: printf + exit(0)
                print("Initial error too large.")
                print("Please loose plane determination criteria for more planes.")
                print("The optimization is terminated.")
                sys.exit(0)
            raise
        unpack_poses(poses, x_stats)
        return trace
