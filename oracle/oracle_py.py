"""ctypes binding of the CPU oracle (oracle/balm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (balm_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile liboracle.so / liboracle_native.so with gcc (no-op if up to date)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "balm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return so


class _Problem(C.Structure):
    _fields_ = [("n_poses", C.c_int), ("n_voxels", C.c_int64), ("row_ptr", C.c_void_p),
                ("pose_idx", C.c_void_p), ("obs10", C.c_void_p), ("fix10", C.c_void_p), ("coe", C.c_void_p)]


class LmOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("u0", C.c_double), ("v0", C.c_double), ("rel_tol", C.c_double),
                ("hess_includes_fix", C.c_int), ("threads", C.c_int), ("gauge_mode", C.c_int),
                ("min_planes_per_pose", C.c_int)]


class Trace(C.Structure):
    _fields_ = [("r1", C.c_double), ("r2", C.c_double), ("u", C.c_double), ("v", C.c_double),
                ("q", C.c_double), ("q1", C.c_double), ("accepted", C.c_int), ("recomputed_hess", C.c_int)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Holds one problem (CSR voxel->observation layout, see balm_oracle.h)."""

    def __init__(self, n_poses, row_ptr, pose_idx, obs10, coe, fix10=None, native=False):
        build()
        name = "liboracle_native.so" if native else "liboracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        self.N = int(n_poses)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.pose_idx = np.ascontiguousarray(pose_idx, dtype=np.int32)
        self.obs10 = np.ascontiguousarray(obs10, dtype=np.float64).reshape(-1, 10)
        self.coe = np.ascontiguousarray(coe, dtype=np.float64)
        self.fix10 = None if fix10 is None else np.ascontiguousarray(fix10, dtype=np.float64).reshape(-1, 10)
        self.M = len(self.row_ptr) - 1
        assert self.row_ptr[-1] == len(self.pose_idx) == len(self.obs10)
        self.pb = _Problem(self.N, self.M, _p(self.row_ptr), _p(self.pose_idx), _p(self.obs10),
                           _p(self.fix10), _p(self.coe))
        self.lib.orc_divide_thread_left.restype = C.c_double

    def evaluate(self, poses12, head=0, end=None, include_fix=False):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), dtype=np.float64, order="F")
        g = np.zeros(n)
        r = C.c_double()
        end = self.M if end is None else end
        self.lib.orc_left_evaluate_acc2(C.byref(self.pb), _p(poses12), C.c_int64(head), C.c_int64(end),
                                        C.c_int(int(include_fix)), _p(H), _p(g), C.byref(r))
        return H, g, r.value

    def evaluate_threads(self, poses12, threads=4, include_fix=False):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), dtype=np.float64, order="F")
        g = np.zeros(n)
        r = self.lib.orc_divide_thread_left(C.byref(self.pb), _p(poses12), C.c_int(threads),
                                            C.c_int(int(include_fix)), _p(H), _p(g))
        return H, g, r

    def residual(self, poses12):
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        r = C.c_double()
        self.lib.orc_evaluate_only_residual(C.byref(self.pb), _p(poses12), C.byref(r))
        return r.value

    def lm_step(self, H, g, u, poses12):
        n = 6 * self.N
        H = np.asfortranarray(H, dtype=np.float64)
        g = np.ascontiguousarray(g, dtype=np.float64)
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        dx = np.zeros(n)
        trial = np.zeros_like(poses12)
        q1 = C.c_double()
        self.lib.orc_lm_step(C.c_int(self.N), _p(H), _p(g), C.c_double(u), _p(poses12), _p(dx), _p(trial),
                             C.byref(q1))
        return dx, trial, q1.value

    def damping_iter(self, poses12, max_iter=10, u0=0.01, v0=2.0, rel_tol=1e-6, hess_includes_fix=False,
                     threads=4, gauge_mode=0, min_planes_per_pose=20):
        poses = np.array(poses12, dtype=np.float64, order="C", copy=True)
        opts = LmOpts(max_iter, u0, v0, rel_tol, int(hess_includes_fix), threads, gauge_mode,
                      min_planes_per_pose)
        trace = (Trace * max_iter)()
        n_it = C.c_int()
        per_iter = np.zeros((max_iter, self.N, 12))
        st = self.lib.orc_damping_iter(C.byref(self.pb), _p(poses), C.byref(opts), trace, C.byref(n_it),
                                       _p(per_iter))
        tr = [dict(r1=t.r1, r2=t.r2, u=t.u, v=t.v, q=t.q, q1=t.q1, accepted=bool(t.accepted),
                   recomputed_hess=bool(t.recomputed_hess)) for t in trace[:n_it.value]]
        return st, poses, tr, per_iter[:n_it.value]


def _lib():
    build()
    return C.CDLL(os.path.join(_HERE, "liboracle.so"))


def exp_so3(phi):
    R = np.zeros(9)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    _lib().orc_exp_so3(_p(phi), _p(R))
    return R.reshape(3, 3).T  # col-major -> numpy matrix


def log_so3(R):
    Rcm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T).reshape(9)
    phi = np.zeros(3)
    _lib().orc_log_so3(_p(Rcm), _p(phi))
    return phi


def eig3(A):
    Acm = np.ascontiguousarray(np.asarray(A, dtype=np.float64).T).reshape(9)
    lam = np.zeros(3)
    U = np.zeros(9)
    _lib().orc_eig3(_p(Acm), _p(lam), _p(U))
    return lam, U.reshape(3, 3).T


def ldlt_solve(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    zp = _lib().orc_ldlt_solve(C.c_int(len(b)), _p(A), _p(b), _p(x))
    return x, zp


def rmse(est, gt):
    est = np.ascontiguousarray(est, dtype=np.float64)
    gt = np.ascontiguousarray(gt, dtype=np.float64)
    rot, tran = C.c_double(), C.c_double()
    _lib().orc_rmse(C.c_int(len(est)), _p(est), _p(gt), C.byref(rot), C.byref(tran))
    return rot.value, tran.value
