"""ctypes binding of oracle/_ref/libbalm_ref.so: the REFERENCE'S OWN code (bavoxel.hpp + tools.hpp, compiled where they lie
under /root/reference against the Eigen/PCL/ROS stand-ins of oracle/ref_stubs).  TEST INFRASTRUCTURE ONLY.

Used by tests/test_reference_pin.py to pin the restated oracles against what the reference's code computes. The library is
built by `make -C oracle` when /root/reference is present (this container); on a box without the reference the prebuilt
.so travels with the repo snapshot, and `available()` says whether it can be loaded."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libbalm_ref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_problem_create.restype = C.c_void_p
        L.ref_problem_create.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 4
        L.ref_problem_destroy.argtypes = [C.c_void_p]
        L.ref_problem_pushed.restype = C.c_int64
        L.ref_problem_pushed.argtypes = [C.c_void_p]
        L.ref_problem_coeffs.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_left_evaluate_acc2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_evaluate_only_residual.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_acc_evaluate2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_divide_thread_left.restype = C.c_double
        L.ref_divide_thread_left.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_damping_iter.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_lm_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.ref_exp.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_log.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_session_create.restype = C.c_void_p
        L.ref_session_create.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.ref_session_destroy.argtypes = [C.c_void_p]
        L.ref_session_cut_voxel.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_session_recut.argtypes = [C.c_void_p, C.c_int]
        L.ref_session_marginalize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_session_export.restype = C.c_int64
        L.ref_session_export.argtypes = [C.c_void_p, C.c_int]
        L.ref_session_num_obs.restype = C.c_int64
        L.ref_session_num_obs.argtypes = [C.c_void_p]
        L.ref_session_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ref_session_fetch_layers.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Problem:
    """VOX_HESS filled through the reference's push_voxel from CSR arrays (every voxel a vector<PointCluster> of win_size slots)."""

    def __init__(self, n_poses, row_ptr, pose_idx, obs10, fix10=None):
        self.N = int(n_poses)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.pose_idx = np.ascontiguousarray(pose_idx, dtype=np.int32)
        self.obs10 = np.ascontiguousarray(obs10, dtype=np.float64)
        self.fix10 = None if fix10 is None else np.ascontiguousarray(fix10, dtype=np.float64)
        self.M = len(self.row_ptr) - 1
        self.h = C.c_void_p(lib().ref_problem_create(self.N, self.M, _p(self.row_ptr), _p(self.pose_idx), _p(self.obs10), _p(self.fix10)))

    def __del__(self):
        try:
            lib().ref_problem_destroy(self.h)
        except Exception:
            pass

    def pushed(self):
        return lib().ref_problem_pushed(self.h)

    def coeffs(self):
        c = np.zeros(self.pushed())
        lib().ref_problem_coeffs(self.h, _p(c))
        return c

    def left_evaluate_acc2(self, poses12, head=0, end=None):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), order="F")
        g = np.zeros(n)
        r = C.c_double()
        lib().ref_left_evaluate_acc2(self.h, _p(poses12), head, self.pushed() if end is None else end, _p(H), _p(g), C.byref(r))
        return H, g, r.value

    def acc_evaluate2(self, poses12, head=0, end=None):
        """The reference's right-update evaluator (bavoxel.hpp:53-158) -> (Hess, JacT, residual)."""
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), order="F")
        g = np.zeros(n)
        r = C.c_double()
        lib().ref_acc_evaluate2(self.h, _p(poses12), head, self.pushed() if end is None else end, _p(H), _p(g), C.byref(r))
        return H, g, r.value

    def evaluate_only_residual(self, poses12):
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        r = C.c_double()
        lib().ref_evaluate_only_residual(self.h, _p(poses12), C.byref(r))
        return r.value

    def divide_thread_left(self, poses12):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), order="F")
        g = np.zeros(n)
        r = lib().ref_divide_thread_left(self.h, _p(poses12), _p(H), _p(g))
        return H, g, r

    def damping_iter(self, poses12):
        p = np.array(poses12, dtype=np.float64, order="C", copy=True)
        lib().ref_damping_iter(self.h, _p(p))
        return p


def lm_solve(H, g, u):
    n = len(g)
    H = np.asfortranarray(H, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    dx = np.zeros(n)
    lib().ref_lm_solve(n, _p(H), _p(g), float(u), _p(dx))
    return dx


def exp_so3(phi):
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    R = np.zeros(9)
    lib().ref_exp(_p(phi), _p(R))
    return R.reshape(3, 3).T


def log_so3(R):
    Rcm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T).reshape(9)
    phi = np.zeros(3)
    lib().ref_log(_p(Rcm), _p(phi))
    return phi


class Session:
    """The reference's octree (unordered_map<VOXEL_LOC, OCTO_TREE_ROOT*>) kept alive between calls."""

    def __init__(self, n_poses, voxel_size=2.0, layer_limit=2, min_ps=15, eigen_value_array=(1 / 16, 1 / 16, 1 / 16)):
        self.N = int(n_poses)
        e = np.array(list(eigen_value_array)[:3], dtype=np.float64)
        self.h = C.c_void_p(lib().ref_session_create(self.N, float(voxel_size), int(layer_limit), int(min_ps), _p(e)))

    def __del__(self):
        try:
            lib().ref_session_destroy(self.h)
        except Exception:
            pass

    def cut_voxel(self, xyz, pose12, fnum):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        pose12 = np.ascontiguousarray(pose12, dtype=np.float64)
        lib().ref_session_cut_voxel(self.h, len(xyz), _p(xyz), _p(pose12), int(fnum))

    def recut(self, win_count):
        lib().ref_session_recut(self.h, int(win_count))

    def marginalize(self, mg_size, poses12, win_count):
        p = None if poses12 is None else np.ascontiguousarray(poses12, dtype=np.float64)
        lib().ref_session_marginalize(self.h, int(mg_size), _p(p), int(win_count))

    def export(self, win_count, with_layers=False):
        """tras_opt over all roots -> (keys, row_ptr, pose_idx, obs10, fix10, coe[, layers]), sorted by node key."""
        M = lib().ref_session_export(self.h, int(win_count))
        K = lib().ref_session_num_obs(self.h)
        keys = np.zeros(M, dtype=np.uint64)
        rp = np.zeros(M + 1, dtype=np.int64)
        pi = np.zeros(K, dtype=np.int32)
        ob = np.zeros((K, 10))
        fx = np.zeros((M, 10))
        co = np.zeros(M)
        lib().ref_session_fetch(self.h, _p(keys), _p(rp), _p(pi), _p(ob), _p(fx), _p(co))
        out = (keys.astype(np.int64), rp, pi, ob, fx, co)
        if with_layers:
            lay = np.zeros(M, dtype=np.int32)
            lib().ref_session_fetch_layers(self.h, _p(lay))
            out = out + (lay,)
        return out


# ---------------- the consistency experiment's code (src/simulation/BAs_left.hpp + toolss.hpp) ----------------
LIB_SIM = os.path.join(_HERE, "_ref", "libbalm_ref_sim.so")
_lib_sim = None


def sim_available():
    return os.path.exists(LIB_SIM)


def lib_sim():
    global _lib_sim
    if _lib_sim is None:
        L = C.CDLL(LIB_SIM)
        L.ref_sim_problem_create.restype = C.c_void_p
        L.ref_sim_problem_create.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 5
        L.ref_sim_problem_destroy.argtypes = [C.c_void_p]
        L.ref_sim_left_jacobian_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_sim_left_evaluate_acc2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_sim_push_points.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _lib_sim = L
    return _lib_sim


class SimProblem:
    """The sim's VOX_HESS (coe = 1, BAs_left.hpp:45) with per-cluster c_cov (K x 9 x 9)."""

    def __init__(self, n_poses, row_ptr, pose_idx, obs10, fix10=None, c_cov=None):
        self.N = int(n_poses)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.pose_idx = np.ascontiguousarray(pose_idx, dtype=np.int32)
        self.obs10 = np.ascontiguousarray(obs10, dtype=np.float64)
        self.fix10 = None if fix10 is None else np.ascontiguousarray(fix10, dtype=np.float64)
        self.cc = None if c_cov is None else np.ascontiguousarray(c_cov, dtype=np.float64).reshape(-1, 81)
        self.M = len(self.row_ptr) - 1
        self.h = C.c_void_p(lib_sim().ref_sim_problem_create(self.N, self.M, _p(self.row_ptr), _p(self.pose_idx), _p(self.obs10),
                                                             _p(self.fix10), _p(self.cc)))

    def __del__(self):
        try:
            lib_sim().ref_sim_problem_destroy(self.h)
        except Exception:
            pass

    def left_jacobian_point(self, poses12, beg=0, end=None):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        R = np.zeros((n, n), order="F")
        lib_sim().ref_sim_left_jacobian_point(self.h, _p(poses12), beg, self.M if end is None else end, _p(R))
        return R

    def left_evaluate_acc2(self, poses12, head=0, end=None):
        n = 6 * self.N
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((n, n), order="F")
        g = np.zeros(n)
        r = C.c_double()
        lib_sim().ref_sim_left_evaluate_acc2(self.h, _p(poses12), head, self.M if end is None else end, _p(H), _p(g), C.byref(r))
        return H, g, r.value


def sim_push_points(xyz, point_noise):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    o = np.zeros(10)
    cc = np.zeros(81)
    lib_sim().ref_sim_push_points(len(xyz), _p(xyz), float(point_noise), _p(o), _p(cc))
    return o, cc.reshape(9, 9)
