"""Thin object wrapper over the C ABI context (include/balm_b200.h)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One GPU, one registered set of plane voxels, N poses."""

    def __init__(self, n_poses, device=0, precision=L.PREC_FP64):
        self._h = C.c_void_p()
        self.N = int(n_poses)
        self.n = 6 * self.N
        L.check(L.lib().balm_create(C.byref(self._h), self.N, int(device), int(precision)))

    def close(self):
        if self._h:
            L.lib().balm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- problem ----
    def set_voxels(self, row_ptr, pose_idx, obs10, coe, fix10=None):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        pose_idx = np.ascontiguousarray(pose_idx, dtype=np.int32)
        obs10 = np.ascontiguousarray(obs10, dtype=np.float64)
        coe = np.ascontiguousarray(coe, dtype=np.float64)
        fix10 = None if fix10 is None else np.ascontiguousarray(fix10, dtype=np.float64)
        self.M = len(row_ptr) - 1
        L.check(L.lib().balm_set_voxels(self._h, self.M, _p(row_ptr), _p(pose_idx), _p(obs10), _p(fix10), _p(coe)))

    def synth_virtual(self, n_voxels, first_voxel=0, pts_size=40, point_noise=0.01, surf_range=2.0, seed=10):
        gt = np.zeros((self.N, 12))
        init = np.zeros((self.N, 12))
        L.check(L.lib().balm_synth_virtual(self._h, int(n_voxels), int(first_voxel), int(pts_size),
                                           float(point_noise), float(surf_range), int(seed), _p(gt), _p(init)))
        self.M = int(n_voxels)
        return gt, init

    def cut_voxels(self, xyz, frame, poses12, voxel_size=2.0, layer_limit=2, min_ps=15,
                   eigen_value_array=(1.0 / 16, 1.0 / 16, 1.0 / 9)):
        """GPU association (cut_voxel + recut + tras_opt): raw body-frame points + poses -> registered plane voxels."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        frame = np.ascontiguousarray(frame, dtype=np.int32)
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        o = L.AssocOpts(float(voxel_size), int(layer_limit), int(min_ps), (C.c_double * 3)(*eigen_value_array))
        M, K = C.c_int64(), C.c_int64()
        L.check(L.lib().balm_cut_voxels(self._h, len(frame), _p(xyz), _p(frame), _p(poses12), C.byref(o), C.byref(M),
                                        C.byref(K)))
        self.M = M.value
        return M.value, K.value

    def download_voxels(self):
        K = L.lib().balm_num_obs(self._h)
        row_ptr = np.zeros(self.M + 1, dtype=np.int64)
        pose_idx = np.zeros(K, dtype=np.int32)
        obs10 = np.zeros((K, 10))
        coe = np.zeros(self.M)
        L.check(L.lib().balm_download_voxels(self._h, _p(row_ptr), _p(pose_idx), _p(obs10), _p(coe)))
        return row_ptr, pose_idx, obs10, coe

    def download_voxel_range(self, head, end):
        """CSR arrays of voxels [head, end) only (row_ptr re-based to 0)."""
        K = C.c_int64()
        L.check(L.lib().balm_download_voxel_range(self._h, int(head), int(end), None, None, None, None, C.byref(K)))
        row_ptr = np.zeros(end - head + 1, dtype=np.int64)
        pose_idx = np.zeros(K.value, dtype=np.int32)
        obs10 = np.zeros((K.value, 10))
        coe = np.zeros(end - head)
        L.check(L.lib().balm_download_voxel_range(self._h, int(head), int(end), _p(row_ptr), _p(pose_idx), _p(obs10),
                                                  _p(coe), C.byref(K)))
        return row_ptr, pose_idx, obs10, coe

    # ---- evaluation ----
    def evaluate(self, poses12, head=0, end=None, include_fix=False, want_H=True):
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        H = np.zeros((self.n, self.n), order="F") if want_H else None
        g = np.zeros(self.n)
        r = C.c_double()
        end = self.M if end is None else end
        L.check(L.lib().balm_evaluate(self._h, _p(poses12), head, end, int(include_fix), _p(H), _p(g), C.byref(r)))
        return H, g, r.value

    def residual(self, poses12):
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        r = C.c_double()
        L.check(L.lib().balm_residual(self._h, _p(poses12), C.byref(r)))
        return r.value

    def solve(self, u):
        dx = np.zeros(self.n)
        q1 = C.c_double()
        bad = C.c_int()
        L.check(L.lib().balm_solve(self._h, float(u), _p(dx), C.byref(q1), C.byref(bad)))
        return dx, q1.value, bool(bad.value)

    def damping_iter(self, poses12, max_iter=10, u0=0.01, v0=2.0, rel_tol=1e-6, hess_includes_fix=False,
                     gauge_mode=0, min_planes_per_pose=20, verbose=False, want_per_iter=False, force_hess=False):
        poses = np.array(poses12, dtype=np.float64, order="C", copy=True)
        opts = L.LmOpts(max_iter, u0, v0, rel_tol, int(hess_includes_fix), gauge_mode, min_planes_per_pose,
                        int(verbose), int(force_hess))
        trace = (L.Trace * max_iter)()
        n_it = C.c_int()
        per_iter = np.zeros((max_iter, self.N, 12)) if want_per_iter else None
        L.check(L.lib().balm_damping_iter(self._h, _p(poses), C.byref(opts), trace, C.byref(n_it), _p(per_iter)))
        tr = [dict(r1=t.r1, r2=t.r2, u=t.u, v=t.v, q=t.q, q1=t.q1, accepted=bool(t.accepted),
                   recomputed_hess=bool(t.recomputed_hess), not_pd=bool(t.not_pd)) for t in trace[:n_it.value]]
        return poses, tr, (per_iter[:n_it.value] if want_per_iter else None)

    def marginalize(self, mg_size, poses12, min_ps=15):
        """OCTO_TREE_ROOT::marginalize on the registered voxel set (bavoxel.hpp:948-963, 778-816) -> (n_voxels, n_obs)."""
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        M, K = C.c_int64(), C.c_int64()
        L.check(L.lib().balm_marginalize(self._h, int(mg_size), _p(poses12), int(min_ps), C.byref(M), C.byref(K)))
        self.M = M.value
        return M.value, K.value

    def append_scan(self, xyz, poses12, slot):
        """Associates a new scan (body-frame points) with the voxel set in HBM: observation in pose slot `slot`, every voxel
        re-judged as recut / judge_eigen do. -> (n_voxels, n_obs, n_points_matched)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        M, K, m = C.c_int64(), C.c_int64(), C.c_int64()
        L.check(L.lib().balm_append_scan(self._h, len(xyz), _p(xyz), _p(poses12), int(slot), C.byref(M), C.byref(K), C.byref(m)))
        self.M = M.value
        return M.value, K.value, m.value

    def download_keys(self, with_layers=False):
        keys = np.zeros(self.M, dtype=np.uint64)
        layers = np.zeros(self.M, dtype=np.int32)
        L.check(L.lib().balm_download_keys(self._h, _p(keys), _p(layers)))
        return (keys, layers) if with_layers else keys

    def download_fix(self):
        fix = np.zeros((self.M, 10))
        L.check(L.lib().balm_download_fix(self._h, _p(fix)))
        return fix

    def pose_covariance(self, poses12, point_noise=0.0, c_cov=None, include_fix=False, want_raw=True, want_cov=True):
        """left_jacobian_point / multi_second / H^-1 Rcov H^-T (BAs_left.hpp:342-473, 995-1023, 1089-1096).
        -> (Rcov_raw, Rcov); c_cov: K x 9 x 9 per-observation cluster covariances, or None for isotropic point noise."""
        poses12 = np.ascontiguousarray(poses12, dtype=np.float64)
        cc = None if c_cov is None else np.ascontiguousarray(c_cov, dtype=np.float64).reshape(-1, 81)
        raw = np.zeros((self.n, self.n), order="F") if want_raw else None
        cov = np.zeros((self.n, self.n), order="F") if want_cov else None
        L.check(L.lib().balm_pose_covariance(self._h, _p(poses12), _p(cc), float(point_noise), int(include_fix), _p(raw),
                                             _p(cov)))
        return raw, cov

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        L.check(L.lib().balm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        L.check(L.lib().balm_comm_init(self._h, int(rank), int(world), buf))

    # ---- instrumentation ----
    def timings(self):
        t = L.Timings()
        L.check(L.lib().balm_get_timings(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in L.Timings._fields_}

    def reset_counters(self):
        L.check(L.lib().balm_reset_counters(self._h))

    def timer_begin(self):
        L.check(L.lib().balm_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float()
        L.check(L.lib().balm_timer_end(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        L.check(L.lib().balm_sync(self._h))
