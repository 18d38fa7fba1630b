"""numpy restatement of the reference's RIGHT-update evaluator VOX_HESS::acc_evaluate2
(/root/reference/src/benchmark/bavoxel.hpp:53-158), residual and gradient only.  TEST INFRASTRUCTURE.

acc_evaluate2 is dead code in the reference (its only call site is commented out, bavoxel.hpp:1108), but it is the
one remaining INDEPENDENT reference-side statement of the cost and its first derivative:
it perturbs the rotation on the right and the translation in the world frame,  R <- R Exp(phi),  p <- p + dt
(what its Jacobian differentiates -- established by finite differences in tests/test_oracle.py), and builds that
Jacobian from body-frame quantities (P_i, v_i, R_i^T u_k) instead of the world-frame 4x4 products of
left_evaluate_acc2.  To first order the left update  R <- Exp(phi_l) R, p <- Exp(phi_l) p + dt_l  is the same motion when

    dx_left = LL dx_right,   LL_i = [[R_i, 0], [hat(p_i) R_i, I]]     =>     g_right = LL^T g_left

(the commented block bavoxel.hpp:279-300 sketches this map with R_i in the lower-right block, i.e. for a body-frame
translation increment; acc_evaluate2 itself uses the world-frame one).

The Hessians agree only up to parametrisation-dependent second-order terms, so only r and g are compared.
Loop structure follows the reference line by line (variable names kept: vBar, uk, RiTuk, combo1, combo2, Auk, jjt).
"""
import numpy as np


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _unpack(o10):
    P = np.array([[o10[0], o10[1], o10[2]], [o10[1], o10[3], o10[4]], [o10[2], o10[4], o10[5]]])
    return P, o10[6:9].copy(), float(o10[9])


def acc_evaluate2(n_poses, row_ptr, pose_idx, obs10, coe, poses12, fix10=None, head=0, end=None):
    """-> (JacT (6N), residual) of bavoxel.hpp:53-158 over voxels [head, end) (the fix cluster is INCLUDED, :68)."""
    end = len(row_ptr) - 1 if end is None else end
    kk = 0
    JacT = np.zeros(6 * n_poses)
    residual = 0.0
    Rs = [poses12[i][:9].reshape(3, 3).T for i in range(n_poses)]
    ps = [poses12[i][9:12] for i in range(n_poses)]
    for a in range(head, end):
        c = coe[a]
        sigP, sigv, sigN = (np.zeros((3, 3)), np.zeros(3), 0.0) if fix10 is None else _unpack(fix10[a])
        slots = range(row_ptr[a], row_ptr[a + 1])
        for s in slots:                                        # :69-74  sig += transform(sig_orig[i], xs[i])
            P, v, n = _unpack(obs10[s])
            R, p = Rs[pose_idx[s]], ps[pose_idx[s]]
            Rv = R @ v
            sigP = sigP + R @ P @ R.T + np.outer(Rv, p) + np.outer(p, Rv) + n * np.outer(p, p)   # tools.hpp:333-339
            sigv = sigv + Rv + n * p
            sigN += n
        vBar = sigv / sigN                                     # :76
        lmbd, U = np.linalg.eigh(sigP / sigN - np.outer(vBar, vBar))   # :77-79
        NN = int(sigN)                                         # :80
        uk = U[:, kk]
        for s in slots:                                        # :92-118
            Pi, vi, ni = _unpack(obs10[s])
            i = pose_idx[s]
            Ri = Rs[i]
            vihat = hat(vi)
            RiTuk = Ri.T @ uk
            RiTukhat = hat(RiTuk)
            PiRiTuk = Pi @ RiTuk
            ti_v = ps[i] - vBar
            ukTti_v = uk @ ti_v
            combo1 = hat(PiRiTuk) + vihat * ukTti_v
            combo2 = Ri @ vi + ni * ti_v
            Auk = np.zeros((3, 6))
            Auk[:, 0:3] = (Ri @ Pi + np.outer(ti_v, vi)) @ RiTukhat - Ri @ combo1
            Auk[:, 3:6] = np.outer(combo2, uk) + (combo2 @ uk) * np.eye(3)
            Auk /= NN
            jjt = Auk.T @ uk                                   # :117
            JacT[6 * i:6 * i + 6] += c * jjt                   # :118
        residual += c * lmbd[kk]                               # :150
    return JacT, residual


def left_to_right_gradient(g_left, poses12):
    """g_right = LL^T g_left with LL_i = [[R_i, 0], [hat(p_i) R_i, I]] (see the module docstring)."""
    out = np.zeros_like(g_left)
    for i in range(len(poses12)):
        R = poses12[i][:9].reshape(3, 3).T
        p = poses12[i][9:12]
        LL = np.block([[R, np.zeros((3, 3))], [hat(p) @ R, np.eye(3)]])
        out[6 * i:6 * i + 6] = LL.T @ g_left[6 * i:6 * i + 6]
    return out


def right_update(poses12, dx, exp_so3):
    """R <- R Exp(phi), p <- p + dt  (the update acc_evaluate2 differentiates; only used for finite differences)."""
    out = np.array(poses12, dtype=np.float64, copy=True)
    for i in range(len(out)):
        R = out[i][:9].reshape(3, 3).T
        p = out[i][9:12]
        Rn = R @ exp_so3(dx[6 * i:6 * i + 3])
        pn = p + dx[6 * i + 3:6 * i + 6]
        out[i][:9] = Rn.T.reshape(9)
        out[i][9:12] = pn
    return out
