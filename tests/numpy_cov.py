"""numpy restatement of the pose-covariance propagation of BALM 2.0's consistency experiment.
TEST INFRASTRUCTURE (the oracle of balm_pose_covariance).

Follows /root/reference/src/simulation/BAs_left.hpp
  g1, g2                         :321-340   (d(C w)/dc for the 9 cluster parameters; [hat(w_xyz); w_3 I])
  VOX_HESS::left_jacobian_point  :342-473   (Ls = d(gradient)/d(cluster parameters of observation (a, j)), 6N x 9;
                                             Rcov += Ls c_cov_j Ls^T)
  BALM2::multi_second            :995-1023  (4 threads over voxel ranges, summed)
  BALM2::damping_iter            :1089-1096 (Rcov <- H^-1 Rcov H^-T)
and /root/reference/src/simulation/toolss.hpp:311-343 (PointCluster::push with POINT_NOISE: c_cov += Bf p_cov Bf^T).
The loop nest and the names (SpTul, T_FC, UlTC, g2_combos, G, Gkl, Lp, Ls) are the reference's. Differences, both
neutral for the sim (coeffs are 1 there, BAs_left.hpp:45): the voxel weight coe multiplies Ls (the gradient it
differentiates is sum coe * g_0), and c_cov may be given per observation or derived from (P, v, N) for isotropic point
noise, which is what push() accumulates.
"""
import numpy as np

I33 = np.eye(3)


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def g1(w):
    """BAs_left.hpp:321-331: (C w) as a linear map of the 9 parameters (P00,P01,P02,P11,P12,P22,v0,v1,v2)."""
    return np.array([[w[0], w[1], w[2], 0, 0, 0, w[3], 0, 0],
                     [0, w[0], 0, w[1], w[2], 0, 0, w[3], 0],
                     [0, 0, w[0], 0, w[1], w[2], 0, 0, w[3]],
                     [0, 0, 0, 0, 0, 0, w[0], w[1], w[2]]], dtype=np.float64)


def g2(w):
    """BAs_left.hpp:333-340."""
    out = np.zeros((6, 3))
    out[0:3] = hat(w[0:3])
    out[3:6] = w[3] * I33
    return out


# Bi = x E[0] + y E[1] + z E[2]  (toolss.hpp:317-324)
_E = np.zeros((3, 6, 3))
_E[0][0, 0] = 2; _E[0][1, 1] = 1; _E[0][2, 2] = 1
_E[1][1, 0] = 1; _E[1][3, 1] = 2; _E[1][4, 2] = 1
_E[2][2, 0] = 1; _E[2][4, 1] = 1; _E[2][5, 2] = 2


def cluster_cov_isotropic(o10, pnoise):
    """c_cov of a cluster built by push() under p_cov = pnoise^2 I (toolss.hpp:311-343), from its moments alone:
    sum Bf Bf^T = [[sum_kl P_kl E_k E_l^T, sum_k v_k E_k], [sym, N I]]."""
    P = np.array([[o10[0], o10[1], o10[2]], [o10[1], o10[3], o10[4]], [o10[2], o10[4], o10[5]]])
    v = o10[6:9]
    out = np.zeros((9, 9))
    for k in range(3):
        for l in range(3):
            out[0:6, 0:6] += P[k, l] * _E[k] @ _E[l].T
        out[0:6, 6:9] += v[k] * _E[k]
    out[6:9, 0:6] = out[0:6, 6:9].T
    out[6:9, 6:9] = o10[9] * I33
    return pnoise * pnoise * out


def _C4(o10):
    C = np.zeros((4, 4))
    C[0, 0], C[0, 1], C[0, 2], C[1, 1], C[1, 2], C[2, 2] = o10[0:6]
    C[1, 0], C[2, 0], C[2, 1] = C[0, 1], C[0, 2], C[1, 2]
    C[0:3, 3] = o10[6:9]
    C[3, 0:3] = o10[6:9]
    C[3, 3] = o10[9]
    return C


def left_jacobian_point(n_poses, row_ptr, pose_idx, obs10, coe, poses12, fix10=None, c_cov=None, pnoise=None,
                        beg=0, end=None, return_Ls=False):
    """BAs_left.hpp:342-473 over voxels [beg, end). -> Rcov (6N x 6N); with return_Ls also {(a, s): Ls}."""
    end = len(row_ptr) - 1 if end is None else end
    n = 6 * n_poses
    Rcov = np.zeros((n, n))
    l = 0
    T = []
    for i in range(n_poses):
        Ti = np.eye(4)
        Ti[0:3, 0:3] = poses12[i][:9].reshape(3, 3).T
        Ti[0:3, 3] = poses12[i][9:12]
        T.append(Ti)
    Sp = np.zeros((3, 4)); Sp[0:3, 0:3] = I33
    F = np.zeros((4, 4)); F[3, 3] = 1
    all_Ls = {}
    for a in range(beg, end):
        C = np.zeros((4, 4)) if fix10 is None else _C4(fix10[a])
        slots = list(range(row_ptr[a], row_ptr[a + 1]))
        seen = [pose_idx[s] for s in slots]
        TC, TCT = {}, {}
        for s, j in zip(slots, seen):
            Co = _C4(obs10[s])
            TC[j] = T[j] @ Co
            TCT[j] = TC[j] @ T[j].T
            C = C + TCT[j]
        NN = C[3, 3]
        C = C / NN
        v_bar = C[0:3, 3]
        lmbd, Uev = np.linalg.eigh(C[0:3, 0:3] - np.outer(v_bar, v_bar))
        u = [Uev[:, 0], Uev[:, 1], Uev[:, 2]]
        U = []
        for k in range(3):
            Uk = np.zeros((6, 4))
            Uk[0:3, 0:3] = hat(-u[k])
            Uk[3:6, 3] = u[k]
            U.append(Uk)
        SpTul = Sp.T @ u[l]
        T_FC, UlTC, g2_combos = {}, {}, {}
        for p in seen:
            T_FC[p] = T[p].T - F @ C
            UlTC[p] = U[l] @ TC[p]
            w2 = TC[p] @ T_FC[p] @ SpTul
            g2_combos[p] = g2(w2) + UlTC[p] @ T_FC[p] @ Sp.T
        for s, j in zip(slots, seen):
            g1_TSu = g1(T[j].T @ SpTul)
            G = np.zeros((3, 9))
            for k in range(3):
                if k != l:
                    Gkl = T_FC[j].T @ g1_TSu - T[j] @ g1(F @ C @ Sp.T @ u[l])
                    G += 1.0 / (lmbd[l] - lmbd[k]) / NN * np.outer(u[k], u[k]) @ Sp @ Gkl
            Ls = np.zeros((n, 9))
            for p in seen:
                Lp = g2_combos[p] @ G
                Lp += -1.0 / NN * UlTC[p] @ F @ T[j] @ g1_TSu
                if j == p:
                    Lp += U[l] @ T[p] @ g1(T_FC[p] @ SpTul)
                Ls[6 * p:6 * p + 6] = 2.0 / NN * Lp
            Ls *= coe[a]
            cc = c_cov[s] if c_cov is not None else cluster_cov_isotropic(obs10[s], pnoise)
            Rcov += Ls @ cc @ Ls.T
            if return_Ls:
                all_Ls[(a, s)] = Ls
    return (Rcov, all_Ls) if return_Ls else Rcov


def pose_covariance(H, Rcov):
    """BAs_left.hpp:1093-1094: hess_inv * Rcov * hess_inv^T."""
    Hi = np.linalg.inv(H)
    return Hi @ Rcov @ Hi.T
