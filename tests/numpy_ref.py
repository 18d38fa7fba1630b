"""Independent float64 numpy restatement of the hot path (second oracle), written from the clean-notation
spec in SURVEY.md Appendix B rather than from the reference loop nest:

  H = sum_v coe_v * G_v diag(w_v) G_v^T + blockdiag(D_i),   G_v = [a | g_1 | g_2]

Reference: /root/reference/src/benchmark/bavoxel.hpp:304-426 (H,g,r) and :428-470 (r only).
Used only by tests to cross-check oracle/balm_oracle.c.
"""
import numpy as np


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _world_cluster(o10, p12):
    R = p12[:9].reshape(3, 3).T
    p = p12[9:12]
    P = np.array([[o10[0], o10[1], o10[2]], [o10[1], o10[3], o10[4]], [o10[2], o10[4], o10[5]]])
    v = o10[6:9]
    n = o10[9]
    Rv = R @ v
    Pw = R @ P @ R.T + np.outer(Rv, p) + np.outer(p, Rv) + n * np.outer(p, p)
    vw = Rv + n * p
    return Pw, vw, n


def residual(sc, poses12, use_fix=True):
    r = 0.0
    for a in range(len(sc["row_ptr"]) - 1):
        P = np.zeros((3, 3)); v = np.zeros(3); n = 0.0
        if use_fix and sc.get("fix10") is not None:
            f = sc["fix10"][a]
            P = np.array([[f[0], f[1], f[2]], [f[1], f[3], f[4]], [f[2], f[4], f[5]]]); v = f[6:9].copy(); n = f[9]
        for s in range(sc["row_ptr"][a], sc["row_ptr"][a + 1]):
            Pw, vw, nn = _world_cluster(sc["obs10"][s], poses12[sc["pose_idx"][s]])
            P = P + Pw; v = v + vw; n = n + nn
        vb = v / n
        lam = np.linalg.eigvalsh(P / n - np.outer(vb, vb))
        r += sc["coe"][a] * lam[0]
    return r


def evaluate(sc, poses12, include_fix=False):
    N = sc["n_poses"]; n6 = 6 * N
    H = np.zeros((n6, n6)); g = np.zeros(n6); r = 0.0
    for a in range(len(sc["row_ptr"]) - 1):
        coe = sc["coe"][a]
        sl = range(sc["row_ptr"][a], sc["row_ptr"][a + 1])
        W = [_world_cluster(sc["obs10"][s], poses12[sc["pose_idx"][s]]) for s in sl]
        P = sum(w[0] for w in W); v = sum(w[1] for w in W); NN = sum(w[2] for w in W)
        if include_fix and sc.get("fix10") is not None:
            f = sc["fix10"][a]
            P = P + np.array([[f[0], f[1], f[2]], [f[1], f[3], f[4]], [f[2], f[4], f[5]]]); v = v + f[6:9]; NN = NN + f[9]
        vb = v / NN
        lam, U = np.linalg.eigh(P / NN - np.outer(vb, vb))
        r += coe * lam[0]
        u0, u1, u2 = U[:, 0], U[:, 1], U[:, 2]
        w = np.array([-2.0 / NN ** 2, 2.0 / (lam[0] - lam[1]), 2.0 / (lam[0] - lam[2])])
        G = np.zeros((n6, 3))
        for s, (Pw, vw, nn) in zip(sl, W):
            i = sc["pose_idx"][s]
            Mt = Pw - np.outer(vw, vb)           # 3x3 top of M_i
            mb = vw - nn * vb                    # bottom row of M_i

            def gk(uk):
                top = -np.cross(uk, Mt @ u0) - np.cross(u0, Mt @ uk)
                bot = uk * (mb @ u0) + u0 * (mb @ uk)
                return np.concatenate([top, bot]) / NN
            a_i = np.concatenate([-np.cross(u0, vw), nn * u0])
            G[6 * i:6 * i + 6, 0] = a_i
            G[6 * i:6 * i + 6, 1] = gk(u1)
            G[6 * i:6 * i + 6, 2] = gk(u2)
            g[6 * i:6 * i + 6] += coe * gk(u0)
            E = hat(Mt @ u0) @ hat(u0) / NN
            V0 = np.zeros((6, 4)); V0[:3, :3] = hat(-u0); V0[3:, 3] = u0
            TCT = np.zeros((4, 4)); TCT[:3, :3] = Pw; TCT[:3, 3] = vw; TCT[3, :3] = vw; TCT[3, 3] = nn
            D = 2.0 / NN * V0 @ TCT @ V0.T
            D[:3, :3] += E + E.T
            H[6 * i:6 * i + 6, 6 * i:6 * i + 6] += coe * D
        H += coe * (G * w) @ G.T
    return H, g, r
