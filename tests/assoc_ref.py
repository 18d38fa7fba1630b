"""numpy restatement of the reference's association step -- the oracle of the GPU association (balm_cut_voxels).

  cut_voxel                                     src/benchmark/bavoxel.hpp:1170-1223
      world point = R p + t; float32 voxel coordinate loc = world / voxel_size, "-1 for negatives", truncation to
      int64 (:1178-1184); root centre (0.5 + idx) * voxel_size stored as float (:1213-1215), quater_length = size/4
  OCTO_TREE_NODE::recut / judge_eigen / cut_func bavoxel.hpp:654-776
      a node with <= min_ps points is dropped; planar (lambda0/lambda1 < eigen_value_array[layer], covariance of ALL the
      node's world points) -> plane leaf; otherwise split into octants around the float centre (child centre =
      centre +- quater, quater/2) until layer_limit
  tras_opt -> VOX_HESS::push_voxel              bavoxel.hpp:908-929, 30-51
      leaves with >= min_ps points and >= 2 observing frames become plane voxels; per-frame clusters are the moments
      of the BODY-frame points (sig_orig), coe = total point count
Returns CSR arrays ordered by the 63-bit node key (root x,y,z biased by 2^18 in 19 bits each, then the two octant
digits, 7 = "not split"), which is also the order the GPU implementation emits.
"""
import numpy as np

DEFAULTS = dict(voxel_size=2.0, layer_limit=2, min_ps=15, layer_size=(30, 30, 30, 30),
                eigen_value_array=(1.0 / 16, 1.0 / 16, 1.0 / 9, 1.0 / 16))


def root_index(world, voxel_size):
    loc = (world / voxel_size).astype(np.float32)
    loc = np.where(loc < 0, loc - np.float32(1.0), loc)
    return np.trunc(loc).astype(np.int64)


def node_key(root, o1, o2):
    b = 1 << 18
    return (((int(root[0]) + b) << 44) | ((int(root[1]) + b) << 25) | ((int(root[2]) + b) << 6) | (o1 << 3) | o2)


def cut_voxels(points_body, frames, poses, with_layers=False, **kw):
    """points_body: n x 3 float64 (already rounded to the float32 the PCD stores), frames: n int, poses: list of (R, p).
    with_layers: also return every leaf's octree layer -- a key digit 7 is both "not split" and octant 7, so the layer
    cannot be read back from the key."""
    o = dict(DEFAULTS)
    o.update(kw)
    vs, lim, min_ps = o["voxel_size"], o["layer_limit"], o["min_ps"]
    R = np.stack([r for r, _ in poses])
    t = np.stack([p for _, p in poses])
    world = np.einsum("nij,nj->ni", R[frames], points_body) + t[frames]
    key = root_index(world, vs)
    order = np.lexsort((frames, key[:, 2], key[:, 1], key[:, 0]))
    key, pb, pw, fr = key[order], points_body[order], world[order], frames[order]
    change = np.any(key[1:] != key[:-1], axis=1)
    starts = np.concatenate([[0], np.nonzero(change)[0] + 1, [len(key)]])
    leaves = []

    def recut(pb_, pw_, fr_, center, quater, layer, root, path):
        n = len(pb_)
        if n <= min_ps:
            return
        c = pw_.sum(0) / n
        cov = pw_.T @ pw_ / n - np.outer(c, c)
        lam = np.linalg.eigvalsh(cov)
        if lam[0] / lam[1] < float(np.float32(o["eigen_value_array"][layer])):  # `float eigen_value_array[]`, bavoxel.hpp:11
            leaves.append((node_key(root, path[0], path[1]), pb_, fr_, layer))
            return
        if layer == lim:
            return
        octant = pw_ > center[None, :]
        leaf = 4 * octant[:, 0] + 2 * octant[:, 1] + octant[:, 2]
        for lf in range(8):
            m = leaf == lf
            if not m.any():
                continue
            xyz = np.array([(lf >> 2) & 1, (lf >> 1) & 1, lf & 1])
            cc = (center.astype(np.float32) + (2 * xyz - 1).astype(np.float32) * np.float32(quater)).astype(np.float64)
            p2 = (lf, 7) if layer == 0 else (path[0], lf)
            recut(pb_[m], pw_[m], fr_[m], cc, quater / 2, layer + 1, root, p2)

    for a, b in zip(starts[:-1], starts[1:]):
        center = ((0.5 + key[a]) * vs).astype(np.float32).astype(np.float64)
        recut(pb[a:b], pw[a:b], fr[a:b], center, vs / 4.0, 0, key[a], (7, 7))
    leaves.sort(key=lambda x: x[0])
    row_ptr, pose_idx, obs, coe, keys, layers = [0], [], [], [], [], []
    for k, q, f, lay in leaves:
        fs = np.unique(f)
        if len(q) < min_ps or len(fs) < 2:
            continue
        for ff in fs:
            x = q[f == ff]
            P = x.T @ x
            v = x.sum(0)
            obs.append([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], float(len(x))])
            pose_idx.append(int(ff))
        row_ptr.append(len(pose_idx))
        coe.append(float(len(q)))
        keys.append(k)
        layers.append(lay)
    out = (np.array(row_ptr, dtype=np.int64), np.array(pose_idx, dtype=np.int32), np.array(obs, dtype=np.float64),
           np.array(coe, dtype=np.float64), np.array(keys, dtype=np.int64))
    return out + (np.array(layers, dtype=np.int32),) if with_layers else out


def synthetic_scans(n_poses=12, pts_per_scan=6000, seed=5, room=6.0):
    """Lidar-like scans of a box room with a few interior planes and clutter, seen from a short trajectory with noisy
    initial poses -- exercises planar roots, split roots (wall/floor corners) and non-planar clutter."""
    import scenes
    rng = np.random.default_rng(seed)
    Rs = [scenes.exp_so3(np.array([0.02 * i, -0.015 * i, 0.05 * i])) for i in range(n_poses)]
    ps = [np.array([0.15 * i, 0.1 * np.sin(0.5 * i), 0.02 * i]) for i in range(n_poses)]
    pts, frs = [], []
    for i in range(n_poses):
        n = pts_per_scan
        face = rng.integers(0, 8, n)
        u, v = rng.uniform(-room, room, n), rng.uniform(-room, room, n)
        w = np.zeros((n, 3))
        for fidx in range(6):  # the six room faces
            m = face == fidx
            ax, sgn = fidx // 2, (1 if fidx % 2 else -1)
            q = np.zeros((m.sum(), 3))
            q[:, ax] = sgn * room
            q[:, (ax + 1) % 3] = u[m]
            q[:, (ax + 2) % 3] = v[m]
            w[m] = q
        m = face == 6  # a slanted interior plane
        w[m] = np.stack([u[m] * 0.5, v[m] * 0.5, 0.3 * u[m] + 0.2 * v[m] - 1.0], axis=1)
        m = face == 7  # clutter (non-planar blobs)
        w[m] = rng.normal(0, 0.6, (m.sum(), 3)) + rng.choice([-3.0, 0.0, 3.0], (m.sum(), 3))
        w += rng.normal(0, 0.01, w.shape)
        body = (w - ps[i]) @ Rs[i]  # R^T (x - p)
        pts.append(body.astype(np.float32).astype(np.float64))
        frs.append(np.full(n, i, dtype=np.int32))
    noisy = [(Rs[i] @ scenes.exp_so3(rng.normal(0, 0.003, 3)), ps[i] + rng.normal(0, 0.01, 3)) for i in range(n_poses)]
    noisy[0] = (Rs[0], ps[0])
    return np.concatenate(pts), np.concatenate(frs), noisy


def marginalize_ref(n_poses, row_ptr, pose_idx, obs10, fix10, poses12, mg_size, min_ps=15):
    """numpy restatement of the sliding-window step on a flattened set of plane leaves:
      OCTO_TREE_NODE::to_margi   bavoxel.hpp:778-816  (sig_tran = transform(sig_orig, x_poses); fix_point += the oldest
                                                       mg_size clusters if fix_point.N < 50 and push_state == 1; shift)
      OCTO_TREE_NODE::tras_opt   bavoxel.hpp:908-929  (fewer than min_ps points left -> not pushed)
      VOX_HESS::push_voxel       bavoxel.hpp:30-51    (fewer than 2 observing scans -> not pushed; coe = points left)
    -> row_ptr, pose_idx, obs10, fix10, coe of the voxels that are pushed again, in the original order."""
    rp, pi, ob, fx, co = [0], [], [], [], []
    for a in range(len(row_ptr) - 1):
        fix = np.zeros(10) if fix10 is None else np.array(fix10[a], dtype=np.float64)
        absorb = int(fix[9]) < 50
        keep = []
        for s in range(row_ptr[a], row_ptr[a + 1]):
            i = pose_idx[s]
            if i < mg_size:
                if absorb:
                    o = obs10[s]
                    R = poses12[i][:9].reshape(3, 3).T
                    p = poses12[i][9:12]
                    P = np.array([[o[0], o[1], o[2]], [o[1], o[3], o[4]], [o[2], o[4], o[5]]])
                    v, n = o[6:9], o[9]
                    Rv = R @ v                                                   # PointCluster::transform, tools.hpp:333-339
                    Pw = R @ P @ R.T + np.outer(Rv, p) + np.outer(p, Rv) + n * np.outer(p, p)
                    fix += np.array([Pw[0, 0], Pw[0, 1], Pw[0, 2], Pw[1, 1], Pw[1, 2], Pw[2, 2], *(Rv + n * p), n])
            else:
                keep.append(s)
        pts = sum(obs10[s][9] for s in keep)
        if int(pts) < min_ps or len(keep) < 2:
            continue
        for s in keep:
            ob.append(obs10[s])
            pi.append(pose_idx[s] - mg_size)
        rp.append(len(pi))
        fx.append(fix)
        co.append(float(pts))
    return (np.array(rp, dtype=np.int64), np.array(pi, dtype=np.int32), np.array(ob).reshape(-1, 10),
            np.array(fx).reshape(-1, 10), np.array(co))


def point_keys(world, voxel_size):
    """Full 63-bit keys of world points: root index (cut_voxel) + the octants cut_func chooses at layers 1 and 2 with the
    reference's float32 centres (bavoxel.hpp:1213-1216, 700-735). -> (root n x 3, oct1 n, oct2 n)."""
    root = root_index(world, voxel_size)
    c0 = ((0.5 + root) * voxel_size).astype(np.float32)
    quater = np.float32(voxel_size / 4.0)
    b1 = (world > c0.astype(np.float64)).astype(np.int64)
    c1 = c0 + (2 * b1 - 1).astype(np.float32) * quater
    b2 = (world > c1.astype(np.float64)).astype(np.int64)
    return root, 4 * b1[:, 0] + 2 * b1[:, 1] + b1[:, 2], 4 * b2[:, 0] + 2 * b2[:, 1] + b2[:, 2]


def append_scan_ref(keys, layers, row_ptr, pose_idx, obs10, fix10, coe, points_body, poses12, slot, **kw):
    """numpy restatement of balm_append_scan: the new scan's points find their plane leaf (cut_voxel + cut_func), their
    body-frame moments become the leaf's observation in pose slot `slot`, and every leaf is re-judged as recut does
    (bavoxel.hpp:737-776 -> judge_eigen :654-699): fix_point + all window clusters transformed by the current poses must be
    planar (eigen ratio below the layer's threshold), hold more than min_ps window points and be seen by >= 2 scans.
    A point descends root -> octant -> octant and stops at the first plane leaf on its path (cut_func :700-735): a leaf
    only matches at its own layer (`layers`; the key digit 7 alone is ambiguous with octant 7).
    -> keys, layers, row_ptr, pose_idx, obs10, fix10, coe of the leaves that are pushed again, and the number of matched
    points."""
    o = dict(DEFAULTS)
    o.update(kw)
    vs, min_ps = o["voxel_size"], o["min_ps"]
    keymap = {int(k): a for a, k in enumerate(keys)}
    R = poses12[slot][:9].reshape(3, 3).T
    t = poses12[slot][9:12]
    world = points_body @ R.T + t
    root, o1, o2 = point_keys(world, vs)
    new_pts = {}
    matched = 0
    for i in range(len(points_body)):
        for lay, cand in enumerate((node_key(root[i], 7, 7), node_key(root[i], int(o1[i]), 7),
                                    node_key(root[i], int(o1[i]), int(o2[i])))):
            a = keymap.get(cand)
            if a is not None and int(layers[a]) == lay:
                new_pts.setdefault(a, []).append(points_body[i])
                matched += 1
                break
    out_k, out_l, rp, pi, ob, fx, co = [], [], [0], [], [], [], []
    for a, k in enumerate(keys):
        rows = [(int(pose_idx[s]), np.array(obs10[s], dtype=np.float64)) for s in range(row_ptr[a], row_ptr[a + 1])]
        if a in new_pts:
            x = np.array(new_pts[a])
            P = x.T @ x
            v = x.sum(0)
            assert all(p < slot for p, _ in rows)
            rows.append((slot, np.array([P[0, 0], P[0, 1], P[0, 2], P[1, 1], P[1, 2], P[2, 2], v[0], v[1], v[2], float(len(x))])))
        fix = np.zeros(10) if fix10 is None else np.array(fix10[a], dtype=np.float64)
        W = fix.copy()
        for p, c in rows:                                               # covMat = fix_point + sum sig_tran (:656-658)
            Rp = poses12[p][:9].reshape(3, 3).T
            tp = poses12[p][9:12]
            Pm = np.array([[c[0], c[1], c[2]], [c[1], c[3], c[4]], [c[2], c[4], c[5]]])
            Rv = Rp @ c[6:9]
            Pw = Rp @ Pm @ Rp.T + np.outer(Rv, tp) + np.outer(tp, Rv) + c[9] * np.outer(tp, tp)
            W += np.array([Pw[0, 0], Pw[0, 1], Pw[0, 2], Pw[1, 1], Pw[1, 2], Pw[2, 2], *(Rv + c[9] * tp), c[9]])
        cen = W[6:9] / W[9]
        cov = np.array([[W[0], W[1], W[2]], [W[1], W[3], W[4]], [W[2], W[4], W[5]]]) / W[9] - np.outer(cen, cen)
        lam = np.linalg.eigvalsh(cov)
        layer = int(layers[a])
        pts = sum(c[9] for _, c in rows)
        if not (lam[0] / lam[1] < float(np.float32(o["eigen_value_array"][layer]))) or int(pts) <= min_ps or len(rows) < 2:
            continue
        for p, c in rows:
            pi.append(p)
            ob.append(c)
        rp.append(len(pi))
        out_k.append(int(k))
        out_l.append(layer)
        fx.append(fix)
        co.append(float(pts))
    return (np.array(out_k, dtype=np.int64), np.array(out_l, dtype=np.int32), np.array(rp, dtype=np.int64),
            np.array(pi, dtype=np.int32),
            np.array(ob).reshape(-1, 10), np.array(fx).reshape(-1, 10), np.array(co), matched)
