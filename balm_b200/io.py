"""On-disk formats of the reference's real-world benchmark -- host-side readers/writers.

  alidarPose.csv   four comma-separated lines per pose = the ROWS of the 4x4 matrix [R p; 0 0 0 t]
                   (read_pose, src/benchmark/benchmark_realworld.cpp:31-73: 16 numbers fill a column-major Matrix4d
                   which is then transposed; element (3,3) carries the scan time)
  full<i>.pcd      PCL point clouds, loaded as pcl::PointXYZI and copied field by field
                   (read_file, benchmark_realworld.cpp:75-106). The dataset ships "DATA binary" files with
                   FIELDS x y z intensity normal_x normal_y normal_z curvature (8 x float32 = 32 B per point).

Only what the benchmark needs is implemented: ascii and binary PCD (not binary_compressed), float32/float64/integer
fields of COUNT 1. Points are returned as float32 x,y,z -- exactly the values the reference's cut_voxel sees.
"""
import os

import numpy as np

_PCD_TYPES = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32,
              ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32}


def read_pose_csv(path):
    """-> (R [n,3,3], p [n,3], t [n]). A trailing incomplete group of lines is ignored, like the reference's loop."""
    rows = []
    with open(path) as f:
        for ln in f:
            ln = ln.strip()
            if not ln:
                continue
            rows.append([float(x) for x in ln.rstrip(",").split(",") if x.strip() != ""])
    n = len(rows) // 4
    R = np.zeros((n, 3, 3))
    p = np.zeros((n, 3))
    t = np.zeros(n)
    for i in range(n):
        m = np.array([r[:4] for r in rows[4 * i:4 * i + 4]], dtype=np.float64)
        if m.shape != (4, 4):
            raise ValueError(f"{path}: pose {i} does not have four lines of four numbers")
        R[i], p[i], t[i] = m[:3, :3], m[:3, 3], m[3, 3]
    return R, p, t


def write_pose_csv(path, R, p, t=None):
    with open(path, "w") as f:
        for i in range(len(R)):
            m = np.eye(4)
            m[:3, :3], m[:3, 3] = R[i], p[i]
            m[3, 3] = 0.0 if t is None else t[i]
            for r in range(4):
                f.write(",".join(repr(float(x)) for x in m[r]) + "\n")


def read_pcd(path, want_intensity=False):
    """-> xyz float32 [n,3] (and intensity float32 [n] if asked)."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, npts, data = [], [], [], [], None, None
        while True:
            raw = f.readline()
            if not raw:
                raise ValueError(f"{path}: no DATA line")
            ln = raw.decode("ascii", "replace").strip()
            if not ln or ln.startswith("#"):
                continue
            key, _, rest = ln.partition(" ")
            if key == "FIELDS":
                fields = rest.split()
            elif key == "SIZE":
                sizes = [int(x) for x in rest.split()]
            elif key == "TYPE":
                types = rest.split()
            elif key == "COUNT":
                counts = [int(x) for x in rest.split()]
            elif key == "POINTS":
                npts = int(rest)
            elif key == "WIDTH" and npts is None:
                npts = int(rest)  # overwritten by POINTS when present (unorganised clouds: WIDTH == POINTS)
            elif key == "DATA":
                data = rest.strip()
                break
        if not counts:
            counts = [1] * len(fields)
        if any(c != 1 for c in counts):
            raise ValueError(f"{path}: fields with COUNT != 1 are not supported")
        for need in ("x", "y", "z"):
            if need not in fields:
                raise ValueError(f"{path}: no field '{need}'")
        dt = np.dtype([(n_, _PCD_TYPES[(t_, s_)]) for n_, t_, s_ in zip(fields, types, sizes)])
        if data == "binary":
            buf = f.read(npts * dt.itemsize)
            if len(buf) < npts * dt.itemsize:
                raise ValueError(f"{path}: truncated (expected {npts} points)")
            rec = np.frombuffer(buf, dtype=dt, count=npts)
        elif data == "ascii":
            txt = np.loadtxt(f, dtype=np.float64, ndmin=2)
            if len(txt) != npts:
                raise ValueError(f"{path}: {len(txt)} points, header says {npts}")
            rec = np.zeros(npts, dtype=dt)
            for j, n_ in enumerate(fields):
                rec[n_] = txt[:, j]
        else:
            raise ValueError(f"{path}: DATA {data} is not supported (ascii and binary are)")
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float32)
    if want_intensity:
        inten = rec["intensity"].astype(np.float32) if "intensity" in fields else np.zeros(npts, np.float32)
        return xyz, inten
    return xyz


def write_pcd(path, xyz, intensity=None, binary=True):
    """Writes the 8-float layout of the dataset (normals and curvature zero)."""
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = len(xyz)
    rec = np.zeros((n, 8), dtype=np.float32)
    rec[:, :3] = xyz
    if intensity is not None:
        rec[:, 3] = intensity
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
           "FIELDS x y z intensity normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4 4\n"
           "TYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'binary' if binary else 'ascii'}\n")
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if binary:
            f.write(rec.tobytes())
        else:
            for r in rec:
                f.write((" ".join(repr(float(x)) for x in r) + "\n").encode("ascii"))


def read_realworld_dir(prename, max_scans=None):
    """read_file (benchmark_realworld.cpp:75-106): poses from <dir>/alidarPose.csv, scan m from <dir>/full<m>.pcd.
    -> R, p, t, list of float32 xyz arrays."""
    R, p, t = read_pose_csv(os.path.join(prename, "alidarPose.csv"))
    n = len(R) if max_scans is None else min(len(R), int(max_scans))
    scans = [read_pcd(os.path.join(prename, f"full{m}.pcd")) for m in range(n)]
    return R[:n], p[:n], t[:n], scans
