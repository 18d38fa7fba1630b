"""Voxel sharding for multi-GPU runs: the feature-parallel split of BALM2::divide_thread_left
(/root/reference/src/benchmark/bavoxel.hpp:1044-1047 cuts the voxel list into 4 contiguous ranges of equal
COUNT for its 4 threads). Across GPUs the cut is still contiguous but balanced on the Hessian-accumulation
work, sum of k_v^2 (k_v = poses observing voxel v), with a linear term for the O(K) passes."""
import numpy as np


def partition_voxels(row_ptr, world, quad_weight=1.0, lin_weight=16.0):
    """-> list of (head, end) voxel ranges, one per rank, contiguous, covering [0, M) exactly once."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    M = len(row_ptr) - 1
    k = np.diff(row_ptr).astype(np.float64)
    w = quad_weight * k * k + lin_weight * k
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        c = int(np.searchsorted(cum, target))
        c = min(max(c, cuts[-1]), M)
        cuts.append(c)
    cuts.append(M)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_arrays(row_ptr, pose_idx, obs10, coe, fix10, head, end):
    """Slice the CSR arrays of voxels [head, end) (row_ptr re-based to 0)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    s0, s1 = row_ptr[head], row_ptr[end]
    rp = row_ptr[head:end + 1] - s0
    return (rp, np.asarray(pose_idx)[s0:s1], np.asarray(obs10)[s0:s1], np.asarray(coe)[head:end],
            None if fix10 is None else np.asarray(fix10)[head:end])
