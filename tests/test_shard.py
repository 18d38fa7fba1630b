"""Host-side logic of the multi-GPU path on CPU: the voxel partition and, with two gloo ranks, the
shard -> partial (H, g, r) -> all-reduce(sum) data flow that the library performs with NCCL on GPUs
(= divide_thread_left's `Hess += hessians[i]` reduction, bavoxel.hpp:1049-1056). The partial results here come
from the CPU oracle (test-only), so the test pins the sharding/reduction semantics, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
from balm_b200 import shard
from oracle import oracle_py as orc


def test_partition_covers_every_voxel_once_and_balances_work():
    sc = scenes.make_scene(n_poses=12, n_planes=101, seed=31, drop=0.5, pts_size=5)
    for world in (1, 2, 3, 8):
        parts = shard.partition_voxels(sc["row_ptr"], world)
        assert parts[0][0] == 0 and parts[-1][1] == 101
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        k = np.diff(sc["row_ptr"]).astype(float)
        w = k * k + 16 * k
        loads = [w[h:e].sum() for h, e in parts]
        assert max(loads) <= w.sum() / world + w.max() + 1e-9
    # fewer voxels than ranks: still contiguous and covering, the surplus ranks get EMPTY shards -- which the library
    # registers (balm_set_voxels with n_voxels = 0) so that those ranks keep joining the collectives
    for M, world in ((1, 4), (3, 8), (0, 2)):
        parts = shard.partition_voxels(np.arange(M + 1) * 3, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == M
        assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(parts, parts[1:] + [(M, M)]))
        assert sum(e - h for h, e in parts) == M and sum(e > h for h, e in parts) == min(M, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.make_scene(n_poses=9, n_planes=60, seed=32, drop=0.3, pts_size=5)
    head, end = shard.partition_voxels(sc["row_ptr"], world)[rank]
    rp, pi, ob, co, fx = shard.shard_arrays(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], None, head, end)
    n = 6 * sc["n_poses"]
    if end > head:
        o = orc.Oracle(sc["n_poses"], rp, pi, ob, co)
        H, g, r = o.evaluate(sc["poses_init"])
    else:
        H, g, r = np.zeros((n, n)), np.zeros(n), 0.0
    buf = torch.from_numpy(np.concatenate([np.asarray(H).reshape(-1), g, [r]]))  # [H | g | r], one all-reduce
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = scenes.make_scene(n_poses=9, n_planes=60, seed=32, drop=0.3, pts_size=5)
    o = orc.Oracle(sc["n_poses"], sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
    H, g, r = o.evaluate(sc["poses_init"])
    n = 6 * sc["n_poses"]
    assert np.abs(out[:n * n].reshape(n, n) - H).max() <= 1e-12 * np.abs(H).max()
    assert np.abs(out[n * n:n * n + n] - g).max() <= 1e-12 * np.abs(g).max()
    assert abs(out[-1] - r) <= 1e-13 * abs(r)
