"""Two-rank parity check (torchrun): voxel shards + NCCL all-reduce == single GPU evaluation / LM run."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import balm_b200, scenes
from balm_b200 import shard

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for prec in (0, 1):
    sc = scenes.make_scene(n_poses=40, n_planes=400, seed=51, drop=0.2, pts_size=8)
    parts = shard.partition_voxels(sc["row_ptr"], world)
    c = balm_b200.Context(40, lr, prec)
    c.set_voxels(*[x for x in shard.shard_arrays(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"], None, *parts[rank])][:4])
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(balm_b200.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    c.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    H, g, r = c.evaluate(sc["poses_init"])
    poses, tr, _ = c.damping_iter(sc["poses_init"], gauge_mode=2, min_planes_per_pose=0)
    if rank == 0:
        c1 = balm_b200.Context(40, lr, prec)
        c1.set_voxels(sc["row_ptr"], sc["pose_idx"], sc["obs10"], sc["coe"])
        H1, g1, r1 = c1.evaluate(sc["poses_init"])
        p1, tr1, _ = c1.damping_iter(sc["poses_init"], gauge_mode=2, min_planes_per_pose=0)
        eH = np.abs(H - H1).max() / np.abs(H1).max(); eg = np.abs(g - g1).max() / np.abs(g1).max()
        ep = np.abs(poses - p1).max()
        print(f"prec {prec}: relH {eH:.2e} relg {eg:.2e} r {abs(r - r1) / abs(r1):.1e} dpose {ep:.2e} iters {len(tr)} {len(tr1)}")
        ok = ok and eH < (1e-12 if prec == 0 else 2e-8) and eg < 1e-12 and ep < 1e-6 and len(tr) == len(tr1)
    dist.barrier()
if rank == 0:
    print("MGPU_OK" if ok else "MGPU_FAIL")
dist.destroy_process_group()
