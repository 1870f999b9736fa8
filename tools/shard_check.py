"""Multi-process check of the frame-sharded warp (run under torchrun on N GPUs):
each rank warps its contiguous frame range of a synthetic clip; rank 0 gathers the ranges
and compares them bit for bit with the clip warped on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import pwstablenet_b200 as pw
import synth
from pwstablenet_b200 import sharding


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, C, H, W = 24, 3, 270, 480
    frames = torch.from_numpy(synth.make_frames(n, C, H, W, seed=5)).cuda()
    grid = torch.from_numpy(synth.make_map("smooth", n, H, W, False, seed=6)).cuda()
    sh = sharding.shard_frames(n, world, rank)
    mine = pw.warp2d_forward(frames[sh.begin:sh.end], grid[sh.begin:sh.end], 0, False)
    got = sharding.gather_frames(mine, sh, dst=0)
    if rank == 0:
        full = pw.warp2d_forward(frames, grid, 0, False)
        assert torch.equal(got, full), "sharded result differs from the single-GPU result"
        print(f"shard_check ok: world={world}, {n} frames, bit-identical to the 1-GPU run")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
