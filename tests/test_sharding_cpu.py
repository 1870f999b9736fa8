"""Host-side multi-GPU logic on CPU: shard arithmetic and a world_size-2 gloo run of the
gather plumbing.  The warp path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pwstablenet_b200 import sharding


@pytest.mark.parametrize("frames,world", [(300, 1), (300, 2), (300, 4), (300, 8), (7, 8), (0, 4), (31, 3)])
def test_shards_partition_the_clip(frames, world):
    shards = sharding.all_shards(frames, world)
    assert shards[0].begin == 0 and shards[-1].end == frames
    for a, b in zip(shards, shards[1:]):
        assert a.end == b.begin
    sizes = [s.count for s in shards]
    assert sum(sizes) == frames and max(sizes) - min(sizes) <= 1
    for s in shards:
        if s.count:
            assert s.halo_begin == max(0, s.begin - 15) and s.halo_end == min(frames, s.end + 15)
            # every input frame netG needs for an owned output lies inside the halo range
            for f in (s.begin, s.end - 1):
                idx = sharding.window_indices(f, frames)
                assert len(idx) == 31 and min(idx) >= s.halo_begin and max(idx) < s.halo_end


def test_batch_sharding_matches_reference_batch():
    parts = [sharding.shard_batch(16, 8, r) for r in range(8)]
    assert [p.stop - p.start for p in parts] == [2] * 8
    assert sharding.shard_batch(16, 1, 0) == slice(0, 16)
    with pytest.raises(ValueError):
        sharding.shard_frames(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = sharding.shard_frames(frames, world, rank)
        # stand-in for "warp my frames": frame i becomes a tensor filled with i
        local = torch.stack([torch.full((2, 3), float(i)) for i in range(sh.begin, sh.end)]) if sh.count else torch.empty(0, 2, 3)
        got = sharding.gather_frames(local, sh, dst=0)
        if rank == 0:
            torch.save(got, out)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_gather_is_in_clip_order(tmp_path):
    out = str(tmp_path / "g.pt")
    frames = 11
    mp.spawn(_worker, args=(2, _free_port(), frames, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got.shape == (frames, 2, 3)
    assert torch.equal(got[:, 0, 0], torch.arange(frames, dtype=torch.float32))
