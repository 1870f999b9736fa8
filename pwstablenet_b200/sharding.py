"""Frame / clip sharding of the warp path across the GPUs of one box.

The warp has no cross-frame dependence: output frame i needs its own frame and its own map
(R/main_new.py:716), and the map needs input frames i-15..i+15 of the 31-frame gray window
(period = 30, R/lib/cfg.py:4; R/main_new.py:642-673; feeding stabilised frames back is
commented out at :712-715).  So a clip is cut into contiguous frame ranges, one per rank,
each with a read-only halo of `period // 2` INPUT frames on both sides for netG; the warp
itself needs no halo and no collective.  One process per GPU; torch.distributed is used
only to agree on the split and, optionally, to collect results on rank 0.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

PERIOD = 30  # R/lib/cfg.py:4


@dataclass(frozen=True)
class FrameShard:
    rank: int
    world: int
    begin: int        # first output frame owned by this rank
    end: int          # one past the last owned output frame
    halo_begin: int   # first input frame netG needs for `begin` (clamped to the clip)
    halo_end: int     # one past the last input frame netG needs for `end - 1`

    @property
    def count(self) -> int:
        return self.end - self.begin


def shard_frames(num_frames: int, world: int, rank: int, halo: int = PERIOD // 2) -> FrameShard:
    """Contiguous split of `num_frames` over `world` ranks: the first `num_frames % world`
    ranks get one frame more (sizes differ by at most 1; empty shards when world > frames)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if num_frames < 0:
        raise ValueError("num_frames must be >= 0")
    base, extra = divmod(num_frames, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return FrameShard(rank, world, begin, end, max(0, begin - halo) if end > begin else begin,
                      min(num_frames, end + halo) if end > begin else begin)


def all_shards(num_frames: int, world: int, halo: int = PERIOD // 2) -> List[FrameShard]:
    return [shard_frames(num_frames, world, r, halo) for r in range(world)]


def shard_batch(batch: int, world: int, rank: int) -> slice:
    """Batch-dimension split used by a data-parallel training step (per-GPU batch 16/n)."""
    s = shard_frames(batch, world, rank, halo=0)
    return slice(s.begin, s.end)


def window_indices(frame: int, num_frames: int, period: int = PERIOD) -> List[int]:
    """Indices of the `period + 1` input frames netG sees for output `frame`, with the
    reference's edge replication (R/main_new.py:612-633 pre-fills the history with the first
    frame; the tail repeats the last one)."""
    half = period // 2
    return [min(max(frame - half + k, 0), num_frames - 1) for k in range(period + 1)]


def gather_frames(local, shard: FrameShard, group=None, dst: int = 0):
    """Collect per-rank results (a tensor whose dim 0 is the shard's frames) on rank `dst`
    in clip order.  Pure plumbing for writing the video out; the warp path never calls it."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or shard.world == 1:
        return local
    sizes = [s.count for s in all_shards_from(shard, group)]
    rank = dist.get_rank(group)
    most = max(sizes)
    # dist.gather wants equal shapes: pad every shard to the largest one, trim after
    padded = torch.zeros((most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in sizes] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0) if rank == dst else None


def all_shards_from(shard: FrameShard, group=None) -> Sequence[FrameShard]:
    import torch
    import torch.distributed as dist

    total = torch.tensor([shard.count], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        total = total.cuda()
    dist.all_reduce(total, group=group)
    return all_shards(int(total.item()), shard.world)
