"""BASELINE config 4: the reference's process() loop (R/main_new.py:612-737) on a synthetic 300-frame 1080p clip,
frame-sharded over the GPUs of one box.

Per rank: its contiguous frame range plus the 15-frame input halo netG's window needs (pwstablenet_b200.sharding);
gray 256 x 256 windows built on the device in batches (pwstablenet_b200.windows); netG in eval mode, batched; then the
inference site itself as ONE fused kernel per batch -- uint8 HWC frame in, the 256 x 256 map upsampled inside the sampler
(UpsamplingBilinear2d, :706-710), bilinear sample (:716), truncation to uint8 (:717-721), uint8 HWC frame out.
No collective on the path; rank 0 gathers the stabilised frames only to compare them with the 1-GPU result.

What is checked: the frames every rank produces are bit-identical to the frames rank 0 produces for the same indices
when it processes the whole clip alone.  What is reported: warp-only frames/s (fused kernel, CUDA events), end-to-end
frames/s per rank and in aggregate (host uint8 frames in -> host uint8 frames out, netG included), the share of netG.
"""
from __future__ import annotations

import json
import os
import sys
import time

import torch

H, W = 1080, 1920
CLIP = 300
BATCH = 12      # output frames per netG / warp batch


def synthetic_clip(n, seed=0):
    """Deterministic moving-texture clip, uint8 HWC on the HOST (what cv2.VideoCapture would hand over)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randint(0, 256, (H + 64, W + 64, 3), dtype=torch.uint8, generator=g)
    jit = torch.randint(0, 48, (n, 2), generator=g)
    return torch.stack([base[int(jit[i, 0]):int(jit[i, 0]) + H, int(jit[i, 1]):int(jit[i, 1]) + W] for i in range(n)], 0)


def stabilise_range(netG, clip_host, begin, end, dev, timers):
    """Stabilised frames [begin, end) of the clip as a uint8 HWC device tensor."""
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import sharding, windows
    n = clip_host.size(0)
    half = sharding.PERIOD // 2
    lo, hi = max(0, begin - half), min(n, end + half)
    t0 = time.perf_counter()
    frames = clip_host[lo:hi].to(dev, non_blocking=True)                 # the shard and its halo: uint8 HWC
    gray = windows.gray_small(frames)                                    # (hi-lo, 256, 256) in [-1,1]
    outs = []
    ev = []
    for b in range(begin, end, BATCH):
        e = min(end, b + BATCH)
        win = windows.window_batch(gray, lo, b, e, n)                    # (B,31,256,256)
        with torch.no_grad():
            lattice = netG(win, False)                                   # (B,256,256,2): planar storage, permuted view
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pw.warp_fused(frames[b - lo:e - lo].permute(0, 3, 1, 2), drift=lattice, upsample="aligned", out_size=(H, W),
                            out_dtype=torch.uint8, out_channels_last=True)
        e1.record()
        ev.append((e0, e1, e - b))
        outs.append(out.permute(0, 2, 3, 1))
    res = torch.cat(outs, 0)
    host = torch.empty(res.shape, dtype=torch.uint8).pin_memory()
    host.copy_(res, non_blocking=True)
    torch.cuda.synchronize()
    timers["e2e_s"] = time.perf_counter() - t0
    timers["warp_ms"] = sum(a.elapsed_time(b) for a, b, _ in ev)
    timers["frames"] = end - begin
    return res, host


def run(args):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
    import pwstablenet_b200 as pw  # noqa: F401
    from pwstablenet_b200 import numa, sharding
    from harness.netg_standin import build_netg

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = numa.bind_to_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(123)
    netG = build_netg().to(dev).eval()
    with torch.no_grad():
        # random-init netG emits the degenerate everything-to-the-centre map; a near-identity affine head makes
        # the synthetic weights produce a stabilisation-like warp (identity + small drift), as a trained netG does
        netG.linear.bias.copy_(torch.tensor([1.0, 0.01, 0.0, -0.01, 1.0, 0.0]))
    clip = synthetic_clip(CLIP, seed=0).pin_memory()
    # Shard in whole netG batches: every batch then has the same composition and size whichever rank runs it, so cuDNN makes
    # the same calls and the maps -- hence the frames -- are bit-identical to the 1-GPU run (autotuning off for the same reason)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    nb = (CLIP + BATCH - 1) // BATCH
    bs = sharding.shard_frames(nb, world, rank, halo=0)
    sh = sharding.FrameShard(rank, world, min(CLIP, bs.begin * BATCH), min(CLIP, bs.end * BATCH), 0, 0)

    t_warm = {}
    stabilise_range(netG, clip, sh.begin, min(sh.end, sh.begin + BATCH), dev, t_warm)       # warm-up (cuDNN autotune, allocator)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    timers = {}
    mine, _ = stabilise_range(netG, clip, sh.begin, sh.end, dev, timers)

    # bit-identity with the 1-GPU run: rank 0 recomputes every other rank's range alone and compares
    identical = True
    if world > 1:
        counts = [min(CLIP, b.end * BATCH) - min(CLIP, b.begin * BATCH) for b in sharding.all_shards(nb, world, halo=0)]
        most = max(counts)
        padded = torch.zeros((most,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)
        padded[:mine.shape[0]] = mine
        bufs = [torch.empty_like(padded) for _ in counts] if rank == 0 else None
        dist.gather(padded, bufs, dst=0)
        if rank == 0:
            got = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
            t1 = {}
            full, _ = stabilise_range(netG, clip, 0, CLIP, dev, t1)
            identical = bool(torch.equal(got, full))
    stats = torch.tensor([timers["e2e_s"], timers["warp_ms"], float(timers["frames"])], device=dev, dtype=torch.float64)
    if world > 1:
        alls = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(alls, stats)
    else:
        alls = [stats]
    if rank == 0:
        e2e_s = max(float(a[0]) for a in alls)
        warp_ms = max(float(a[1]) for a in alls)
        px_bytes = 6 * H * W + 2 * 256 * 256 * 4                                          # uint8 in + out + the map lattice
        peak = 6535.7
        try:
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        warp_fps = CLIP / (warp_ms * 1e-3)                                               # ranks run concurrently: max over ranks
        line = {"config": "clip", "metric": "1080p stabilised frames/s (300-frame clip, frame-sharded)",
                "value": CLIP / e2e_s, "unit": "frames/s", "n_gpus": world, "higher_is_better": True, "scaling": "strong",
                "dtype": "u8 frames, f32 maps", "data": "synthetic",
                "frames": CLIP, "frames_per_rank": [int(a[2]) for a in alls],
                "e2e_s_per_rank": [float(a[0]) for a in alls], "warp_ms_per_rank": [float(a[1]) for a in alls],
                "warp_only_frames_per_s": warp_fps,
                "warp_roofline": {"bytes_per_frame": px_bytes, "achieved_GBs": warp_fps * px_bytes / 1e9 / world, "peak": peak,
                                  "frac": warp_fps * px_bytes / 1e9 / world / peak, "kernel": "fwd_fused_u8 (pws_warp2d_forward_fused)"},
                "netg": "stand-in, eval mode, batch %d" % BATCH, "bit_identical_to_1gpu": identical, "host_placement": placement,
                "what": "host uint8 HWC clip -> H2D (shard + 15-frame halo) -> gray 256x256 windows on device -> netG(eval) -> fused "
                        "upsample+sample+uint8 kernel -> D2H; per-rank wall clock, max over ranks"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
