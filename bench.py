#!/usr/bin/env python
"""bench.py -- 1080p warped frames/s of the pixel-wise warp hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic frames:
forward warp + backward warp (gradients to BOTH the frame and the map) of
FRAMES 1080p fp32 RGB frames per GPU, through the C ABI of libpwswarp.so.

  value        frames/s, whole job, inputs resident in HBM (inputs are 0.9 GB per GPU,
               larger than the 126 MB L2, so every step streams from DRAM)
  e2e          same metric through the user-facing call (pwstablenet_b200.grid_sample
               pipeline) with PINNED HOST buffers: H2D of frames/map/grad_output and
               D2H of output/grad_frame/grad_map inside the timed region
  roofline     backward kernel (bwd_tma_kernel, the dominant one: 71 % of the step live, 72 % in the ncu launch
               list profiles/r01_launches_bench.csv): algorithmic bytes (52 B/pixel) / CUDA-event time
               of the backward call, against MEASURED_PEAKS.json; `traffic` = DRAM bytes of the same launch from ncu
  cpu_baseline the reference's own CPU path -- torch.nn.functional.grid_sample on CPU
               tensors exactly as R/main_new.py:106,116,716 call it -- timed on this box's
               host cores over a bounded sample
  --impl reference   runs only that CPU path (rank 0) and prints the same JSON line

N > 1 (torchrun): every rank owns its own frames (frame/clip sharding, no collective on
the warp path); the timed region is bracketed by barrier + synchronize and the MAX over
ranks is taken.  scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

H, W, C = 1080, 1920, 3
FRAMES = 16                      # frames per GPU per step
FWD_BYTES_PX, BWD_BYTES_PX = 32, 52   # algorithmic bytes per output pixel, fp32 C=3
NCU_BWD_DRAM_BYTES = 1_162_755_000 + 616_722_000   # ncu --set full, 16-frame backward launch: read + write (profiles/r01c_bwd_tma.txt)
WORKLOAD = ("1080p (1920x1080) fp32 RGB bilinear warp, forward + backward (grad to frame and map), "
            f"{FRAMES} frames/GPU/step, zeros padding, align_corners=False, NCHW frames, planar-stored map "
            "= identity + 0.03*tanh(low-pass noise)")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_inputs(n_frames, seed, device):
    """Frames U(0,255), grad_output U(0,1) on the device; the map is built on the host
    from the seeded generator shared with the tests (tests/synth.py) and stored planar."""
    import synth
    g = torch.Generator(device="cpu").manual_seed(seed)
    nmap = min(n_frames, 4)
    m = torch.from_numpy(synth.make_map("smooth", nmap, H, W, False, seed=seed))
    m = m.repeat((n_frames + nmap - 1) // nmap, 1, 1, 1)[:n_frames]
    planar = m.permute(0, 3, 1, 2).contiguous()          # (N,2,H,W) storage, as netG returns it
    frames = torch.rand((n_frames, C, H, W), generator=g) * 255
    gout = torch.rand((n_frames, C, H, W), generator=g)
    if device is not None:
        frames, gout, planar = frames.to(device), gout.to(device), planar.to(device)
    return frames, planar.permute(0, 2, 3, 1), gout       # map view: strides (2HW, W, 1, HW)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        if os.environ.get("PWS_BENCH_NO_SMI"):
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: take the nearest samples
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_pass(frames, grid, gout, threads):
    """The reference's CPU path: F.grid_sample + autograd backward on CPU tensors."""
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    fi = frames.clone().requires_grad_(True)
    gi = grid.clone().requires_grad_(True)
    t0 = time.perf_counter()
    out = F.grid_sample(fi, gi, mode="bilinear", padding_mode="zeros", align_corners=False)
    out.backward(gout)
    return time.perf_counter() - t0


def cpu_baseline(budget_s=20.0):
    """Bounded sample of the bench workload on the host cores. ATen parallelises this op over
    the batch only, so the sample batch equals the thread count."""
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    frames, grid, gout = make_inputs(threads, seed=7, device=None)
    cpu_reference_pass(frames[:1], grid[:1], gout[:1], threads)       # warm the op
    t, passes, spent = [], 0, 0.0
    while passes < 3 and (spent < budget_s or passes == 0):
        dt = cpu_reference_pass(frames, grid, gout, threads)
        t.append(dt); spent += dt; passes += 1
    best = min(t)
    return {"value": threads / best, "unit": "frames/s", "cores": threads, "kind": "reference",
            "sample": f"torch.nn.functional.grid_sample fwd+bwd on CPU tensors (ATen CPU kernel, the op the reference's "
                      f"call sites run), batch {threads} 1080p fp32 frames, best of {passes} passes, "
                      f"{threads} threads of {cores} host cores",
            "ms_per_pass": best * 1e3}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    frames, grid, gout = make_inputs(threads, seed=7, device=None)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(frames, grid, gout, threads)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(frames, grid, gout, threads)
    dt = (time.perf_counter() - t0) / steps
    value = threads / dt
    line = {
        "impl": "reference", "metric": "1080p warped frames/s (forward+backward)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": threads,
                   "note": "reference CPU path on host cores; each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "reference",
                         "sample": f"torch CPU grid_sample fwd+bwd, batch {threads} 1080p frames per step"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import _lib
    _lib.load()  # fails loudly when the CUDA library is missing: there is no fallback

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)

    frames, grid, gout = make_inputs(FRAMES, seed=100 + rank, device=dev)
    px = FRAMES * H * W

    def step_device():
        out = pw.warp2d_forward(frames, grid, 0, False)
        gin, ggrid = pw.warp2d_backward(gout, frames, grid, 0, False, (True, True))
        return out, gin, ggrid

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) + per-call CUDA-event split
    # warm up in the same pattern as the timed loop (results held until reassigned), so the caching allocator has
    # reached its steady state and no cudaMalloc lands inside the timed region
    for _ in range(warmup):
        out, gin, ggrid = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    l0 = _lib.launch_count()
    barrier()
    t_wall0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(steps):
        ev[k][0].record()
        out = pw.warp2d_forward(frames, grid, 0, False)
        ev[k][1].record()
        gin, ggrid = pw.warp2d_backward(gout, frames, grid, 0, False, (True, True))
        ev[k][2].record()
    stop.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1)
    elapsed_ms = start.elapsed_time(stop)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))

    # ---------------- end to end through the public API with host buffers
    hf, hg, hgo = (t.cpu().pin_memory() for t in (frames, grid.permute(0, 3, 1, 2).contiguous(), gout))
    h_out = torch.empty_like(hf).pin_memory()
    h_gin = torch.empty_like(hf).pin_memory()
    h_gg = torch.empty_like(hg).pin_memory()
    del out, gin, ggrid
    h2d = hf.numel() * 4 + hg.numel() * 4 + hgo.numel() * 4
    d2h = h_out.numel() * 4 + h_gin.numel() * 4 + h_gg.numel() * 4

    # the user-facing host API: chunks of 2 frames, upload / warp / download overlapped on three streams
    pipe = pw.HostWarpPipeline(2, C, (H, W), device=dev, backward=True)

    def step_e2e():
        pipe.run(hf, hg, h_out, hgo, h_gin, h_gg)

    e2e_steps = max(3, min(steps, 10))
    for _ in range(2):
        step_e2e()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(e2e_steps):
        step_e2e()
    e2.record()
    barrier()
    e2e_ms = s2.elapsed_time(e2)

    # ---------------- max over ranks
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, fwd_ms, bwd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, fwd_ms, bwd_ms = [float(x) for x in t.tolist()]
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    if rank == 0:
        peak, peak_src = peaks()
        value = world * FRAMES * steps / (elapsed_ms * 1e-3)
        bwd_gbs = BWD_BYTES_PX * px / (bwd_ms * 1e-3) / 1e9
        fwd_gbs = FWD_BYTES_PX * px / (fwd_ms * 1e-3) / 1e9
        step_gbs = (FWD_BYTES_PX + BWD_BYTES_PX) * px * steps / (elapsed_ms * 1e-3) / 1e9
        line = {
            "metric": "1080p warped frames/s (forward+backward)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": FRAMES, "global_frames_per_step": FRAMES * world,
                       "sharding": f"frames split over {world} rank(s), no collective on the warp path",
                       "l2": "inputs (0.9 GB/GPU) exceed the 126 MB L2; no explicit flush needed"},
            "clocks": clocks,
            "e2e": {"value": world * FRAMES * e2e_steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "pwstablenet_b200.HostWarpPipeline.run(): pinned host buffers in and out, 2-frame chunks, H2D / fwd+bwd through the C ABI / D2H overlapped on 3 streams"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "bwd_tma_kernel (pws_warp2d_backward)", "achieved": bwd_gbs, "peak": peak,
                         "unit": "GB/s", "frac": bwd_gbs / peak, "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one 16-frame launch (profiles/r01c_bwd_tma.txt)
                         "traffic": NCU_BWD_DRAM_BYTES if FRAMES == 16 else None,
                         "algorithmic_bytes_per_launch_set": BWD_BYTES_PX * px, "ms": bwd_ms,
                         "frac_of_8TBs_nominal": bwd_gbs / 8000.0,
                         "forward": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "ms": fwd_ms},
                         "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak}},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
