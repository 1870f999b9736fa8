"""Development microbench (not the contract bench): ours vs torch CUDA grid_sample,
fwd and bwd, device-resident inputs larger than L2, CUDA-event timing."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import pwstablenet_b200 as pw
import synth


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    cases = [("1080p", 16, 3, 1080, 1920), ("720p", 32, 3, 720, 1280), ("256", 16, 3, 256, 256)]
    if os.environ.get("QB_ONLY_1080"): cases = cases[:1]
    kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["smooth", "random", "centre"]
    print(torch.cuda.get_device_name(0), "chunkMB", os.environ.get("PWS_BWD_CHUNK_MB", "32"))
    for name, N, C, H, W in cases:
        for kind in kinds:
            nmap = min(N, 4)
            g = torch.from_numpy(synth.make_map(kind, nmap, H, W, False, seed=1)).cuda()
            grid = g.repeat((N + nmap - 1) // nmap, 1, 1, 1)[:N].contiguous()
            planar = grid.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
            frames = torch.rand(N, C, H, W, device="cuda") * 255
            gout = torch.rand(N, C, H, W, device="cuda")
            px = N * H * W
            fb, bb = 32 * px, 52 * px
            for gname, gr in (("interleaved", grid), ("planar", planar)):
                res = {}
                for impl, fwd in ((("ours", pw.warp2d_forward),) if os.environ.get("QB_NO_TORCH") else (("ours", pw.warp2d_forward), ("torch", None))):
                    if impl == "ours":
                        f = lambda: pw.warp2d_forward(frames, gr, 0, False)
                        b = lambda: pw.warp2d_backward(gout, frames, gr, 0, False, (True, True))
                        bg = lambda: pw.warp2d_backward(gout, frames, gr, 0, False, (False, True))
                        bi = lambda: pw.warp2d_backward(gout, frames, gr, 0, False, (True, False))
                    else:
                        f = lambda: torch.ops.aten.grid_sampler_2d(frames, gr, 0, 0, False)
                        b = lambda: torch.ops.aten.grid_sampler_2d_backward(gout, frames, gr, 0, 0, False, [True, True])
                        bg = lambda: torch.ops.aten.grid_sampler_2d_backward(gout, frames, gr, 0, 0, False, [False, True])
                        bi = lambda: torch.ops.aten.grid_sampler_2d_backward(gout, frames, gr, 0, 0, False, [True, False])
                    tf, _ = timeit(f)
                    tb, _ = timeit(b)
                    tbg, _ = timeit(bg)
                    tbi, _ = timeit(bi)
                    res[impl] = (tf, tb)
                    print(f"{name:6s} {kind:7s} {gname:11s} {impl:5s} fwd {tf:8.3f} ms {fb/tf/1e6:7.0f} GB/s | bwd {tb:8.3f} ms {bb/tb/1e6:7.0f} GB/s"
                          f" | bwd-grid-only {tbg:8.3f} | bwd-in-only {tbi:8.3f} | f+b {N/(tf+tb)*1e3:9.0f} frames/s", flush=True)


if __name__ == "__main__":
    main()
