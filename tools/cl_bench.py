"""Development: forward on channels-last 1080p frames (inference site) vs NCHW, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwstablenet_b200 as pw
from pwstablenet_b200 import _lib
import synth
N, C, H, W = 16, 3, 1080, 1920
g = torch.from_numpy(synth.make_map("smooth", 4, H, W, False, seed=1)).cuda().repeat(4, 1, 1, 1)
g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
fr = torch.rand(N, C, H, W, device="cuda") * 255
cl = fr.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
def t(fn, k=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
for name, x in (("nchw", fr), ("channels_last", cl)):
    ms = t(lambda: pw.warp2d_forward(x, g, 0, False))
    print(f"{name:14s} ours {ms:.3f} ms ({_lib.last_kernel()})  {32*N*H*W/ms/1e6:.0f} GB/s   aten {t(lambda: torch.ops.aten.grid_sampler_2d(x, g, 0, 0, False)):.3f} ms", flush=True)
