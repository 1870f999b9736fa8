"""Tiny driver for ncu captures: a few fwd+bwd passes at 1080p (planar smooth map)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import pwstablenet_b200 as pw
import synth

N = int(os.environ.get("PROF_N", "8"))
H, W = int(os.environ.get("PROF_H", "1080")), int(os.environ.get("PROF_W", "1920"))
kind = os.environ.get("PROF_KIND", "smooth")
iters = int(os.environ.get("PROF_ITERS", "2"))
g = torch.from_numpy(synth.make_map(kind, 2, H, W, False, seed=1)).cuda()
grid = g.repeat(N // 2, 1, 1, 1).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
frames = torch.rand(N, 3, H, W, device="cuda") * 255
gout = torch.rand(N, 3, H, W, device="cuda")
for _ in range(iters):
    out = pw.warp2d_forward(frames, grid, 0, False)
    gin, gg = pw.warp2d_backward(gout, frames, grid, 0, False, (True, True))
    if os.environ.get("PROF_SPLIT"):
        pw.warp2d_backward(gout, frames, grid, 0, False, (True, False))
        pw.warp2d_backward(gout, frames, grid, 0, False, (False, True))
torch.cuda.synchronize()
print("done", float(out.sum()), float(gin.sum()))
