"""Development: alternating fwd/bwd timing (the bench.py pattern) with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwstablenet_b200 as pw
import synth
N, C, H, W = 16, 3, 1080, 1920
g = torch.from_numpy(synth.make_map("smooth", 4, H, W, False, seed=1)).cuda().repeat(4, 1, 1, 1)
g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
fr = torch.rand(N, C, H, W, device="cuda") * 255
go = torch.rand(N, C, H, W, device="cuda")
for _ in range(5):
    pw.warp2d_forward(fr, g, 0, False); pw.warp2d_backward(go, fr, g, 0, False, (True, True))
torch.cuda.synchronize()
K = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
for k in range(K):
    ev[k][0].record(); pw.warp2d_forward(fr, g, 0, False)
    ev[k][1].record(); pw.warp2d_backward(go, fr, g, 0, False, (True, True))
    ev[k][2].record()
torch.cuda.synchronize()
f = np.mean([e[0].elapsed_time(e[1]) for e in ev]); b = np.mean([e[1].elapsed_time(e[2]) for e in ev])
print(f"alternating: fwd {f:.3f} ms  bwd {b:.3f} ms  -> {N/(f+b)*1e3:.0f} frames/s", flush=True)
