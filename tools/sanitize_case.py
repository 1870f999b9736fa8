"""Development: a small forward + backward through every TMA kernel family, for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pwstablenet_b200 as pw
from pwstablenet_b200 import _lib
import synth
N, C, H, W = 3, 3, 144, 256
for kind in ("smooth", "noisy"):
    g = torch.from_numpy(synth.make_map(kind, N, H, W, False, seed=1)).cuda()
    gp = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    fr = torch.rand(N, C, H, W, device="cuda") * 255
    go = torch.rand(N, C, H, W, device="cuda")
    for grid in (g, gp):
        out = pw.warp2d_forward(fr, grid, 0, False); k1 = _lib.last_kernel()
        cl = fr.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        out_cl = pw.warp2d_forward(cl, grid, 1, True); k2 = _lib.last_kernel()
        out16 = pw.warp2d_forward(fr.bfloat16(), grid, 0, False); k3 = _lib.last_kernel()
        gi, gg = pw.warp2d_backward(go, fr, grid, 0, False, (True, True)); k4 = _lib.last_kernel()
        gi2, _ = pw.warp2d_backward(go, fr, grid, 1, True, (True, False))
        _, gg2 = pw.warp2d_backward(go, fr, grid, 0, False, (False, True))
        torch.cuda.synchronize()
        ref = torch.ops.aten.grid_sampler_2d(fr, grid, 0, 0, False)
        print(kind, k1, k2, k3, k4, bool(torch.equal(out, ref)), float(gi.sum()), flush=True)
print("sanitize case done")
