"""Development: a small forward + backward through every TMA kernel family, for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pwstablenet_b200 as pw
from pwstablenet_b200 import _lib
import synth
_lib.small_problem_elems(0)      # reach the persistent TMA kernels with this small case
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
# the multi-map op and the 16-bit backward
fr = torch.rand(N, C, H, W, device="cuda") * 2 - 1
gs = [torch.from_numpy(synth.make_map("smooth", N, H, W, False, seed=s)).cuda().requires_grad_(True) for s in (1, 2, 3)]
outs = pw.warp_stages(fr, gs, pre=(1.0, 127.5), post=(127.5, -1.0))
torch.autograd.backward(list(outs), [torch.ones_like(o) for o in outs])
gi16, gg16 = pw.warp2d_backward(go.bfloat16(), (fr * 100).bfloat16(), gp, 0, False, (True, True))
torch.cuda.synchronize()
print("stages + 16-bit backward", float(gs[0].grad.abs().sum()), float(gi16.float().abs().sum()), flush=True)
print("sanitize case done")
