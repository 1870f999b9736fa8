"""Development: where the host time of one small warp call goes (the 256 x 256 training shapes are bound by it).
cProfile over forward + backward through autograd at batch 2, ours against torch's own op."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import pwstablenet_b200 as pw

x = torch.rand(2, 3, 256, 256, device="cuda", requires_grad=True)
g = (torch.rand(2, 2, 256, 256, device="cuda") * 2 - 1).permute(0, 2, 3, 1).requires_grad_()

def step(fn, n):
    for _ in range(n):
        y = fn(x, g, align_corners=False)
        y.backward(y.detach())
        x.grad = None; g.grad = None

from pwstablenet_b200 import functional as PF
def ctypes_shim(x, g, align_corners=False):
    return PF._Warp2d.apply(x, g, 0, align_corners)
print("compiled shim:", bool(PF.torch_ext()))
for name, fn in (("torch", F.grid_sample), ("ours", pw.grid_sample), ("ours, ctypes shim", ctypes_shim)):
    step(fn, 50); torch.cuda.synchronize()
    t0 = time.perf_counter(); step(fn, 500); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: host {1e6 * (t1 - t0) / 500:.1f} us per forward+backward (GPU drained {1e6 * (t2 - t1):.0f} us after the loop)")
    with torch.no_grad():
        xd, gd = x.detach(), g.detach()
        fn(xd, gd, align_corners=False); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(500): fn(xd, gd, align_corners=False)
        t1 = time.perf_counter(); torch.cuda.synchronize()
        print(f"{name}: host {1e6 * (t1 - t0) / 500:.1f} us per forward alone (no autograd)")
pr = cProfile.Profile(); pr.enable(); step(pw.grid_sample, 300); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
