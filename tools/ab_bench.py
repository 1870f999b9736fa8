"""Development: backward timings of whichever pwstablenet_b200 package directory is given (default: this tree), for A/B
runs against the round-1 build kept under tools/exp/r01pkg (git-ignored).  usage: python tools/ab_bench.py [PKG_PARENT_DIR]"""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, sys.argv[1] if len(sys.argv) > 1 else root)
sys.path.insert(0, os.path.join(root, "tests"))
import torch
import pwstablenet_b200 as pw
import synth
print("package:", os.path.dirname(pw.__file__), flush=True)


def t(fn, k=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k


def planar(g):
    return g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)


for (N, H, W, reps) in ((64, 720, 1280, 10), (16, 1080, 1920, 20), (16, 256, 256, 100), (64, 256, 256, 50)):
    nm = min(N, 4)
    g = planar(torch.from_numpy(synth.make_map("smooth", nm, H, W, False, seed=1)).cuda().repeat((N + nm - 1) // nm, 1, 1, 1)[:N].contiguous())
    fr = torch.rand(N, 3, H, W, device="cuda") * 255
    go = torch.rand(N, 3, H, W, device="cuda")
    out = torch.empty_like(fr); gin = torch.empty_like(fr); gg = torch.empty_strided(g.shape, g.stride(), device="cuda")
    f = t(lambda: pw.warp2d_forward(fr, g, 0, False, out=out), reps)
    b = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, True), grad_input=gin, grad_grid=gg), reps)
    b1 = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, False), grad_input=gin), reps)
    b2 = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (False, True), grad_grid=gg), reps)
    print(f"{N}x3x{H}x{W}: fwd {f:.3f}  bwd both {b:.3f}  gin-only {b1:.3f}  ggrid-only {b2:.3f} ms", flush=True)
    del fr, go, g, out, gin, gg
