"""Development: per-launch CUDA-event times in the bench.py pattern (outputs held across iterations)."""
import os, sys, time
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
def run(tag, hold):
    for _ in range(5):
        o = pw.warp2d_forward(fr, g, 0, False); gi, gg = pw.warp2d_backward(go, fr, g, 0, False, (True, True))
    torch.cuda.synchronize()
    K = 20
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    host = []
    for k in range(K):
        t0 = time.perf_counter()
        ev[k][0].record()
        if hold: out = pw.warp2d_forward(fr, g, 0, False)
        else: pw.warp2d_forward(fr, g, 0, False)
        t1 = time.perf_counter()
        ev[k][1].record()
        if hold: gin, ggrid = pw.warp2d_backward(go, fr, g, 0, False, (True, True))
        else: pw.warp2d_backward(go, fr, g, 0, False, (True, True))
        ev[k][2].record()
        host.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    torch.cuda.synchronize()
    print(tag, "fwd", " ".join(f"{e[0].elapsed_time(e[1]):.2f}" for e in ev))
    print(tag, "bwd", " ".join(f"{e[1].elapsed_time(e[2]):.2f}" for e in ev))
    print(tag, "host fwd", " ".join(f"{h[0]:.2f}" for h in host))
    print(tag, "host bwd", " ".join(f"{h[1]:.2f}" for h in host), flush=True)
run("nohold", False)
run("hold  ", True)
run("hold2 ", True)
