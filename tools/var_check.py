"""Development: parity against ATen's CUDA kernel + alternating fwd/bwd timing at the bench size, for whichever
library PWS_LIB_PATH selects (tools/var_bench.sh runs it once per variant)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwstablenet_b200 as pw
import synth

def check(N, C, H, W, kind, pad, align, inter=False, seed=1):
    g = torch.from_numpy(synth.make_map(kind, N, H, W, align, seed=seed)).cuda()
    if not inter:
        g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    fr = torch.rand(N, C, H, W, device="cuda") * 255
    go = torch.rand(N, C, H, W, device="cuda")
    gin, gg = pw.warp2d_backward(go, fr, g, pad, align, (True, True))
    rin, rg = torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, pad, align, (True, True))
    e_in = ((gin - rin).abs().max() / rin.abs().max()).item()
    e_g = ((gg - rg).abs().max() / rg.abs().max()).item()
    gin1, _ = pw.warp2d_backward(go, fr, g, pad, align, (True, False))
    _, gg1 = pw.warp2d_backward(go, fr, g, pad, align, (False, True))
    e1 = ((gin1 - rin).abs().max() / rin.abs().max()).item()
    same = torch.equal(gg1, gg)
    ok = e_in <= 1e-4 and e_g <= 1e-6 and e1 <= 1e-4 and same
    print(f"  {'ok ' if ok else 'BAD'} {N}x{C}x{H}x{W} {kind:8s} pad {pad} align {int(align)} inter {int(inter)}: gin {e_in:.1e} ggrid {e_g:.1e} gin-only {e1:.1e} ggrid-only same {same} [{pw._lib.last_kernel()}]", flush=True)
    return ok

ok = True
for kind in ("smooth", "noisy", "random", "centre", "identity"):
    for pad in (0, 1):
        for align in (False, True):
            ok &= check(2, 3, 270, 480, kind, pad, align, inter=(pad == 1))
ok &= check(3, 1, 250, 388, "smooth", 0, False)
ok &= check(2, 3, 1080, 1920, "smooth", 0, False)
print("PARITY", "OK" if ok else "FAILED", flush=True)

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
for mask in ((True, False), (False, True)):
    for _ in range(3):
        pw.warp2d_backward(go, fr, g, 0, False, mask)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        pw.warp2d_backward(go, fr, g, 0, False, mask)
    e.record(); torch.cuda.synchronize()
    print(f"  mask {mask}: {s.elapsed_time(e)/K:.3f} ms", flush=True)
