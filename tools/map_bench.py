"""Development: backward time at the bench size against the shape of the map -- a pure sub-pixel shift (every warp row
reads one box row, hands every east tap over, queues nothing: the floor of the kernel's design), uniform zooms, the bench
map.  Whichever library PWS_LIB_PATH selects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwstablenet_b200 as pw
import synth
N, C, H, W = 16, 3, 1080, 1920

def ident():
    ys = (2 * torch.arange(H, dtype=torch.float64) + 1) / H - 1
    xs = (2 * torch.arange(W, dtype=torch.float64) + 1) / W - 1
    return xs.view(1, 1, W).expand(1, H, W), ys.view(1, H, 1).expand(1, H, W)

def planar(x, y):
    m = torch.stack([x, y], 1).float().cuda().repeat(N, 1, 1, 1).contiguous()     # (N, 2, H, W) storage
    return m.permute(0, 2, 3, 1)

def maps():
    x, y = ident()
    yield "shift (+10.37, +3.61) px", planar(x + 2 * 10.37 / W, y + 2 * 3.61 / H)
    yield "zoom 0.99 (1 % compression)", planar(x * 0.99, y * 0.99)
    yield "zoom 0.97", planar(x * 0.97, y * 0.97)
    yield "rotation 1 degree", planar(*rot(x, y, 1.0))
    yield "rotation 3 degrees", planar(*rot(x, y, 3.0))
    g = torch.from_numpy(synth.make_map("smooth", 4, H, W, False, seed=1)).cuda().repeat(4, 1, 1, 1)
    yield "bench map", g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)

def rot(x, y, deg):
    a = np.deg2rad(deg); c, s = np.cos(a), np.sin(a)
    px, py = x * W / 2, y * H / 2
    return (c * px - s * py) * 0.98 / (W / 2), (s * px + c * py) * 0.98 / (H / 2)

DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[os.environ.get("MAP_BENCH_DTYPE", "f32")]   # frames and grad_output
fr = (torch.rand(N, C, H, W, device="cuda") * 255).to(DT)
go = torch.rand(N, C, H, W, device="cuda").to(DT)
K = 20
GI = torch.empty(N, C, H, W, device="cuda", dtype=torch.float32)   # caller-provided fp32 grad_input: the kernel alone, no rounding pass
for name, g in maps():
    row = []
    for mask in ((True, True), (True, False), (False, True)):
        for _ in range(3):
            pw.warp2d_backward(go, fr, g, 0, False, mask, grad_input=GI if mask[0] else None)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(K):
            pw.warp2d_backward(go, fr, g, 0, False, mask, grad_input=GI if mask[0] else None)
        e.record(); torch.cuda.synchronize()
        row.append(s.elapsed_time(e) / K)
    for _ in range(3):
        pw.warp2d_forward(fr, g, 0, False)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        pw.warp2d_forward(fr, g, 0, False)
    e.record(); torch.cuda.synchronize()
    print(f"{name:32s} bwd both {row[0]:.3f}  gin-only {row[1]:.3f}  ggrid-only {row[2]:.3f}  fwd {s.elapsed_time(e)/K:.3f} ms [{pw._lib.last_kernel()}]", flush=True)
