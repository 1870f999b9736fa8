"""Development: one measured line per BASELINE.json config shape (warp only), CUDA events, inputs larger than L2
where the config allows.  Prints a markdown table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import torch.nn.functional as F
import pwstablenet_b200 as pw
from pwstablenet_b200 import _lib
import synth

PEAK = 6535.7


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


def row(name, ms, bytes_, frames, kern, aten_ms=None):
    gbs = bytes_ / ms / 1e6
    extra = f" | {aten_ms:.3f} |" if aten_ms is not None else " | |"
    print(f"| {name} | {ms:.3f} | {frames / ms * 1e3:,.0f} | {gbs:,.0f} | {gbs / PEAK:.2f} | {kern}{extra}", flush=True)


def maps(kind, N, H, W, align=False):
    nm = min(N, 4)
    g = torch.from_numpy(synth.make_map(kind, nm, H, W, align, seed=1)).cuda()
    return planar(g.repeat((N + nm - 1) // nm, 1, 1, 1)[:N].contiguous())


print("| case | ms | frames/s | GB/s (algorithmic) | of 6 535.7 | kernel | ATen CUDA ms |")
print("|---|---|---|---|---|---|---|")
# config 2: batch 64, 3x720x1280 fp32, forward + backward, three map kinds
N, C, H, W = 64, 3, 720, 1280
fr = torch.rand(N, C, H, W, device="cuda") * 255
go = torch.rand(N, C, H, W, device="cuda")
for kind in ("smooth", "random", "centre"):
    g = maps(kind, N, H, W)
    px = N * H * W
    ms = t(lambda: pw.warp2d_forward(fr, g, 0, False)); k = _lib.last_kernel()
    row(f"config 2: 64x3x720x1280 fp32 forward, {kind} map", ms, 32 * px, N, k, t(lambda: torch.ops.aten.grid_sampler_2d(fr, g, 0, 0, False), 3, 1))
    ms = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, True))); k = _lib.last_kernel()
    row(f"config 2: backward (both gradients), {kind} map", ms, 52 * px, N, k, t(lambda: torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, [True, True]), 3, 1))
del fr, go, g
# config 3: the training-step warps, 16x3x256x256 (L2-resident: 12.6 MB forward working set)
N, C, H, W = 16, 3, 256, 256
fr = torch.rand(N, C, H, W, device="cuda") * 255
go = torch.rand(N, C, H, W, device="cuda")
g = maps("smooth", N, H, W)
px = N * H * W
ms = t(lambda: pw.warp2d_forward(fr, g, 0, False), 50); k = _lib.last_kernel()
row("config 3: 16x3x256x256 forward (L2-resident)", ms, 32 * px, N, k, t(lambda: torch.ops.aten.grid_sampler_2d(fr, g, 0, 0, False), 50))
ms = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (False, True)), 50); k = _lib.last_kernel()
row("config 3: backward, grad to map only (main_new.py:106)", ms, 40 * px, N, k, t(lambda: torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, [False, True]), 50))
ms = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, False)), 50); k = _lib.last_kernel()
row("config 3: backward, grad to frame only (main_new.py:197)", ms, 32 * px, N, k, t(lambda: torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, [True, False]), 50))
del fr, go, g
# config 4: 1080p inference, uint8 HWC frames, 256^2 netG map upsampled in the kernel, uint8 out
N, H, W = 16, 1080, 1920
hwc = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")
ident = F.affine_grid(torch.tensor([[[1.0, 0, 0], [0, 1.0, 0]]], device="cuda"), (1, 3, 256, 256), align_corners=False)
drift = (torch.from_numpy(synth.make_map("smooth", 4, 256, 256, False, seed=1)).cuda() - ident).repeat(4, 1, 1, 1).permute(0, 3, 1, 2).contiguous()  # netG-like: +-0.03, low-pass
theta = torch.tensor([[[1.0, 0.002, 0.0], [-0.002, 1.0, 0.0]]], device="cuda").repeat(N, 1, 1)
fused = lambda: pw.warp_fused(hwc.permute(0, 3, 1, 2), drift=drift.permute(0, 2, 3, 1), base="affine", theta=theta,
                              upsample="aligned", out_size=(H, W), out_dtype=torch.uint8, out_channels_last=True)
def unfused():
    now = hwc.float().permute(0, 3, 1, 2)
    grid = drift.permute(0, 2, 3, 1) + F.affine_grid(theta, (N, 3, 256, 256), align_corners=False)
    gr = torch.nn.UpsamplingBilinear2d(size=(H, W))(grid.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    return torch.ops.aten.grid_sampler_2d(now, gr, 0, 0, False).permute(0, 2, 3, 1).to(torch.uint8)
ms = t(fused); k = _lib.last_kernel()
row("config 4: 16 x 1080p inference, uint8 HWC in/out, map composed + upsampled in the kernel (6 B/px)", ms, 6 * N * H * W, N, k, t(unfused, 3, 1))
def ours_unfused():
    now = hwc.float().permute(0, 3, 1, 2)
    grid = drift.permute(0, 2, 3, 1) + F.affine_grid(theta, (N, 3, 256, 256), align_corners=False)
    gr = torch.nn.UpsamplingBilinear2d(size=(H, W))(grid.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    return pw.warp2d_forward(now, gr, 0, False)
fl = hwc.float().permute(0, 3, 1, 2)
gr = torch.nn.UpsamplingBilinear2d(size=(H, W))((drift.permute(0, 2, 3, 1) + F.affine_grid(theta, (N, 3, 256, 256), align_corners=False)).permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
ms = t(lambda: pw.warp2d_forward(fl, gr, 0, False)); k = _lib.last_kernel()
row("config 4: the plain sample of the same call (fp32 channels-last frame view, upsampled map view), 32 B/px", ms, 32 * N * H * W, N, k, t(lambda: torch.ops.aten.grid_sampler_2d(fl, gr, 0, 0, False), 3, 1))
del hwc, fl, gr
# config 5: 4K bf16 frames, fp32 maps, zeros and border
N, C, H, W = 8, 3, 2160, 3840
fr = (torch.rand(N, C, H, W, device="cuda") * 255).to(torch.bfloat16)
g = maps("smooth", N, H, W)
for pad, pname in ((0, "zeros"), (1, "border")):
    ms = t(lambda: pw.warp2d_forward(fr, g, pad, False), 5, 2); k = _lib.last_kernel()
    row(f"config 5: 8x3x2160x3840 bf16 frames, fp32 maps, {pname}, forward (20 B/px)", ms, 20 * N * H * W, N, k)
go16 = torch.rand(N, C, H, W, device="cuda").to(torch.bfloat16)
for pad, pname in ((0, "zeros"), (1, "border")):
    ms = t(lambda: pw.warp2d_backward(go16, fr, g, pad, False, (True, True)), 3, 1); k = _lib.last_kernel()
    # bf16 grad_out + fp32 map + bf16 frame read, fp32 grad_grid written, grad_input written once (fp32 buffer) + rounded copy
    row(f"config 5: same, backward (both gradients), {pname} (34 B/px)", ms, 34 * N * H * W, N, k)
del fr, g, go16
# the cascade's call-site fusion: three maps, one frame, pre / post scale folded (R/main_new.py:103-107), 16x3x256x256
N, C, H, W = 16, 3, 256, 256
fr = torch.rand(N, C, H, W, device="cuda") * 2 - 1
g3 = [maps("smooth", N, H, W) + 0.002 * i for i in range(3)]
def ref_seq(sampler):
    return [sampler((fr + 1) * 127.5, g) / 127.5 - 1 for g in g3]
ms = t(lambda: pw.warp_stages(fr, g3, pre=(1.0, 127.5), post=(127.5, -1.0)), 50); k = _lib.last_kernel()
aten_ms = t(lambda: ref_seq(lambda x, g: torch.ops.aten.grid_sampler_2d(x, g, 0, 0, False)), 50)
row("config 3: three stage maps on one frame, (x+1)*127.5 .. /127.5-1 folded, forward (one launch)", ms, (3 * 4 + 3 * (8 + 12)) * N * H * W, N, k, aten_ms)
gr3 = [g.clone().requires_grad_(True) for g in g3]
gos = [torch.rand(N, C, H, W, device="cuda") for _ in range(3)]
def fb(fn):
    for g in gr3: g.grad = None
    outs = fn()
    torch.autograd.backward(list(outs), gos)
ms = t(lambda: fb(lambda: pw.warp_stages(fr, gr3, pre=(1.0, 127.5), post=(127.5, -1.0))), 30); k = _lib.last_kernel()
aten_ms = t(lambda: fb(lambda: [torch.ops.aten.grid_sampler_2d((fr + 1) * 127.5, g, 0, 0, False) / 127.5 - 1 for g in gr3]), 30)
row("config 3: the same, forward + backward through autograd (grad to the three maps)", ms, (3 * 4 + 3 * (8 + 12) + 3 * 4 + 3 * (12 + 8 + 8)) * N * H * W, N, k, aten_ms)
ours_ms = t(lambda: fb(lambda: [pw.grid_sample((fr + 1) * 127.5, g, "bilinear", "zeros", False) / 127.5 - 1 for g in gr3]), 30)
print(f"| (the same through three pw.grid_sample calls + torch elementwise ops: {ours_ms:.3f} ms) | | | | | | |")

# sensitivity of the backward to the roughness of the map (the bench map stretches by up to +-12 %)
N, C, H, W = 16, 3, 1080, 1920
fr = torch.rand(N, C, H, W, device="cuda") * 255
go = torch.rand(N, C, H, W, device="cuda")
rng = np.random.default_rng(1)
ident = synth.identity_map(4, H, W, False)
for name, gmap in (("identity map", ident),
                   ("identity + 0.03 tanh(noise on a 3x3 lattice): stretch <= 3 %", ident + synth.smooth_drift(4, H, W, rng, amp=0.03, ncell=2)),
                   ("bench map (9x9 lattice): stretch <= 12 %", synth.make_map("smooth", 4, H, W, False, seed=1))):
    g = planar(torch.from_numpy(np.ascontiguousarray(gmap, dtype=np.float32)).cuda().repeat(4, 1, 1, 1).contiguous())
    ms = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, True)), 20); k = _lib.last_kernel()
    row(f"1080p backward (both gradients), {name}", ms, 52 * N * H * W, N, k)
    ms = t(lambda: pw.warp2d_backward(go, fr, g, 0, False, (True, False)), 20)
    row(f"1080p backward, grad to frame only, {name}", ms, 32 * N * H * W, N, k)
