"""Development check: TMA forward/backward vs ATen CUDA on a few shapes/maps (bit-exact forward)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pwstablenet_b200 as pw
import synth

cases = [(2, 3, 64, 64), (2, 3, 256, 256), (3, 1, 100, 128), (2, 3, 720, 1280), (1, 3, 1080, 1920)]
kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["smooth", "identity", "random", "centre", "noisy"]
bad = 0
for (N, C, H, W) in cases:
    for kind in kinds:
        for align in (False, True):
            for pad in (0, 1):
                for layout in ("planar", "inter"):
                    g = torch.from_numpy(synth.make_map(kind, N, H, W, align, seed=3)).cuda()
                    if layout == "planar":
                        g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
                    fr = torch.rand(N, C, H, W, device="cuda") * 255
                    out = pw.warp2d_forward(fr, g, pad, align)
                    ref = torch.ops.aten.grid_sampler_2d(fr, g, 0, pad, align)
                    torch.cuda.synchronize()
                    ok = torch.equal(out, ref)
                    if not ok:
                        bad += 1
                        d = (out - ref).abs()
                        print("MISMATCH", (N, C, H, W), kind, align, pad, layout, "max", float(d.max()), "count", int((d > 0).sum()))
print("forward check done, mismatches:", bad)

# ---- backward: grad_grid bit-exact vs ATen is not guaranteed (ATen uses its own order), compare with tolerance;
# grad_input within 1e-4 relative of ATen
bad = 0
for (N, C, H, W) in cases:
    for kind in kinds:
        for align in (False, True):
            for pad in (0, 1):
                for layout in ("planar", "inter"):
                    g = torch.from_numpy(synth.make_map(kind, N, H, W, align, seed=3)).cuda()
                    if layout == "planar":
                        g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
                    fr = torch.rand(N, C, H, W, device="cuda") * 255
                    go = torch.rand(N, C, H, W, device="cuda")
                    for mask in ((True, True), (True, False), (False, True)):
                        gi, gg = pw.warp2d_backward(go, fr, g, pad, align, mask)
                        ri, rg = torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, pad, align, list(mask))
                        torch.cuda.synchronize()
                        msgs = []
                        if mask[0]:
                            err = float((gi - ri).abs().max() / ri.abs().max().clamp_min(1e-30))
                            if not err < 1e-4: msgs.append(f"gin rel {err:.3e}")
                        if mask[1]:
                            err = float((gg - rg).abs().max() / rg.abs().max().clamp_min(1e-30))
                            if not err < 1e-5: msgs.append(f"ggrid rel {err:.3e}")
                        if msgs:
                            bad += 1
                            print("MISMATCH", (N, C, H, W), kind, align, pad, layout, mask, *msgs)
print("backward check done, mismatches:", bad)
