"""Tiny driver for ncu captures of the fused inference call (uint8 HWC in/out, lattice upsampled in the kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
import pwstablenet_b200 as pw
import synth
N, H, W = 16, 1080, 1920
hwc = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")
ident = F.affine_grid(torch.tensor([[[1.0, 0, 0], [0, 1.0, 0]]], device="cuda"), (1, 3, 256, 256), align_corners=False)
drift = (torch.from_numpy(synth.make_map("smooth", 4, 256, 256, False, seed=1)).cuda() - ident).repeat(4, 1, 1, 1).permute(0, 3, 1, 2).contiguous()
theta = torch.tensor([[[1.0, 0.002, 0.0], [-0.002, 1.0, 0.0]]], device="cuda").repeat(N, 1, 1)
for _ in range(2):
    out = pw.warp_fused(hwc.permute(0, 3, 1, 2), drift=drift.permute(0, 2, 3, 1), base="affine", theta=theta,
                        upsample="aligned", out_size=(H, W), out_dtype=torch.uint8, out_channels_last=True)
torch.cuda.synchronize()
print("done", int(out.sum()))
