"""Development: a few backward launches at the bench size (for traces printed by experimental builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pwstablenet_b200 as pw
import synth
N, C, H, W = 16, 3, 1080, 1920
g = torch.from_numpy(synth.make_map("smooth", 4, H, W, False, seed=1)).cuda().repeat(4, 1, 1, 1)
g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
fr = torch.rand(N, C, H, W, device="cuda") * 255
go = torch.rand(N, C, H, W, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    pw.warp2d_forward(fr, g, 0, False)
    pw.warp2d_backward(go, fr, g, 0, False, (True, True))
    torch.cuda.synchronize()
    print("---- launch", i, flush=True)
