"""Wall-clock per case (debug): finds configurations that stall."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import pwstablenet_b200 as pw
import synth
for (N, C, H, W) in [(16, 3, 256, 256), (32, 3, 720, 1280), (16, 3, 1080, 1920), (2, 3, 64, 64), (1, 3, 1080, 1920)]:
    g = torch.from_numpy(synth.make_map("smooth", min(N, 4), H, W, False, seed=1)).cuda()
    g = g.repeat((N + 3) // 4, 1, 1, 1)[:N].permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    fr = torch.rand(N, C, H, W, device="cuda") * 255
    go = torch.rand(N, C, H, W, device="cuda")
    for mask in ((True, True), (True, False), (False, True)):
        torch.cuda.synchronize()
        ts = []
        for i in range(8):
            t0 = time.time()
            pw.warp2d_backward(go, fr, g, 0, False, mask)
            torch.cuda.synchronize()
            ts.append((time.time() - t0) * 1e3)
        print((N, C, H, W), mask, " ".join(f"{t:8.3f}" for t in ts), flush=True)
    t0 = time.time()
    for i in range(8):
        pw.warp2d_forward(fr, g, 0, False)
    torch.cuda.synchronize()
    print((N, C, H, W), "fwd x8", f"{(time.time()-t0)*1e3:8.3f} ms", flush=True)
