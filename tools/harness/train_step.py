"""BASELINE config 3: the reference's optimisation step (R/main_new.py:81-216) replayed on synthetic data.

    netG (stand-in or, in this container, the reference's own) -> 11 warps -> losses -> backward -> Adam

The step below follows R/main_new.py line by line (cited inline); the only differences are the synthetic batch in place
of the DataLoader, `vgg16(weights=None)` in place of the downloaded weights (flagged synthetic), the
device-side loss consumers of pwstablenet_b200.consumers in place of the host-synchronising loops (values identical:
tests/test_consumers_cpu.py), and DistributedDataParallel (one process per GPU, NCCL all-reduce of netG's 194 MB of
gradients over NVLink) in place of nn.DataParallel (R/lib/networks_cascading.py:51-52).

The warp is whatever `torch.nn.functional.grid_sample` is bound to: the reference's call sites go through the module
attribute, so `pwstablenet_b200.install()` switches the whole step to the sm_100a kernels without touching this file.
`fused=True` takes the composed call sites instead (multi-map single-read + folded pre/post scale, `pw.warp_stages`).
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.nn.functional as functional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pwstablenet_b200 import consumers  # noqa: E402

PERIOD = 30            # R/lib/cfg.py:4
NUM_LAYER = 3          # opt.num_layer
INPUT_SIZE = 256       # opt.input_size
NUMBER_FEATURE = 400   # opt.number_feature
BLOCK = 16             # opt.block
LAMD = 10              # opt.lamd
SHAPELOSS_WEIGHT = 1.0


def synth_batch(n, device, seed=0):
    """What `next(iter_sample)` yields (R/main_new.py:83): two clips of 37-channel uint8 samples (31 gray + 3 RGB unstable,
    3 RGB stable), 400 feature points x (stable xyz, unstable xyz), the homography between the clips' outputs."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    mk = lambda: torch.randint(0, 256, (n, PERIOD + 1 + 3 + 3, INPUT_SIZE, INPUT_SIZE), generator=g, dtype=torch.uint8)
    feat = lambda: torch.rand((n, NUMBER_FEATURE, 6), generator=g) * 1.9 - 0.95
    eye = torch.tensor([1.0, 0, 0, 0, 1, 0])
    adj = eye.unsqueeze(0) + torch.randn((n, 6), generator=g) * 0.01
    batch = dict(images1=mk(), features1=feat(), images2=mk(), features2=feat(), feature_adjacent=adj)
    return {k: v.to(device) for k, v in batch.items()}


def pre_propossing(images, features):
    # R/lib/utils.py:244-253
    images = images.float() * (1. / 255) * 2 - 1
    images_unstable = images[:, 0:PERIOD + 1 + 3, :, :]
    images_stable = images[:, PERIOD + 1 + 3:, :, :]
    feature_stable = features[:, :, 0:3].permute(0, 2, 1)
    feature_unstable = features[:, :, 3:6].permute(0, 2, 1)
    return images_stable, images_unstable, feature_stable, feature_unstable


class PerceptualLoss(torch.nn.Module):
    """R/lib/utils.py:11-32 with random weights (the real ones need a download): the same 31 feature layers, frozen."""

    def __init__(self):
        super().__init__()
        from torchvision.models.vgg import vgg16
        net = torch.nn.Sequential(*list(vgg16(weights=None).features)[:31]).eval()
        for p in net.parameters():
            p.requires_grad = False
        self.net = net

    def forward(self, out_images, target_images):
        return functional.mse_loss(self.net(out_images), self.net(target_images))


def forward_losses(netG, batch, vgg=None, fused=False, timers=None):
    """R/main_new.py:94-212.  Returns (loss_g, dict of the terms)."""
    import warnings
    warnings.filterwarnings("ignore", message="Default grid_sample and affine_grid behavior has changed")
    image_stable1, image_unstable1, feature_stable1, feature_unstable1 = pre_propossing(batch["images1"], batch["features1"].float())
    image_stable2, image_unstable2, feature_stable2, feature_unstable2 = pre_propossing(batch["images2"], batch["features2"].float())
    period = PERIOD
    n = image_unstable1.size(0)

    def tic():
        if timers is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(); return e
        return None

    def toc(e0, key):
        if timers is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); timers.setdefault(key, []).append((e0, e1))

    def warp_clip(image_unstable, grids):
        # :103-110 -- three cascade maps applied to the SAME frame, then the gray centre frame with the last map
        rgb = image_unstable[:, period + 1:period + 1 + 3, :, :]
        gray = image_unstable[:, period // 2: period // 2 + 1:, :]
        e0 = tic()
        if fused:
            import pwstablenet_b200 as pw
            fakes = list(pw.warp_stages(rgb, grids, pre=(1.0, 127.5), post=(127.5, -1.0)))
            fake_gray = pw.warp_stages(gray, [grids[NUM_LAYER - 1]], pre=(1.0, 127.5), post=(127.5, -1.0))[0]
        else:
            fakes = []
            for nl in range(NUM_LAYER):
                fake_temp = functional.grid_sample((rgb + 1) * 127.5, grids[nl])
                fakes.append(fake_temp / 127.5 - 1)
            fake_gray = functional.grid_sample((gray + 1) * 127.5, grids[NUM_LAYER - 1])
            fake_gray = fake_gray / 127.5 - 1
        toc(e0, "warp_fwd")
        return fakes, fake_gray

    e0 = tic()
    grid1, affine1 = netG(image_unstable1[:, 0:period + 1, :, :])          # :101
    toc(e0, "netg_fwd")
    fake1, fake1_gray = warp_clip(image_unstable1, grid1)
    e0 = tic()
    grid2, affine2 = netG(image_unstable2[:, 0:period + 1, :, :])          # :112
    toc(e0, "netg_fwd")
    fake2, fake2_gray = warp_clip(image_unstable2, grid2)

    basis = consumers.tile_basis(INPUT_SIZE // BLOCK, device=image_stable1.device)
    loss_mse = loss_feature = loss_delta = loss_vgg = loss_g2 = 0
    loss_pixel = 0
    feature_adjacent = batch["feature_adjacent"].float().view(-1, 2, 3)
    for nl in range(NUM_LAYER):                                           # :184-203
        for grid, fs, fu, fake, real in ((grid1[nl], feature_stable1, feature_unstable1, fake1[nl], image_stable1),
                                         (grid2[nl], feature_stable2, feature_unstable2, fake2[nl], image_stable2)):
            loss_feature = loss_feature + consumers.map_feature_loss(grid, fs, fu, INPUT_SIZE, NUMBER_FEATURE)
            loss_mse = loss_mse + torch.mean(torch.abs(real[:, 0:3, :, :] - fake))
            loss_delta = loss_delta + consumers.map_smoothness(grid)
            if vgg is not None:
                loss_vgg = loss_vgg + vgg(fake, real[:, 0:3, :, :])
        grid = functional.affine_grid(feature_adjacent, fake1[nl].size())  # :194-195
        e0 = tic()
        output2_to_output1 = functional.grid_sample(fake2[nl], grid)        # :197  (grad flows to the input)
        toc(e0, "warp_fwd")
        loss_g2 = loss_g2 + torch.mean(torch.abs(output2_to_output1 - fake1[nl]))
        # :202-203 -- overwritten every stage in the reference: only the last stage's term survives
        loss_pixel = (consumers.block_affine_residual(affine1[nl], basis, BLOCK) * SHAPELOSS_WEIGHT
                      + consumers.block_affine_residual(affine2[nl], basis, BLOCK) * SHAPELOSS_WEIGHT)
    loss_g1 = loss_feature + (loss_vgg + loss_mse) + loss_pixel              # :206
    loss_g = loss_g1 + loss_g2 * LAMD                                        # :212
    terms = dict(feature=loss_feature, mse=loss_mse, delta=loss_delta, vgg=loss_vgg, g2=loss_g2, pixel=loss_pixel)
    return loss_g, terms


def train_step(netG, optimizer, batch, vgg=None, fused=False, timers=None):
    optimizer.zero_grad(set_to_none=True)
    loss_g, terms = forward_losses(netG, batch, vgg, fused, timers)
    loss_g.backward()                                                      # :214
    optimizer.step()                                                       # :216
    return loss_g.detach(), terms


def grads_of(netG, batch, vgg=None, fused=False):
    """netG gradients of one step's loss (no optimiser step), as a flat fp32 vector + the loss value."""
    for p in netG.parameters():
        p.grad = None
    loss_g, _ = forward_losses(netG, batch, vgg, fused)
    loss_g.backward()
    return torch.cat([p.grad.reshape(-1) for p in netG.parameters()]), float(loss_g.detach())


def standalone_warps(batch_n, device, reps=5):
    """The 11 warps of one step (6 RGB + 2 gray with grad -> map, 3 chained with grad -> frame) forward + backward on
    synthetic tensors of the step's shapes, timed alone: the warp's share of the step."""
    g = torch.Generator(device="cpu").manual_seed(1)
    rgb = (torch.rand((batch_n, 3, INPUT_SIZE, INPUT_SIZE), generator=g) * 255).to(device)
    gray = (torch.rand((batch_n, 1, INPUT_SIZE, INPUT_SIZE), generator=g) * 255).to(device)
    theta = (torch.tensor([[1.0, 0, 0], [0, 1, 0]]) + torch.randn((batch_n, 2, 3), generator=g) * 0.01).to(device)
    planar = (torch.randn((batch_n, 2, INPUT_SIZE, INPUT_SIZE), generator=g) * 0.02).to(device)
    base = functional.affine_grid(theta, (batch_n, 3, INPUT_SIZE, INPUT_SIZE), align_corners=False)
    times = []
    for r in range(reps + 1):
        maps = [(planar.permute(0, 2, 3, 1) + base).requires_grad_(True) for _ in range(3)]
        chain_in = [rgb.clone().requires_grad_(True) for _ in range(3)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = []
        for clip in range(2):
            for nl in range(3):
                outs.append(functional.grid_sample(rgb, maps[nl], align_corners=False))
            outs.append(functional.grid_sample(gray, maps[2], align_corners=False))
        for nl in range(3):
            outs.append(functional.grid_sample(chain_in[nl], base, align_corners=False))
        torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1))
    return sum(times) / len(times)


def run(args):
    """bench.py --config train: one JSON line with step time, warp share, all-reduce time and gradient parity."""
    import json
    import numpy as np
    import torch.distributed as dist
    import pwstablenet_b200 as pw
    from harness.netg_standin import build_netg

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    global_batch = 16
    per_gpu = max(1, global_batch // world)
    torch.manual_seed(123)                                           # opt.seed
    netG = build_netg().to(dev)
    vgg = PerceptualLoss().to(dev)
    n_params = sum(p.numel() for p in netG.parameters())
    batch = synth_batch(per_gpu, dev, seed=10 + rank)

    # ---- gradient parity: our warp vs stock torch on the same weights and batch (deterministic convolutions)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    g_ref, l_ref = grads_of(netG, batch, vgg)
    g_ref2, _ = grads_of(netG, batch, vgg)                           # noise floor of the rest of the graph
    pw.install()
    g_pw, l_pw = grads_of(netG, batch, vgg)
    g_fused, l_fused = grads_of(netG, batch, vgg, fused=True)
    scale = float(g_ref.abs().max())
    parity = {"max_abs_grad": scale,
              "torch_vs_torch": float((g_ref2 - g_ref).abs().max()) / scale,
              "ours_vs_torch": float((g_pw - g_ref).abs().max()) / scale,
              "fused_vs_torch": float((g_fused - g_ref).abs().max()) / scale,
              "loss_torch": l_ref, "loss_ours": l_pw, "loss_fused": l_fused}
    torch.backends.cudnn.deterministic = False
    torch.backends.cudnn.benchmark = True                            # R/lib/cfg.py:57

    def timed_steps(model, fused, steps, warmup):
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))   # R/main_new.py:63
        for _ in range(warmup):
            train_step(model, opt, batch, vgg, fused)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            train_step(model, opt, batch, vgg, fused)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    model = netG
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(netG, device_ids=[local], gradient_as_bucket_view=True)
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    ms_ours = timed_steps(model, False, steps, warmup)
    ms_fused = timed_steps(model, True, steps, warmup)
    warps_ours = standalone_warps(per_gpu, dev)
    pw.uninstall()
    ms_torch = timed_steps(model, False, steps, warmup)
    warps_torch = standalone_warps(per_gpu, dev)

    # ---- the collective, alone: all-reduce of a flat fp32 buffer of netG's size
    allreduce_ms = None
    if world > 1:
        flat = torch.zeros(n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_ms = float(t.item())

    if rank == 0:
        line = {"config": "train", "metric": "netG training steps/s (batch 16, 256x256, 11 warps per step)",
                "value": 1e3 / ms_ours, "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_ours, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
                "global_batch": global_batch, "per_gpu_batch": per_gpu, "netg": "stand-in (tools/harness/netg_standin.py)",
                "netg_params": n_params, "allreduce_bytes": n_params * 4, "allreduce_ms_alone": allreduce_ms,
                "ms_per_step_fused_call_sites": ms_fused, "ms_per_step_stock_torch_warp": ms_torch,
                "warps_alone_ms": {"ours": warps_ours, "stock_torch": warps_torch,
                                   "what": "the step's 11 warps forward + backward at the step's shapes, timed alone"},
                "warp_share_of_step": warps_ours / ms_ours,
                "grad_parity_rel_to_max": parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
