"""GPU parity of the multi-map single-read op (csrc/warp_stages.cu, pw.warp_stages): K cascade maps applied to the SAME
frame in one launch with the reference's pre / post scale folded in (R/main_new.py:103-110), forward and backward.

The contract is bit-identity with the sequence of torch calls it replaces:
    fake[k] = F.grid_sample((frame + 1) * 127.5, grid[k]) / 127.5 - 1
both for the values and, through autograd, for every map's gradient (grad_input: <= 1e-4, atomic order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pw():
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import _lib
    _lib.load()
    return pw


def cascade_maps(n, h, w, align, k=3, seed=3, planar=True):
    """k maps that refine each other, stored planar like netG's (R/lib/networks_cascading.py:235)."""
    base = synth.make_map("smooth", n, h, w, align, seed=seed)
    rng = np.random.default_rng(seed + 100)
    maps = []
    for i in range(k):
        m = torch.from_numpy((base + rng.standard_normal(base.shape).astype(np.float32) * (0.004 * i)).astype(np.float32)).cuda()
        if planar:
            m = m.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
        maps.append(m)
    return maps


def unfused(frame, grids, pad, align, pre, post, sampler):
    return [sampler((frame + pre[0]) * pre[1], g, mode="bilinear", padding_mode=pad, align_corners=align) / post[0] + post[1] for g in grids]


@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("shape", [(16, 3, 256, 256), (2, 1, 256, 256), (3, 3, 100, 180)])
def test_stages_forward_is_bit_identical_to_the_reference_sequence(pw, pad, align, shape):
    from pwstablenet_b200 import _lib
    n, c, h, w = shape
    frame = torch.from_numpy(synth.make_frames(n, c, h, w, seed=5)).cuda() / 127.5 - 1          # the reference's [-1,1] frames
    grids = cascade_maps(n, h, w, align, planar=(pad == "zeros"))
    got = pw.warp_stages(frame, grids, pad, align, pre=(1.0, 127.5), post=(127.5, -1.0))
    assert _lib.last_kernel() == "stages_fwd"
    for a, b in zip(got, unfused(frame, grids, pad, align, (1.0, 127.5), (127.5, -1.0), pw.grid_sample)):
        assert torch.equal(a, b)
    if not (pad == "zeros" and align):   # (that combination detours to cuDNN in torch: 1 ulp apart, test_warp_gpu.py)
        for a, b in zip(got, unfused(frame, grids, pad, align, (1.0, 127.5), (127.5, -1.0), F.grid_sample)):
            assert torch.equal(a, b)
    # no scales: the plain sampler
    plain = pw.warp_stages(frame, grids[:2], pad, align)
    for a, g in zip(plain, grids):
        assert torch.equal(a, pw.grid_sample(frame, g, "bilinear", pad, align))


@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("align", [False, True])
def test_stages_backward_matches_autograd_of_the_reference_sequence(pw, pad, align):
    n, c, h, w = 4, 3, 128, 192
    frame = (torch.from_numpy(synth.make_frames(n, c, h, w, seed=6)).cuda() / 127.5 - 1)
    grids = cascade_maps(n, h, w, align, seed=8)
    gouts = [torch.from_numpy(synth.make_gout(n, c, h, w, seed=20 + k)).cuda() for k in range(3)]

    def grads(fn, want_frame):
        f = frame.clone().requires_grad_(want_frame)
        gs = [g.clone().requires_grad_(True) for g in grids]
        outs = fn(f, gs)
        torch.autograd.backward(list(outs), gouts)
        return [g.grad for g in gs], f.grad

    fused = lambda f, gs: pw.warp_stages(f, gs, pad, align, pre=(1.0, 127.5), post=(127.5, -1.0))
    ours = lambda f, gs: unfused(f, gs, pad, align, (1.0, 127.5), (127.5, -1.0), pw.grid_sample)
    aten = lambda f, gs: unfused(f, gs, pad, align, (1.0, 127.5), (127.5, -1.0),
                                 lambda x, g, mode, padding_mode, align_corners: torch.ops.aten.grid_sampler_2d(x, g, 0, {"zeros": 0, "border": 1}[padding_mode], align_corners))
    gg_f, _ = grads(fused, False)
    gg_o, _ = grads(ours, False)
    gg_a, gin_a = grads(aten, True)
    for a, b, r in zip(gg_f, gg_o, gg_a):
        assert a.stride() == b.stride()
        assert torch.equal(a, b)                                                    # the unfused path of this library
        assert float((a - r).abs().max()) <= 1e-5 * float(r.abs().max())            # ATen's CUDA kernel
    # gradient to the frame as well (the training sites never ask for it; the chain rule through (x + 1) * 127.5 is there)
    gg_f2, gin_f = grads(fused, True)
    for a, b in zip(gg_f2, gg_f):
        assert torch.equal(a, b)
    assert float((gin_f - gin_a).abs().max()) <= 1e-4 * float(gin_a.abs().max())


def test_stages_errors(pw):
    f = torch.zeros(1, 3, 8, 8, device="cuda")
    g = torch.zeros(1, 8, 8, 2, device="cuda")
    with pytest.raises(ValueError):
        pw.warp_stages(f, [])
    with pytest.raises(ValueError):
        pw.warp_stages(f, [g] * 5)
    with pytest.raises(NotImplementedError):
        pw.warp_stages(f, [g], padding_mode="reflection")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pw.warp_stages(f.cpu(), [g])
    with pytest.raises(RuntimeError):
        pw.warp_stages(torch.zeros(1, 2, 8, 8, device="cuda"), [g])     # C must be 1 or 3
