"""The compiled PyTorch shim (csrc/torch_binding.cpp -> _pws_torch.so) against the ctypes shim of functional.py: both call
the same libpwswarp.so entry points, so forward and grad_grid must be bit-identical and grad_input (atomic order) within
the gradient tolerance; masks, layouts and errors must match."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pw():
    import pwstablenet_b200 as m
    from pwstablenet_b200 import functional
    if not functional.torch_ext():
        pytest.fail("pwstablenet_b200/_pws_torch.so is not built (python -m pwstablenet_b200._build)")
    return m


def _inputs(N, C, H, W, kind="smooth", planar=True, dtype=torch.float32):
    f = torch.from_numpy(synth.make_frames(N, C, H, W, seed=5)).cuda().to(dtype)
    g = torch.from_numpy(synth.make_map(kind, N, H, W, False, seed=6)).cuda()
    if planar:
        g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    go = torch.from_numpy(synth.make_gout(N, C, H, W, seed=7)).cuda().to(dtype)
    return f, g, go


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (2, 1, 250, 388), (4, 3, 540, 960)])
@pytest.mark.parametrize("padding_mode", ["zeros", "border"])
@pytest.mark.parametrize("align", [False, True])
def test_ext_matches_ctypes_shim(pw, shape, padding_mode, align):
    from pwstablenet_b200 import functional as F
    f, g, go = _inputs(*shape)
    pad = F._PADDING[padding_mode]
    fa, ga = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
    out = pw.grid_sample(fa, ga, padding_mode=padding_mode, align_corners=align)     # compiled shim
    assert type(out.grad_fn).__name__ != "_Warp2dBackward"
    out.backward(go)
    fb, gb = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
    ref = F._Warp2d.apply(fb, gb, pad, align)                                         # ctypes shim
    ref.backward(go)
    assert torch.equal(out, ref)
    assert torch.equal(ga.grad, gb.grad)
    assert ga.grad.stride() == gb.grad.stride() == g.stride()        # grad_grid keeps the planar layout
    scale = fb.grad.abs().max().item()
    assert (fa.grad - fb.grad).abs().max().item() <= 1e-4 * scale     # tolerance: atomic order only
    assert fa.grad.is_contiguous()


def test_ext_output_masks(pw):
    f, g, go = _inputs(2, 3, 64, 96)
    fa = f.clone().requires_grad_(True)
    pw.grid_sample(fa, g, align_corners=False).backward(go)
    assert fa.grad is not None
    ga = g.clone().requires_grad_(True)
    l0 = pw._lib.launch_count()
    pw.grid_sample(f, ga, align_corners=False).backward(go)
    assert ga.grad is not None and pw._lib.launch_count() - l0 == 2   # forward + one backward kernel (no memset, no grad_input)
    with torch.no_grad():
        out = pw.grid_sample(f, g, align_corners=False)
    assert out.grad_fn is None and not out.requires_grad


def test_ext_half_frames(pw):
    f, g, go = _inputs(1, 3, 128, 160, dtype=torch.bfloat16)
    fa, ga = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
    pw.grid_sample(fa, ga, align_corners=False).backward(go)
    assert fa.grad.dtype == torch.bfloat16 and ga.grad.dtype == torch.float32
    gin, gg = pw.warp2d_backward(go, f, g, 0, False, (True, True))
    assert torch.equal(ga.grad, gg)
    assert (fa.grad.float() - gin.float()).abs().max().item() <= 2e-2 * gin.float().abs().max().item()


def test_ext_errors(pw):
    ext = pw.functional.torch_ext()
    f, g, _ = _inputs(1, 3, 32, 32)
    with pytest.raises(NotImplementedError):
        ext.warp2d_forward(f, g, 2, False)            # reflection padding: outside this library's scope
    with pytest.raises(RuntimeError):
        ext.warp2d_forward(f, g[..., :1], 0, False)   # grid.size(-1) != 2


def test_aten_override_routes_the_operator_itself(pw):
    """install(aten_override=True): aten::grid_sampler_2d{,_backward} on CUDA are the C-ABI entry points, reached here through
    torch.grid_sampler (no Python patch involved); ATen's autograd formula drives the backward.  uninstall() restores ATen."""
    f, g, go = _inputs(2, 3, 96, 128)
    fa, ga = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
    ref = pw.grid_sample(fa, ga, padding_mode="border", align_corners=False)
    ref.backward(go)
    pw.install(aten_override=True)
    try:
        fb, gb = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
        l0 = pw._lib.launch_count()
        out = torch.grid_sampler(fb, gb, 0, 1, False)
        assert pw._lib.launch_count() > l0, "the operator did not reach libpwswarp.so"
        assert torch.equal(out, ref)
        l1 = pw._lib.launch_count()
        out.backward(go)
        assert pw._lib.launch_count() > l1
        assert torch.equal(gb.grad, ga.grad)
        assert (fb.grad - fa.grad).abs().max().item() <= 1e-4 * fa.grad.abs().max().item()
        gc = g.clone().requires_grad_(True)
        torch.grid_sampler(f, gc, 0, 0, False).backward(go)          # output_mask (False, True)
        assert gc.grad is not None
        with pytest.raises(NotImplementedError):
            torch.grid_sampler(f, g, 1, 0, False)                    # nearest: outside the library's scope, no fallback
    finally:
        pw.uninstall()
    l2 = pw._lib.launch_count()
    out2 = torch.grid_sampler(f, g, 0, 1, False)                     # ATen's own kernel again
    assert pw._lib.launch_count() == l2
    assert torch.allclose(out2, ref.detach(), atol=1e-3, rtol=1e-5)
