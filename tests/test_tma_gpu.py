"""GPU parity of the TMA-pipelined persistent kernels (csrc/warp_fwd_tma.cu, csrc/warp_bwd_tma.cu) on the
shapes and maps that exercise their special cases: partial tiles, frames smaller than a box, the three box
shapes, border tiles (masked body on the box), tiles whose box fits no shape and NaN / inf maps (global
gather inside the persistent kernel), both map layouts, more frames than one launch of the in-kernel
zero-fill handles, repeated launches (counter slots are recycled).

Forward: bit-exact against ATen's CUDA kernel and, at small sizes, the CPU oracle.  Backward: grad_grid
bit-exact against the oracle at small sizes, <= 1e-5 relative against ATen; grad_input <= 1e-4 relative
(atomic ordering), plus the size-independent checksum sum(grad_input) == sum(grad_out * valid tap weights)."""
import numpy as np
import pytest
import torch

import synth
from oracle import oracle

pytestmark = pytest.mark.gpu
PAD = {"zeros": 0, "border": 1}


@pytest.fixture(scope="module")
def pw():
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import _lib
    _lib.load()
    # this module is about the persistent TMA kernels: reach them with small inputs too (pws_small_problem_elems)
    prev = _lib.small_problem_elems(0)
    yield pw
    _lib.small_problem_elems(prev)


def last_kernel():
    from pwstablenet_b200 import _lib
    return _lib.last_kernel()


def planar(g):
    return g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)


def make(kind, N, C, H, W, Ho, Wo, align, layout, seed=11):
    g = torch.from_numpy(synth.make_map(kind, N, Ho, Wo, align, seed=seed)).cuda()
    if layout == "planar":
        g = planar(g)
    frames = torch.from_numpy(synth.make_frames(N, C, H, W, seed=seed + 1)).cuda()
    gout = torch.from_numpy(synth.make_gout(N, C, Ho, Wo, seed=seed + 2)).cuda()
    return frames, g, gout


# (N, C, H, W, Ho, Wo): all TMA-eligible (W, Wo multiples of 4)
SHAPES = [
    (2, 3, 64, 64, 64, 64),        # frame smaller than every box
    (3, 1, 100, 128, 100, 128),    # C = 1, partial tile rows
    (2, 3, 48, 200, 40, 132),      # output size != frame size, partial tile columns
    (1, 3, 270, 480, 270, 480),    # many tiles, all three box shapes with the smooth map
    (5, 3, 16, 64, 16, 64),        # exactly one tile per frame
    (2, 3, 8, 4, 8, 4),            # smaller than one tile in both directions
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["smooth", "identity", "random", "centre", "noisy"])
@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_tma_forward_backward_parity(pw, shape, kind, pad, layout):
    N, C, H, W, Ho, Wo = shape
    for align in (False, True):
        frames, g, gout = make(kind, N, C, H, W, Ho, Wo, align, layout)
        out = pw.warp2d_forward(frames, g, PAD[pad], align)
        assert last_kernel() == "fwd_tma"
        ref = torch.ops.aten.grid_sampler_2d(frames, g, 0, PAD[pad], align)
        assert torch.equal(out, ref)
        gi, gg = pw.warp2d_backward(gout, frames, g, PAD[pad], align, (True, True))
        assert last_kernel() == "bwd_tma"
        ri, rg = torch.ops.aten.grid_sampler_2d_backward(gout, frames, g, 0, PAD[pad], align, [True, True])
        assert float((gi - ri).abs().max()) <= 1e-4 * max(float(ri.abs().max()), 1e-30)
        assert float((gg - rg).abs().max()) <= 1e-5 * max(float(rg.abs().max()), 1e-30)
        # one gradient at a time (the reference's own call sites want exactly one)
        gi1, none = pw.warp2d_backward(gout, frames, g, PAD[pad], align, (True, False))
        assert none is None and float((gi1 - ri).abs().max()) <= 1e-4 * max(float(ri.abs().max()), 1e-30)
        none, gg1 = pw.warp2d_backward(gout, frames, g, PAD[pad], align, (False, True))
        assert none is None and torch.equal(gg1, gg)


@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("align", [False, True])
def test_tma_against_oracle_small(pw, pad, align):
    N, C, H, W = 2, 3, 40, 72
    for kind in ("smooth", "noisy"):
        frames = synth.make_frames(N, C, H, W, seed=3)
        grid = synth.make_map(kind, N, H, W, align, seed=4)
        gout = synth.make_gout(N, C, H, W, seed=5)
        f, g, go = (torch.from_numpy(a).cuda() for a in (frames, grid, gout))
        out = pw.warp2d_forward(f, planar(g), PAD[pad], align)
        assert last_kernel() == "fwd_tma"
        np.testing.assert_array_equal(out.cpu().numpy(), oracle.forward(frames, grid, pad, align))
        gi, gg = pw.warp2d_backward(go, f, planar(g), PAD[pad], align, (True, True))
        assert last_kernel() == "bwd_tma"
        ogi, ogg = oracle.backward(gout, frames, grid, pad, align)
        np.testing.assert_array_equal(gg.cpu().numpy(), ogg)
        assert np.abs(gi.cpu().numpy() - ogi).max() <= 1e-4 * max(np.abs(ogi).max(), 1e-30)


def test_tma_nan_inf_and_far_coordinates(pw):
    # NaN / inf / huge coordinates send a tile down the generic path of the persistent kernel; ATen's CUDA
    # semantics (-100 guard, fmin/fmax clipping of NaN under border padding) must survive
    N, C, H, W = 1, 3, 64, 128
    for pad in (0, 1):
        frames, g, gout = make("smooth", N, C, H, W, H, W, False, "planar")
        g = g.clone()
        g[0, 3, 5, 0] = float("nan"); g[0, 20, 70, 1] = float("inf"); g[0, 40, 100, 0] = -float("inf")
        g[0, 50, 9, 1] = 3.0e38; g[0, 60, 64, 0] = -7.5
        g = planar(g)
        out = pw.warp2d_forward(frames, g, pad, False)
        assert last_kernel() == "fwd_tma"
        ref = torch.ops.aten.grid_sampler_2d(frames, g, 0, pad, False)
        assert torch.equal(out, ref)
        gi, gg = pw.warp2d_backward(gout, frames, g, pad, False, (True, True))
        ri, rg = torch.ops.aten.grid_sampler_2d_backward(gout, frames, g, 0, pad, False, [True, True])
        assert float((gi - ri).abs().max()) <= 1e-4 * float(ri.abs().max())
        finite = torch.isfinite(rg)
        assert torch.equal(torch.isfinite(gg), finite)
        assert float((gg[finite] - rg[finite]).abs().max()) <= 1e-5 * float(rg[finite].abs().max())


def test_tma_more_frames_than_one_launch_and_slot_recycling(pw):
    # 300 frames > the 256 frames one launch of the in-kernel zero-fill covers; then enough launches to wrap
    # around the 64 counter slots
    N, C, H, W = 300, 1, 16, 64
    frames, g, gout = make("smooth", N, C, H, W, H, W, False, "planar")
    gi, gg = pw.warp2d_backward(gout, frames, g, 0, False, (True, True))
    assert last_kernel() == "bwd_tma"
    ri, rg = torch.ops.aten.grid_sampler_2d_backward(gout, frames, g, 0, 0, False, [True, True])
    assert float((gi - ri).abs().max()) <= 1e-4 * float(ri.abs().max())
    assert float((gg - rg).abs().max()) <= 1e-5 * float(rg.abs().max())
    f2, g2, go2 = make("noisy", 3, 3, 48, 64, 48, 64, False, "interleaved")
    r2, _ = torch.ops.aten.grid_sampler_2d_backward(go2, f2, g2, 0, 0, False, [True, False])
    for _ in range(150):
        gi2, _ = pw.warp2d_backward(go2, f2, g2, 0, False, (True, False))
    assert float((gi2 - r2).abs().max()) <= 1e-4 * float(r2.abs().max())


def test_tma_grad_input_checksum_at_1080p(pw):
    # size-independent property at the bench size: sum(grad_input) == sum over pixels of grad_out * (sum of valid tap weights)
    N, C, H, W = 2, 3, 1080, 1920
    frames, g, gout = make("smooth", N, C, H, W, H, W, False, "planar")
    gi, gg = pw.warp2d_backward(gout, frames, g, 0, False, (True, True))
    assert last_kernel() == "bwd_tma"
    x0, y0, mask, w = pw.warp_taps(g, H, W, "zeros", False)
    valid = torch.stack([(mask >> k) & 1 for k in range(4)], dim=-1).to(torch.float64)
    wsum = (w.to(torch.float64) * valid).sum(-1)                      # (N,H,W)
    expect = float((gout.to(torch.float64) * wsum[:, None]).sum())
    got = float(gi.to(torch.float64).sum())
    assert abs(got - expect) <= 1e-6 * abs(expect)
    out = pw.warp2d_forward(frames, g, 0, False)
    assert torch.equal(out, torch.ops.aten.grid_sampler_2d(frames, g, 0, 0, False))


def test_layouts_tma_cannot_describe_fall_back(pw):
    # rows that are not 16-byte aligned (W % 4 != 0) take the non-TMA kernels; results stay exact
    frames, g, gout = make("smooth", 2, 3, 37, 53, 37, 53, False, "planar")
    out = pw.warp2d_forward(frames, g, 0, False)
    assert last_kernel() in ("fwd_lean", "fwd_direct")
    assert torch.equal(out, torch.ops.aten.grid_sampler_2d(frames, g, 0, 0, False))
    gi, gg = pw.warp2d_backward(gout, frames, g, 0, False, (True, True))
    assert last_kernel() in ("bwd_lean", "bwd_march")


@pytest.mark.parametrize("shape", SHAPES + [(1, 3, 1080, 1920, 1080, 1920)])
@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_tma_forward_channels_last_frames(pw, shape, pad):
    # the inference site (R/main_new.py:679-684,716): the frame is permute(0,3,1,2) of an HWC buffer; the
    # channels-last box path must equal ATen bit for bit and return a dense NCHW result like ATen does
    N, C, H, W, Ho, Wo = shape
    if C != 3:
        pytest.skip("channels-last only differs from NCHW for C > 1")
    for kind in ("smooth", "random", "noisy") if H < 1000 else ("smooth",):
        for align in (False, True):
            for layout in ("planar", "interleaved"):
                frames, g, _ = make(kind, N, C, H, W, Ho, Wo, align, layout)
                cl = frames.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
                assert cl.stride(1) == 1 and cl.stride(3) == C
                out = pw.warp2d_forward(cl, g, PAD[pad], align)
                assert last_kernel() == "fwd_tma_cl"
                assert out.is_contiguous()
                ref = torch.ops.aten.grid_sampler_2d(frames, g, 0, PAD[pad], align)
                assert torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_tma_forward_16bit_frames_fp32_maps(pw, dtype, pad):
    # BASELINE config 5: bf16 frames with fp32 maps.  Semantics of every 16-bit kernel of this library: taps upcast to
    # fp32, ATen's fp32 arithmetic, one rounding of the result -- so the expectation is ATen on the upcast frame, rounded
    for shape in [(2, 3, 64, 64, 64, 64), (3, 1, 100, 128, 100, 128), (2, 3, 48, 200, 40, 136), (1, 3, 270, 480, 270, 480),
                  (1, 3, 2160, 3840, 2160, 3840)]:
        N, C, H, W, Ho, Wo = shape
        for kind in ("smooth", "random", "noisy") if H < 1000 else ("smooth",):
            for align in (False, True):
                for layout in ("planar", "interleaved"):
                    frames, g, _ = make(kind, N, C, H, W, Ho, Wo, align, layout)
                    f16 = frames.to(dtype)
                    out = pw.warp2d_forward(f16, g, PAD[pad], align)
                    assert last_kernel() == "fwd_tma_16"
                    assert out.dtype == dtype
                    ref = torch.ops.aten.grid_sampler_2d(f16.float(), g, 0, PAD[pad], align).to(dtype)
                    assert torch.equal(out, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("shape", [(2, 3, 270, 480), (1, 3, 2160, 3840), (2, 1, 96, 160)])
def test_backward_16bit_frames_fp32_maps(pw, dtype, pad, shape):
    # BASELINE config 5 ("4K warp, bf16 frames with fp32 maps, zeros and border padding"), backward: torch rejects the dtype
    # mix, so the oracle is the fp32 backward on the UPCAST frame and grad_output -- grad_grid bit-exact against this
    # library's own fp32 path, <= 1e-5 against ATen; grad_input accumulated in fp32 (<= 1e-4 against ATen) and rounded to the
    # frame type once
    N, C, H, W = shape
    frames, g, gout = make("smooth", N, C, H, W, H, W, False, "planar", seed=61)
    f16, go16 = frames.to(dtype), gout.to(dtype)
    gin, gg = pw.warp2d_backward(go16, f16, g, PAD[pad], False, (True, True))
    assert last_kernel() == ("bwd_tma_16" if C == 3 else "bwd_lean_16")   # the 16-bit TMA backward has the RGB instantiations only
    assert gin.dtype == dtype and gg.dtype == torch.float32
    rin, rg = pw.warp2d_backward(go16.float(), f16.float(), g, PAD[pad], False, (True, True))
    assert torch.equal(gg, rg)
    ain, ag = torch.ops.aten.grid_sampler_2d_backward(go16.float(), f16.float(), g, 0, PAD[pad], False, (True, True))
    assert float((gg - ag).abs().max()) <= 1e-5 * float(ag.abs().max())
    acc = torch.empty(f16.shape, dtype=torch.float32, device="cuda")
    pw.warp2d_backward(go16, f16, g, PAD[pad], False, (True, False), grad_input=acc)
    assert float((acc - ain).abs().max()) <= 1e-4 * float(ain.abs().max())
    assert float((gin.float() - ain).abs().max()) <= (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -10) * float(ain.abs().max()) * 1.01
    # through autograd, as a user would call it
    fa = f16.clone().requires_grad_(True)
    ga = g.clone().requires_grad_(True)
    out = pw.grid_sample(fa, ga, "bilinear", pad, False)
    out.backward(go16)
    assert fa.grad.dtype == dtype and torch.equal(ga.grad, gg)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_tma_backward_16bit_layouts_masks_and_rough_maps(pw, dtype, pad):
    # the persistent backward with 16-bit grad_output tiles and frame boxes (bwd_tma_16): every map kind, both align modes,
    # planar and interleaved maps, partial tiles, each output mask -- against this library's fp32 path on the upcast inputs
    # (grad_grid bit-exact: same arithmetic on the same values) and ATen (grad_input <= 1e-4, accumulated in fp32)
    for shape in [(2, 3, 144, 256, 144, 256), (1, 3, 270, 480, 250, 392), (3, 3, 96, 200, 96, 200)]:   # (16-bit rows: widths in multiples of 8)
        N, C, H, W, Ho, Wo = shape
        for kind in ("smooth", "noisy", "random", "centre"):
            for align in (False, True):
                for layout in ("planar", "interleaved"):
                    frames, g, gout = make(kind, N, C, H, W, Ho, Wo, align, layout, seed=71)
                    f16, go16 = frames.to(dtype), gout.to(dtype)
                    acc = torch.empty(f16.shape, dtype=torch.float32, device="cuda")
                    _, gg = pw.warp2d_backward(go16, f16, g, PAD[pad], align, (True, True), grad_input=acc)
                    assert last_kernel() == "bwd_tma_16"
                    rin, rg = pw.warp2d_backward(go16.float(), f16.float(), g, PAD[pad], align, (True, True))
                    assert last_kernel() == "bwd_tma"
                    assert torch.equal(gg, rg)
                    ain, _ = torch.ops.aten.grid_sampler_2d_backward(go16.float(), f16.float(), g, 0, PAD[pad], align, (True, False))
                    scale = float(ain.abs().max())
                    assert float((acc - ain).abs().max()) <= 1e-4 * scale
                    acc1 = torch.full(f16.shape, 7.0, dtype=torch.float32, device="cuda")       # need not arrive zeroed
                    pw.warp2d_backward(go16, f16, g, PAD[pad], align, (True, False), grad_input=acc1)
                    assert float((acc1 - ain).abs().max()) <= 1e-4 * scale
                    _, gg1 = pw.warp2d_backward(go16, f16, g, PAD[pad], align, (False, True))
                    assert torch.equal(gg1, rg)
