"""pwstablenet_b200.consumers against the reference's own loss helpers (R/lib/utils.py:339-362,
405-447), imported from /root/reference when it is there (this container) and against a restated loop otherwise."""
import os
import sys

import numpy as np
import pytest
import torch

from pwstablenet_b200 import consumers as C

REF = "/root/reference"


def restated_feature_loss(grid, fs, fu, input_size, number_feature):
    # R/lib/utils.py:341-347, statement for statement (host-side indices, one sample at a time)
    n = grid.size(0)
    loss = 0
    for i in range(n):
        ys = ((fs[i, 1, :] + 1) * input_size / 2).int().cpu().numpy()
        xs = ((fs[i, 0, :] + 1) * input_size / 2).int().cpu().numpy()
        pos = grid[i, ys, xs, :]
        loss = loss + torch.pow(torch.dist(fu[i, 0:2, :], torch.t(pos)), 2) / number_feature
    return loss / n


def inputs(n=4, size=64, p=50, seed=0):
    g = torch.Generator().manual_seed(seed)
    grid = (torch.rand((n, size, size, 2), generator=g) * 2 - 1).requires_grad_(True)
    fs = torch.rand((n, 3, p), generator=g) * 1.9 - 0.95
    fu = torch.rand((n, 3, p), generator=g) * 2 - 1
    return grid, fs, fu


def test_feature_loss_equals_the_per_sample_loop():
    grid, fs, fu = inputs()
    ref = restated_feature_loss(grid, fs, fu, 64, 50)
    ref.backward()
    g_ref = grid.grad.clone(); grid.grad = None
    got = C.map_feature_loss(grid, fs, fu, 64, 50)
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref))
    assert float((grid.grad - g_ref).abs().max()) <= 1e-7


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only in the build container")
def test_against_the_reference_functions():
    argv, sys.argv = sys.argv, ["x"]
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from lib import utils as U
        from lib.cfg import opt
    finally:
        sys.argv = argv
        sys.path.remove(REF)
    n = opt.batchSize
    g = torch.Generator().manual_seed(1)
    grid = (torch.rand((n, 256, 256, 2), generator=g) * 2 - 1).requires_grad_(True)
    fs = torch.rand((n, 3, 400), generator=g) * 1.9 - 0.95
    fu = torch.rand((n, 3, 400), generator=g) * 2 - 1
    fake = torch.rand((n, 3, 256, 256), generator=g)
    real = torch.rand((n, 6, 256, 256), generator=g)
    mse, delta, feat = U.loss_calulate(grid, fs, fu, fake, real, n)          # R/lib/utils.py:339
    feat.backward()
    g_ref = grid.grad.clone(); grid.grad = None
    mine = C.map_feature_loss(grid, fs, fu, opt.input_size, opt.number_feature)
    mine.backward()
    assert abs(float(mine) - float(feat)) <= 1e-6 * abs(float(feat))
    assert float((grid.grad - g_ref).abs().max()) <= 1e-9
    assert abs(float(C.map_smoothness(grid)) - float(delta)) <= 1e-7
    # shape loss: the reference's function calls .cuda(); its basis generator is host-side
    basis_ref = U.generate_affine_matrix(opt.block, opt.block)[:16, :16, 0, :].reshape(256, 4)   # one 16 x 16 tile
    np.testing.assert_allclose(C.tile_basis(16).numpy(), basis_ref, rtol=0, atol=1e-15)


def test_block_affine_residual_equals_the_concatenated_least_squares():
    # R/lib/utils.py:405-425 restated on the CPU (the reference hard-codes .cuda())
    n, size, blocks = 2, 64, 4
    t = size // blocks
    g = torch.Generator().manual_seed(2)
    drift = torch.rand((n, size, size, 2), generator=g) * 0.1
    basis = C.tile_basis(t)
    a_whole = basis.reshape(t, t, 1, 4).repeat(blocks, blocks, 1, 1).unsqueeze(0).repeat(n, 1, 1, 1, 1)
    bl, al = [], []
    d64 = drift.to(torch.float64)
    for i in range(blocks):
        for j in range(blocks):
            bl.append(d64[:, t * i:t * (i + 1), t * j:t * (j + 1), :].reshape(n, t * t, 2))
            al.append(a_whole[:, t * i:t * (i + 1), t * j:t * (j + 1), :, :].reshape(n, t * t, -1))
    bt, at = torch.cat(bl, 0), torch.cat(al, 0)
    ab = torch.bmm(at, torch.bmm(torch.bmm(torch.inverse(torch.bmm(at.permute(0, 2, 1), at)), at.permute(0, 2, 1)), bt))
    ref = torch.dist(ab, bt, 1).to(torch.float32)
    got = C.block_affine_residual(drift, basis, blocks)
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))


def restated_frame_clip_batchsize(sequence, affine, size):
    # R/lib/utils.py:304-336 on the CPU (the reference's own function hard-codes .cuda() and opt.batchSize)
    n = sequence.size(0)
    boundary = torch.tensor([[-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]], dtype=torch.float32).expand(n, 4, 3).permute(0, 2, 1)
    bound = torch.matmul(affine.view(-1, 2, 3), boundary)
    x_start = bound[:, 0, [0, 2]].numpy(); x_end = bound[:, 0, [1, 3]].numpy()
    y_start = bound[:, 1, [0, 1]].numpy(); y_end = bound[:, 1, [2, 3]].numpy()
    xs, xe, ys, ye = -1, 1, -1, 1
    for i in range(n):
        xs = max(xs, max(max(x_start[i]), -1)); xe = min(xe, min(min(x_end[i]), 1))
        ys = max(ys, max(max(y_start[i]), -1)); ye = min(ye, min(min(y_end[i]), 1))
    out = torch.empty_like(sequence)
    m = torch.nn.Upsample(size=size, mode="bilinear")
    for i in range(n):
        out[i] = m(sequence[i:i + 1, :, int((ys + 1) * size / 2):int((ye + 1) * size / 2), int((xs + 1) * size / 2):int((xe + 1) * size / 2)])[0]
    return out


def test_crop_resize_equals_slice_plus_upsample():
    g = torch.Generator().manual_seed(5)
    n, size = 4, 64
    seq = torch.rand((n, 3, size, size), generator=g)
    affine = torch.tensor([1.0, 0, 0, 0, 1, 0]).repeat(n, 1) + torch.randn((n, 6), generator=g) * 0.05
    affine[:, 0] = 0.8 + 0.05 * torch.rand(n, generator=g)      # shrink: the warped frame leaves a border to crop away
    affine[:, 4] = 0.85 + 0.05 * torch.rand(n, generator=g)
    ref = restated_frame_clip_batchsize(seq, affine, size)
    got = C.crop_resize_batch(seq, affine)
    assert float((got - ref).abs().max()) <= 2e-6
