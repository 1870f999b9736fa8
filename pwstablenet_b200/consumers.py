"""Consumers of the warp map in the reference's training loss, restated without host round trips.

R/lib/utils.py:339-362 `loss_calulate` gathers the map at the stable feature points of every sample in a Python loop
with two `.cpu().numpy()` conversions per sample (32 device synchronisations per call, 6 calls per step): every step of
the data-parallel job stalls on them.  `map_feature_loss` does the same gather with one advanced-indexing call on the
device; values and gradients are those of the reference's loop (tests/test_consumers_cpu.py compares them against the
reference's own function where /root/reference is available, and against a restated loop elsewhere).

These stay PyTorch (they are elementwise / gather glue around the hot path, not the hot path): no kernels here.
"""
from __future__ import annotations

import torch


def map_feature_loss(grid: torch.Tensor, feature_stable: torch.Tensor, feature_unstable: torch.Tensor,
                     input_size: int, number_feature: int) -> torch.Tensor:
    """sum_i ||feature_unstable[i, 0:2, :] - grid[i, y_i, x_i, :]^T||_2^2 / number_feature / N
    with (x, y) = int((feature_stable[i, 0 or 1, :] + 1) * input_size / 2)   (R/lib/utils.py:341-347).

    grid (N, H, W, 2); feature_* (N, 3, P) as `pre_propossing` returns them (R/lib/utils.py:244-253)."""
    n = grid.size(0)
    # .int() truncates toward zero, as the reference's host-side cast does
    ys = ((feature_stable[:, 1, :] + 1) * input_size / 2).int().long()
    xs = ((feature_stable[:, 0, :] + 1) * input_size / 2).int().long()
    b = torch.arange(n, device=grid.device).unsqueeze(1).expand_as(xs)
    pos = grid[b, ys, xs, :]                                   # (N, P, 2): one gather, no synchronisation
    diff = feature_unstable[:, 0:2, :] - pos.transpose(1, 2)
    # torch.dist(a, b) ** 2 == sum of squares; keep the reference's sqrt-then-square form per sample so the gradient
    # at zero distance behaves identically
    per = torch.pow(torch.sqrt((diff * diff).sum(dim=(1, 2))), 2) / number_feature
    return per.sum() / n


def map_smoothness(grid: torch.Tensor) -> torch.Tensor:
    """(mean |d grid / dx| + mean |d grid / dy|) / 2 over neighbouring map entries (R/lib/utils.py:352-358)."""
    dx = torch.abs(grid[:, :, :-1, :] - grid[:, :, 1:, :])
    dy = torch.abs(grid[:, :-1, :, :] - grid[:, 1:, :, :])
    return (dx.mean() + dy.mean()) / 2


def block_affine_residual(drift: torch.Tensor, basis: torch.Tensor, blocks: int) -> torch.Tensor:
    """The shape loss R/lib/utils.py:405-425 (`loss_pixel1`): cut the (N, S, S, 2) drift into blocks x blocks tiles,
    least-squares fit every tile with the 4 bilinear corner basis functions (fp64, as the reference) and return the L1
    norm of the residual.  The reference builds the (blocks^2 * N) problems by concatenation in a double Python loop and
    inverts A^T A per problem; the basis is the same for every tile, so one projector P = A (A^T A)^-1 A^T serves all.

    basis: (T, 4) fp64, T = (S / blocks)^2, the corner weights of one tile (`tile_basis`)."""
    n, s, _, _ = drift.shape
    t = s // blocks
    d = drift.to(torch.float64).reshape(n, blocks, t, blocks, t, 2).permute(1, 3, 0, 2, 4, 5).reshape(blocks * blocks * n, t * t, 2)
    a = basis
    proj = a @ torch.linalg.inv(a.T @ a) @ a.T                 # (T, T)
    fit = proj.unsqueeze(0) @ d
    return (fit - d).abs().sum().to(torch.float32)


def tile_basis(t: int, device=None) -> torch.Tensor:
    """Bilinear corner weights of a t x t tile, (t*t, 4) fp64 (R/lib/utils.py:427-447 `generate_affine_matrix`)."""
    y, x = torch.meshgrid(torch.arange(t, dtype=torch.float64, device=device), torch.arange(t, dtype=torch.float64, device=device), indexing="ij")
    x2 = y2 = float(t - 1)
    q11 = (x2 - x) * (y2 - y) / (x2 * y2)
    q21 = x * (y2 - y) / (x2 * y2)
    q12 = (x2 - x) * y / (x2 * y2)
    q22 = x * y / (x2 * y2)
    return torch.stack([q11, q21, q12, q22], dim=-1).reshape(t * t, 4)


def crop_resize_batch(sequence: torch.Tensor, affine: torch.Tensor, sampler=None) -> torch.Tensor:
    """`frame_clip_batchsize` (R/lib/utils.py:304-336, the GAN branch's crop helper) as ONE warp call with no host round trip.

    The reference maps the four frame corners through every sample's affine, intersects the boxes over the batch
    (`.cpu().numpy()` synchronisations), slices that box out of every sample in a Python loop, resizes each slice to the
    frame size with `nn.Upsample(mode='bilinear')` and writes the results into a CPU tensor.  A crop followed by a
    half-pixel bilinear resize IS a bilinear sample on an axis-aligned lattice: here the box stays on the device, the
    lattice `src = (j + 0.5) * crop / size - 0.5 + origin` (clamped to the crop, as the resize clamps at its borders) is
    built with a few elementwise ops and the whole batch is one `grid_sample` call (padding 'border', align_corners=False) --
    through `torch.nn.functional.grid_sample`, i.e. this library's kernel once `install()` has run.

    sequence (N,C,S,S); affine (N,6) or (N,2,3).  Returns (N,C,S,S) on sequence's device."""
    import torch.nn.functional as F
    sampler = sampler or F.grid_sample
    n, c, s, s2 = sequence.shape
    assert s == s2, "the reference crops square frames (opt.input_size)"
    dev, dt = sequence.device, sequence.dtype
    boundary = torch.tensor([[-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]], dtype=dt, device=dev).t()      # (3,4)
    bound = torch.matmul(affine.view(-1, 2, 3).to(dt), boundary)                                           # (N,2,4)
    x_s = bound[:, 0, [0, 2]].max().clamp_min(-1.0); x_e = bound[:, 0, [1, 3]].min().clamp_max(1.0)
    y_s = bound[:, 1, [0, 1]].max().clamp_min(-1.0); y_e = bound[:, 1, [2, 3]].min().clamp_max(1.0)
    # int(...) truncates; the bounds are >= 0 here
    x0 = ((x_s + 1) * s / 2).trunc(); x1 = ((x_e + 1) * s / 2).trunc()
    y0 = ((y_s + 1) * s / 2).trunc(); y1 = ((y_e + 1) * s / 2).trunc()
    j = torch.arange(s, dtype=dt, device=dev) + 0.5
    sx = (j * ((x1 - x0) / s) - 0.5).clamp_min(0.0) + x0           # upsample_bilinear2d's source index, in frame pixels
    sy = (j * ((y1 - y0) / s) - 0.5).clamp_min(0.0) + y0
    sx = torch.minimum(sx, x1 - 1); sy = torch.minimum(sy, y1 - 1)  # the +1 neighbour stays inside the crop
    gx = (2 * sx + 1) / s - 1                                       # normalised, align_corners=False
    gy = (2 * sy + 1) / s - 1
    grid = torch.stack([gx.view(1, s).expand(s, s), gy.view(s, 1).expand(s, s)], dim=-1).unsqueeze(0).expand(n, s, s, 2)
    return sampler(sequence, grid.contiguous(), mode="bilinear", padding_mode="border", align_corners=False)
