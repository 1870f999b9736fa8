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
