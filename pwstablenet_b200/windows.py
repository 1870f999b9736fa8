"""The sliding 31-frame gray window netG sees at inference, built on the device in batches.

R/main_new.py:612-673 keeps a Python list of 31 full-resolution BGR frames; per output frame it converts ONE new frame
to gray, resizes it to 256 x 256 (cv2 INTER_AREA), uploads it and shifts a (1,31,256,256) tensor by one channel with
torch.cat -- strictly sequential, batch 1, one H2D copy and three tiny kernels per frame.  Here the 256 x 256 gray
version of every frame of a shard (its halo included) is computed once on the device, and the windows of B consecutive
output frames are ONE gather (`unfold`-style advanced indexing) producing the (B,31,256,256) batch netG wants, with the
reference's edge replication (the first frame 16 times at the start, the last frame repeated at the end).

cv2's fixed-point BGR2GRAY and its INTER_AREA resampler are host code outside the hot path; the device version uses the
same luma weights in fp32 and an exact area average, which differ from cv2's by rounding only (the harness is
self-consistent: the 1-GPU and the N-GPU run see the same windows).  No kernels of this library are involved.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F

PERIOD = 30  # R/lib/cfg.py:4


def gray_small(frames_hwc_u8: torch.Tensor, size: Tuple[int, int] = (256, 256), chunk: int = 32) -> torch.Tensor:
    """(T,H,W,3) uint8 RGB frames -> (T,h,w) float32 in [-1,1]: luma, exact area average, `/255*2-1` (R/main_new.py:644-650)."""
    out = []
    w = torch.tensor([0.299, 0.587, 0.114], device=frames_hwc_u8.device)
    for b in range(0, frames_hwc_u8.size(0), chunk):
        f = frames_hwc_u8[b:b + chunk].float()
        g = (f * w).sum(-1, keepdim=False).unsqueeze(1)            # (t,1,H,W)
        g = _area_resize(g, size)
        out.append(g.squeeze(1) / 255 * 2 - 1)
    return torch.cat(out, 0)


def _area_resize(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """Exact area-weighted average for arbitrary (also fractional) ratios: separable, as two small matrix products."""
    H, W = x.shape[-2:]
    wy = _area_matrix(H, size[0], x.device)
    wx = _area_matrix(W, size[1], x.device)
    return torch.einsum("oh,nchw,pw->ncop", wy, x, wx)


def _area_matrix(src: int, dst: int, device) -> torch.Tensor:
    """(dst, src) matrix of the overlaps of destination cell [i*s, (i+1)*s) with the source pixels, normalised by s."""
    s = src / dst
    i = torch.arange(dst, dtype=torch.float64, device=device).unsqueeze(1)
    j = torch.arange(src, dtype=torch.float64, device=device).unsqueeze(0)
    lo = torch.maximum(i * s, j)
    hi = torch.minimum((i + 1) * s, j + 1)
    return ((hi - lo).clamp_min(0) / s).to(torch.float32)


def window_batch(gray: torch.Tensor, first_frame: int, begin: int, end: int, num_frames: int, period: int = PERIOD) -> torch.Tensor:
    """Windows of output frames [begin, end) of a clip of `num_frames` frames.

    gray: (t,h,w) gray frames of the clip range starting at clip index `first_frame` (a shard passes its halo range).
    Returns (end-begin, period+1, h, w); window k of frame i is clip frame clamp(i - period/2 + k, 0, num_frames-1)."""
    half = period // 2
    i = torch.arange(begin, end, device=gray.device).unsqueeze(1)
    k = torch.arange(period + 1, device=gray.device).unsqueeze(0)
    idx = (i - half + k).clamp_(0, num_frames - 1) - first_frame
    if int(idx.min()) < 0 or int(idx.max()) >= gray.size(0):
        raise IndexError("window_batch: the gray range does not cover the halo of the requested frames")
    return gray[idx]
