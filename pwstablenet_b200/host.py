"""Host-buffer front end of the warp: forward + backward on batches that live in (pinned)
host memory, with the copies hidden behind the kernels.

The reference moves every frame across PCIe (R/main_new.py:85-92 training batches,
:649-684 inference frames up, :717 warped frame down).  Once the kernels run near the HBM
roofline the step time of a host-resident batch is the PCIe time, so the copies are
pipelined: the batch is cut into chunks of a few frames; chunk k+1 is uploaded and chunk
k-1 downloaded while chunk k is warped, on three CUDA streams with two device staging
slots (PCIe is full duplex: uploads and downloads overlap each other too).

This is plumbing around the C ABI (pws_warp2d_forward / pws_warp2d_backward); it adds no
arithmetic.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .functional import _PADDING, warp2d_backward, warp2d_forward


class HostWarpPipeline:
    """Warp fwd (+ bwd) of host batches of shape (N,C,H,W) with maps (N,2,Ho,Wo) planar-stored,
    exactly the storage netG returns (the kernels see them as (N,Ho,Wo,2) views)."""

    def __init__(self, chunk: int, channels: int, in_size: Tuple[int, int], out_size: Optional[Tuple[int, int]] = None,
                 device=None, backward: bool = True, padding_mode: str = "zeros", align_corners: bool = False,
                 dtype: torch.dtype = torch.float32):
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.chunk, self.C, self.backward = int(chunk), channels, backward
        self.pad, self.align = _PADDING[padding_mode], bool(align_corners)
        H, W = in_size
        Ho, Wo = out_size or in_size
        mk = lambda *shape: torch.empty(shape, dtype=dtype, device=self.dev)
        self.slots = []
        for _ in range(2):
            s = {"frames": mk(chunk, channels, H, W), "map": mk(chunk, 2, Ho, Wo), "out": mk(chunk, channels, Ho, Wo)}
            if backward:
                s.update(gout=mk(chunk, channels, Ho, Wo), gin=mk(chunk, channels, H, W), ggrid=mk(chunk, 2, Ho, Wo))
            s["in_ready"], s["done"], s["drained"] = (torch.cuda.Event() for _ in range(3))
            self.slots.append(s)
        self.s_up, self.s_run, self.s_down = (torch.cuda.Stream(self.dev) for _ in range(3))

    def run(self, frames: torch.Tensor, maps: torch.Tensor, out: torch.Tensor, grad_out: Optional[torch.Tensor] = None,
            grad_frames: Optional[torch.Tensor] = None, grad_maps: Optional[torch.Tensor] = None) -> None:
        """All arguments are HOST tensors (pinned for real overlap).  maps / grad_maps: (N,2,Ho,Wo).
        Returns when the last download has COMPLETED: the host buffers may be read right away."""
        n = frames.size(0)
        if self.backward and (grad_out is None or grad_frames is None or grad_maps is None):
            raise ValueError("backward pipeline needs grad_out, grad_frames and grad_maps host buffers")
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_up, self.s_run, self.s_down):
            st.wait_stream(cur)
        k = 0
        for b in range(0, n, self.chunk):
            e = min(n, b + self.chunk)
            m = e - b
            s = self.slots[k & 1]
            with torch.cuda.stream(self.s_up):
                if k >= 2:
                    self.s_up.wait_event(s["done"])       # slot's inputs were consumed by chunk k-2
                s["frames"][:m].copy_(frames[b:e], non_blocking=True)
                s["map"][:m].copy_(maps[b:e], non_blocking=True)
                if self.backward:
                    s["gout"][:m].copy_(grad_out[b:e], non_blocking=True)
                s["in_ready"].record(self.s_up)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(s["in_ready"])
                if k >= 2:
                    self.s_run.wait_event(s["drained"])   # slot's outputs of chunk k-2 have left
                grid = s["map"][:m].permute(0, 2, 3, 1)   # planar storage, (N,Ho,Wo,2) view
                warp2d_forward(s["frames"][:m], grid, self.pad, self.align, out=s["out"][:m])
                if self.backward:
                    warp2d_backward(s["gout"][:m], s["frames"][:m], grid, self.pad, self.align, (True, True),
                                    grad_input=s["gin"][:m], grad_grid=s["ggrid"][:m].permute(0, 2, 3, 1))
                s["done"].record(self.s_run)
            with torch.cuda.stream(self.s_down):
                self.s_down.wait_event(s["done"])
                out[b:e].copy_(s["out"][:m], non_blocking=True)
                if self.backward:
                    grad_frames[b:e].copy_(s["gin"][:m], non_blocking=True)
                    grad_maps[b:e].copy_(s["ggrid"][:m], non_blocking=True)
                s["drained"].record(self.s_down)
            k += 1
        for st in (self.s_up, self.s_run, self.s_down):
            cur.wait_stream(st)
        # the D2H copies into the caller's buffers are asynchronous: wait for the last one before handing them back
        self.s_down.synchronize()


class HostInferencePipeline:
    """The inference site (R/main_new.py:679-684,697-721) on host-resident clips: uint8 HWC frames up, the 256 x 256
    map lattice netG emits (planar (N,2,h,w) fp32) up, ONE fused kernel per chunk (map upsample + sample + truncation to
    uint8, `warp_fused`), uint8 HWC frames down.  12.4 MB cross the bus per 1080p frame instead of the 132.7 MB of the
    reference's float pipeline.  Copies and kernels overlap on three streams, two device slots."""

    def __init__(self, chunk: int, frame_size: Tuple[int, int], lattice_size: Tuple[int, int] = (256, 256), device=None,
                 upsample: str = "aligned", padding_mode: str = "zeros", align_corners: bool = False):
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.chunk, self.upsample, self.padding_mode, self.align = int(chunk), upsample, padding_mode, bool(align_corners)
        H, W = frame_size
        h, w = lattice_size
        self.size = (H, W)
        self.slots = []
        for _ in range(2):
            s = {"frames": torch.empty((chunk, H, W, 3), dtype=torch.uint8, device=self.dev),
                 "lattice": torch.empty((chunk, 2, h, w), dtype=torch.float32, device=self.dev),
                 "out": None}
            s["in_ready"], s["done"], s["drained"] = (torch.cuda.Event() for _ in range(3))
            self.slots.append(s)
        self.s_up, self.s_run, self.s_down = (torch.cuda.Stream(self.dev) for _ in range(3))

    def run(self, frames: torch.Tensor, lattices: torch.Tensor, out: torch.Tensor) -> None:
        """frames, out: (N,H,W,3) uint8 host tensors (cv2 layout); lattices: (N,2,h,w) fp32 host tensor (netG's stage-3
        map).  Returns when `out` is complete."""
        from .compose import warp_fused
        n = frames.size(0)
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_up, self.s_run, self.s_down):
            st.wait_stream(cur)
        k = 0
        for b in range(0, n, self.chunk):
            e = min(n, b + self.chunk)
            m = e - b
            s = self.slots[k & 1]
            with torch.cuda.stream(self.s_up):
                if k >= 2:
                    self.s_up.wait_event(s["done"])
                s["frames"][:m].copy_(frames[b:e], non_blocking=True)
                s["lattice"][:m].copy_(lattices[b:e], non_blocking=True)
                s["in_ready"].record(self.s_up)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(s["in_ready"])
                if k >= 2:
                    self.s_run.wait_event(s["drained"])
                res = warp_fused(s["frames"][:m].permute(0, 3, 1, 2), drift=s["lattice"][:m].permute(0, 2, 3, 1),
                                 upsample=self.upsample, out_size=self.size, padding_mode=self.padding_mode,
                                 align_corners=self.align, out_dtype=torch.uint8, out_channels_last=True)
                s["out"] = res.permute(0, 2, 3, 1)    # (m,H,W,3) dense; kept alive until the download has run
                s["done"].record(self.s_run)
            with torch.cuda.stream(self.s_down):
                self.s_down.wait_event(s["done"])
                out[b:e].copy_(s["out"], non_blocking=True)
                s["drained"].record(self.s_down)
            k += 1
        for st in (self.s_up, self.s_run, self.s_down):
            cur.wait_stream(st)
        self.s_down.synchronize()


def warp_host(frames: torch.Tensor, maps: torch.Tensor, grad_out: Optional[torch.Tensor] = None, chunk: int = 2,
              padding_mode: str = "zeros", align_corners: bool = False, device=None):
    """Convenience wrapper: allocates pinned result buffers, runs the pipeline once and returns
    (out, grad_frames, grad_maps) as host tensors (the gradients only when grad_out is given)."""
    n, c, h, w = frames.shape
    ho, wo = maps.shape[2], maps.shape[3]
    pipe = HostWarpPipeline(chunk, c, (h, w), (ho, wo), device=device, backward=grad_out is not None,
                            padding_mode=padding_mode, align_corners=align_corners, dtype=frames.dtype)
    out = torch.empty((n, c, ho, wo), dtype=frames.dtype).pin_memory()
    gf = torch.empty_like(frames).pin_memory() if grad_out is not None else None
    gm = torch.empty_like(maps).pin_memory() if grad_out is not None else None
    pipe.run(frames, maps, out, grad_out, gf, gm)
    return out, gf, gm
