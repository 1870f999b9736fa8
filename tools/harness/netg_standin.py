"""A stand-in for the reference's cascading encoder-decoder netG (R/lib/networks_cascading.py:108-237).

The conv stack is out of scope (it stays PyTorch) and the reference source cannot travel to the GPU
box, so the harnesses need something with the SAME CONTRACT at the warp's boundary:

  * input  (N, 31, 256, 256) gray stack in [-1, 1];
  * train mode returns ([map_1, map_2, map_3], [drift_1, drift_2, drift_3]) with
        drift_k = tanh(tanh(conv(...)))            stored planar (N,2,256,256), handed over as .permute(0,2,3,1)
        map_k   = drift_k.permute(0,2,3,1) + F.affine_grid(theta_k, (N,3,256,256))
    i.e. maps whose memory is PLANAR (strides (2HW, W, 1, HW)) exactly like the reference's (R/...:174,235);
  * eval mode returns map_3 only (R/...:236-237);
  * N(0, 0.02) weights, zero biases (R/...:25-46), so random-init maps are the same degenerate near-zero maps;
  * 48.5 M parameters (the reference has 48 535 944): the DDP gradient all-reduce moves the same 194 MB.

It is NOT the reference architecture (a plain three-stage U-Net cascade written for this harness; stages 2 and 3 share
weights and the theta / output heads are shared by all stages, as in the reference).  When /root/reference is present
(this container) `load_reference_netg()` returns the real module instead, for contract checks."""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

SIZE = 256


def _down(cin, cout, k=3, s=2, p=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p), nn.LeakyReLU(0.2, True))


def _up(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1), nn.ReLU(True))


class _Stage(nn.Module):
    """One encoder-decoder pass: 256 -> 2 -> 256 with skips; `extra` channels of the previous stage's features are
    concatenated to the input of every encoder level (the cascade)."""

    def __init__(self, ngf, extra):
        super().__init__()
        w = [ngf, ngf, ngf * 2, ngf * 4, ngf * 4, ngf * 4, ngf * 4, ngf * 4]       # 256,128,64,32,16,8,4,2
        self.extra = extra
        self.enc = nn.ModuleList([_down(w[i] * (2 if extra else 1), w[i + 1]) for i in range(7)])
        self.dec = nn.ModuleList([_up(w[7], w[6])] + [_up(w[7 - i] * 2, w[6 - i]) for i in range(1, 7)])
        self.fuse = nn.Sequential(nn.Conv2d(w[0] * 2, w[0], 3, 1, 1), nn.ReLU(True))

    def forward(self, x0, prev):
        feats = [x0]
        x = x0
        for i, e in enumerate(self.enc):
            inp = torch.cat([x, prev[i]], 1) if self.extra else x
            x = e(inp)
            feats.append(x)
        bottom = x                                   # (N, 4ngf, 2, 2)
        y = self.dec[0](bottom)
        for i in range(1, 7):
            y = self.dec[i](torch.cat([y, feats[7 - i]], 1))
        y = self.fuse(torch.cat([y, feats[0]], 1))   # (N, ngf, 256, 256)
        return y, bottom, feats[:7]


class NetGStandIn(nn.Module):
    def __init__(self, input_nc=31, output_nc=2, ngf=86, head=1087):
        super().__init__()
        self.transfer = _down(input_nc, ngf, 5, 1, 2)
        self.stage1 = _Stage(ngf, extra=False)
        self.stage23 = _Stage(ngf, extra=True)       # shared by stages 2 and 3 (as the reference's *_bottom modules)
        self.flatten = _down(ngf * 4, head, 2, 1, 0)
        self.linear = nn.Conv2d(head, 6, 1)
        self.out = nn.Sequential(nn.Conv2d(ngf, output_nc, 3, 1, 1), nn.Tanh())
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, 0.0, 0.02)
                nn.init.zeros_(m.bias)

    def _heads(self, y, bottom):
        theta = self.linear(self.flatten(bottom)).view(-1, 2, 3)
        affine = F.affine_grid(theta, torch.Size((theta.size(0), 3, SIZE, SIZE)), align_corners=False)
        drift = torch.tanh(self.out(y))              # tanh twice, as the reference does
        return drift.permute(0, 2, 3, 1), affine

    def forward(self, x, is_training=True):
        x0 = self.transfer(x)
        y1, b1, f1 = self.stage1(x0, None)
        y2, b2, f2 = self.stage23(x0, f1)
        y3, b3, _ = self.stage23(x0, f2)
        d3, a3 = self._heads(y3, b3)
        if not is_training:
            return d3 + a3
        d1, a1 = self._heads(y1, b1)
        d2, a2 = self._heads(y2, b2)
        return [d1 + a1, d2 + a2, d3 + a3], [d1, d2, d3]


def build_netg():
    """48 536 385 parameters with the default widths (the reference: 48 535 944)."""
    return NetGStandIn()


def load_reference_netg():
    """The reference's own netG (CPU, this container only); None when /root/reference is absent."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return None
    argv, sys.argv = sys.argv, [sys.argv[0]]
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref)
    try:
        from lib.networks_cascading import define_G
        net = define_G(31, 2, 64, "normal", 0.02)
    finally:
        sys.argv = argv
        sys.path.remove(ref)
    return net
