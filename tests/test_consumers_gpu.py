"""The loss-side consumers on the GPU, with the warp they call routed through libpwswarp.so by install()."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_crop_resize_runs_on_the_library_kernel_and_matches_slice_plus_upsample():
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import _lib, consumers as C
    from test_consumers_cpu import restated_frame_clip_batchsize
    g = torch.Generator().manual_seed(7)
    n, size = 16, 256
    seq = torch.rand((n, 3, size, size), generator=g)
    affine = torch.tensor([1.0, 0, 0, 0, 1, 0]).repeat(n, 1) + torch.randn((n, 6), generator=g) * 0.02
    affine[:, 0] = 0.85; affine[:, 4] = 0.9
    ref = restated_frame_clip_batchsize(seq, affine, size)          # R/lib/utils.py:304-336 restated on the CPU
    pw.install()
    try:
        l0 = _lib.launch_count()
        got = C.crop_resize_batch(seq.cuda(), affine.cuda())
        assert _lib.launch_count() - l0 == 1                        # one launch for the whole batch
    finally:
        pw.uninstall()
    assert float((got.cpu() - ref).abs().max()) <= 2e-6
