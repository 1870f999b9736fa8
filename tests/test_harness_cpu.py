"""The config-3 / config-4 harnesses' host logic on the CPU: the netG stand-in honours the reference's contract at the
warp boundary (checked against the reference's own module where /root/reference exists), the synthetic batch has the
DataLoader's shapes, and the clip window builder produces the reference's 31-frame windows."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from harness import netg_standin as S  # noqa: E402
from harness import train_step as T  # noqa: E402


def test_standin_contract():
    net = S.NetGStandIn(ngf=8, head=16)
    x = torch.rand(2, 31, 256, 256) * 2 - 1
    maps, drifts = net(x)
    assert len(maps) == 3 and len(drifts) == 3
    for m, d in zip(maps, drifts):
        assert tuple(m.shape) == (2, 256, 256, 2) and tuple(d.shape) == (2, 256, 256, 2)
        # planar storage seen through permute(0,2,3,1): what every live call site of the reference hands over
        assert m.stride() == (2 * 256 * 256, 256, 1, 256 * 256)
        assert d.stride() == (2 * 256 * 256, 256, 1, 256 * 256)
    ev = net(x, False)
    assert tuple(ev.shape) == (2, 256, 256, 2) and ev.stride() == (2 * 256 * 256, 256, 1, 256 * 256)


def test_standin_size_matches_the_reference():
    n = sum(p.numel() for p in S.build_netg().parameters())
    assert abs(n - 48535944) / 48535944 < 1e-4          # R/lib/networks_cascading.py UnetGenerator(31, 2, 64)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout is only in the build container")
def test_reference_netg_has_the_same_contract():
    import warnings
    ref = S.load_reference_netg()
    x = torch.rand(1, 31, 256, 256) * 2 - 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        maps, drifts = ref(x)
        ev = ref(x, False)
    own = S.NetGStandIn(ngf=8, head=16)
    m2, d2 = own(x)
    assert sum(p.numel() for p in ref.parameters()) == 48535944
    for a, b in zip(list(maps) + list(drifts) + [ev], list(m2) + list(d2) + [own(x, False)]):
        assert a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype


def test_synthetic_batch_shapes():
    b = T.synth_batch(3, torch.device("cpu"), seed=1)
    assert tuple(b["images1"].shape) == (3, 37, 256, 256) and b["images1"].dtype == torch.uint8
    assert tuple(b["features1"].shape) == (3, 400, 6) and tuple(b["feature_adjacent"].shape) == (3, 6)
    st, un, fs, fu = T.pre_propossing(b["images1"], b["features1"])
    assert tuple(un.shape) == (3, 34, 256, 256) and tuple(st.shape) == (3, 3, 256, 256)
    assert tuple(fs.shape) == (3, 3, 400) and float(un.min()) >= -1 and float(un.max()) <= 1
