"""The reference's training graph (R/main_new.py:94-214, replayed by tools/harness/train_step.py) through
`pwstablenet_b200.install()` on one GPU: the grad_input of the chained warp (:197) feeds the grad_output of the stage
warps (:106,:116), the maps are `permute + affine` views, the frames are slices of a 37-channel sample.  netG's
gradients must equal those of the same graph on stock torch kernels."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.fixture(scope="module")
def setup():
    import pwstablenet_b200 as pw
    from harness import train_step as T
    from harness.netg_standin import NetGStandIn
    torch.manual_seed(123)
    # the stand-in at quarter width (the contract at the warp boundary does not depend on the width)
    net = NetGStandIn(ngf=24, head=128).cuda()
    # random-init maps are degenerate (everything samples the centre): give the output head a bias so that the
    # three stages sample different, spread-out positions
    with torch.no_grad():
        net.linear.bias.copy_(torch.tensor([1.0, 0.02, 0.01, -0.02, 1.0, -0.01]))
    batch = T.synth_batch(2, torch.device("cuda"), seed=3)
    yield pw, T, net, batch
    pw.uninstall()


def test_netg_gradients_through_install_equal_stock_torch(setup):
    pw, T, net, batch = setup
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    try:
        pw.uninstall()
        g_ref, l_ref = T.grads_of(net, batch)
        g_ref2, _ = T.grads_of(net, batch)
        noise = float((g_ref2 - g_ref).abs().max())
        pw.install()
        from pwstablenet_b200 import _lib
        l0 = _lib.launch_count()
        g_pw, l_pw = T.grads_of(net, batch)
        assert _lib.launch_count() - l0 == 20, "the warps did not go through libpwswarp"      # 11 forward + 9 backward (the gray warps feed no loss term: SURVEY 3.1)
        g_fused, l_fused = T.grads_of(net, batch, fused=True)
    finally:
        pw.uninstall()
        torch.backends.cudnn.deterministic = False
    scale = float(g_ref.abs().max())
    assert scale > 0
    tol = max(1e-4 * scale, 4 * noise)
    assert abs(l_pw - l_ref) <= 1e-5 * abs(l_ref) and abs(l_fused - l_ref) <= 1e-5 * abs(l_ref)
    assert float((g_pw - g_ref).abs().max()) <= tol, (float((g_pw - g_ref).abs().max()), scale, noise)
    assert float((g_fused - g_ref).abs().max()) <= tol, (float((g_fused - g_ref).abs().max()), scale, noise)


def test_a_training_step_runs_and_updates_the_weights(setup):
    pw, T, net, batch = setup
    pw.install()
    try:
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.5, 0.999))
        before = net.out[0].weight.detach().clone()
        loss, terms = T.train_step(net, opt, batch, fused=True)
        assert torch.isfinite(loss)
        assert not torch.equal(before, net.out[0].weight.detach())
    finally:
        pw.uninstall()
