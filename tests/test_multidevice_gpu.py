"""One process, several devices and streams ("safe under concurrent Python threads on different devices",
the reference's DataParallel model) and the backward's co-residency requirement.

* the dynamic-shared-memory attribute of the TMA kernels is per device: the first launch on a SECOND GPU of the same
  process must work (round 1 cached it per process);
* two threads driving two devices at once;
* the persistent backward waits for every CTA of its launch: it is launched cooperatively, so it must finish (and be
  right) while another stream keeps the SMs busy with a long kernel;
* more launches in flight than there are counter slots, over several streams."""
import threading

import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pw():
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import _lib
    _lib.load()
    # this module is about the persistent TMA kernels: reach them with small inputs too (pws_small_problem_elems)
    prev = _lib.small_problem_elems(0)
    yield pw
    _lib.small_problem_elems(prev)


def case(dev, n=8, h=540, w=960, seed=1):     # above the small-problem threshold: the TMA kernels take it
    g = torch.from_numpy(synth.make_map("smooth", n, h, w, False, seed=seed)).to(dev)
    g = g.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    fr = torch.from_numpy(synth.make_frames(n, 3, h, w, seed=seed + 1)).to(dev)
    go = torch.from_numpy(synth.make_gout(n, 3, h, w, seed=seed + 2)).to(dev)
    return fr, g, go


def check(pw, fr, g, go):
    from pwstablenet_b200 import _lib
    out = pw.warp2d_forward(fr, g, 0, False)
    assert _lib.last_kernel() == "fwd_tma"
    gin, gg = pw.warp2d_backward(go, fr, g, 0, False, (True, True))
    assert _lib.last_kernel() == "bwd_tma"
    with torch.cuda.device(fr.device):
        ref = torch.ops.aten.grid_sampler_2d(fr, g, 0, 0, False)
        rin, rg = torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, (True, True))
    assert torch.equal(out, ref)
    assert float((gg - rg).abs().max()) <= 1e-5 * float(rg.abs().max())
    assert float((gin - rin).abs().max()) <= 1e-4 * float(rin.abs().max())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process (gpurun --gpus 2)")
def test_second_device_in_the_same_process(pw):
    for d in range(min(torch.cuda.device_count(), 4)):
        dev = torch.device("cuda", d)
        check(pw, *case(dev, seed=10 + d))
    # and back on the first one
    check(pw, *case(torch.device("cuda", 0), seed=99))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process (gpurun --gpus 2)")
def test_two_threads_two_devices(pw):
    errors = []

    def worker(d):
        try:
            dev = torch.device("cuda", d)
            torch.cuda.set_device(dev)
            for it in range(4):
                check(pw, *case(dev, seed=20 + 7 * d + it))
            torch.cuda.synchronize(dev)
        except Exception as e:  # noqa: BLE001
            errors.append((d, repr(e)))

    ts = [threading.Thread(target=worker, args=(d,)) for d in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors


@pytest.mark.timeout(300)
def test_backward_next_to_a_long_kernel_on_another_stream(pw):
    dev = torch.device("cuda", 0)
    fr, g, go = case(dev, n=4, h=540, w=960, seed=31)
    rin, rg = torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, (True, True))
    side = torch.cuda.Stream(dev)
    a = torch.randn(8192, 8192, device=dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(30):        # ~ tens of milliseconds of SM-filling GEMMs
            a = (a @ a) * 1e-4
    for _ in range(5):
        gin, gg = pw.warp2d_backward(go, fr, g, 0, False, (True, True))
    torch.cuda.synchronize()
    assert float((gg - rg).abs().max()) <= 1e-5 * float(rg.abs().max())
    assert float((gin - rin).abs().max()) <= 1e-4 * float(rin.abs().max())


@pytest.mark.timeout(300)
def test_more_launches_in_flight_than_counter_slots(pw):
    dev = torch.device("cuda", 0)
    fr, g, go = case(dev, n=6, h=512, w=512, seed=41)
    rin, rg = torch.ops.aten.grid_sampler_2d_backward(go, fr, g, 0, 0, False, (True, True))
    ref = torch.ops.aten.grid_sampler_2d(fr, g, 0, 0, False)
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    results = []
    torch.cuda.synchronize()
    for it in range(200):          # 64 slots per ring: they are reused many times over, from four streams
        with torch.cuda.stream(streams[it % 4]):
            out = pw.warp2d_forward(fr, g, 0, False)
            gin, gg = pw.warp2d_backward(go, fr, g, 0, False, (True, True))
            if it % 25 == 0:
                results.append((out, gin, gg))
    torch.cuda.synchronize()
    for out, gin, gg in results:
        assert torch.equal(out, ref)
        assert float((gg - rg).abs().max()) <= 1e-5 * float(rg.abs().max())
        assert float((gin - rin).abs().max()) <= 1e-4 * float(rin.abs().max())


def test_stream_capture_takes_the_non_persistent_kernels(pw):
    # a captured launch would bake its counter slot into the graph: under capture the library uses the kernels that need none
    from pwstablenet_b200 import _lib
    dev = torch.device("cuda", 0)
    fr, g, go = case(dev, n=6, h=512, w=512, seed=51)
    ref = torch.ops.aten.grid_sampler_2d(fr, g, 0, 0, False)
    out = torch.empty_like(ref)
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        pw.warp2d_forward(fr, g, 0, False, out=out)       # warm-up outside capture
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            pw.warp2d_forward(fr, g, 0, False, out=out)
            kernel = _lib.last_kernel()
    assert kernel in ("fwd_lean", "fwd_direct")
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
