"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/pwswarp.h declares, and the Python shim keeps torch's argument errors.  No GPU
compute is issued here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pwswarp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pws_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pwstablenet_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert "pws_warp2d_forward" in names and "pws_warp2d_backward" in names
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert lib.pws_abi_version() == 1
    assert isinstance(_lib.launch_count(), int)


def test_abi_rejects_bad_arguments_without_touching_the_gpu():
    from pwstablenet_b200 import _lib
    lib = _lib.load()
    t = _lib.PwsTensor()
    t.data = None; t.dtype = _lib.PWS_F32; t.device = 0
    for i, (sz, st) in enumerate(zip((2, 3, 4, 4), (48, 16, 4, 1))):
        t.size[i] = sz; t.stride[i] = st
    g = _lib.PwsTensor()
    g.data = None; g.dtype = _lib.PWS_F32; g.device = 0
    for i, (sz, st) in enumerate(zip((1, 4, 4, 2), (32, 8, 2, 1))):
        g.size[i] = sz; g.stride[i] = st
    rc = lib.pws_warp2d_forward(ctypes.byref(t), ctypes.byref(g), ctypes.byref(t), 0, 0, 0, None)
    assert rc == _lib.PWS_EINVAL and b"same batch size" in lib.pws_last_error()
    rc = lib.pws_warp2d_forward(ctypes.byref(t), ctypes.byref(g), ctypes.byref(t), 1, 0, 0, None)
    assert rc == _lib.PWS_EUNSUPPORTED and b"bilinear" in lib.pws_last_error()
    rc = lib.pws_warp2d_forward(ctypes.byref(t), ctypes.byref(g), ctypes.byref(t), 0, 7, 0, None)
    assert rc == _lib.PWS_EINVAL
    rc = lib.pws_warp2d_forward(None, ctypes.byref(g), ctypes.byref(t), 0, 0, 0, None)
    assert rc == _lib.PWS_EINVAL


def test_python_shim_argument_errors_match_torch():
    import pwstablenet_b200 as pw
    f, g = torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2)
    with pytest.raises(ValueError, match="expected mode to be"):
        pw.grid_sample(f, g, mode="cubic")
    with pytest.raises(ValueError, match="expected padding_mode to be"):
        pw.grid_sample(f, g, padding_mode="wrap")
    with pytest.raises(NotImplementedError):
        pw.grid_sample(f, g, mode="bicubic", align_corners=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pw.grid_sample(f, g, align_corners=False)


def test_fused_entry_rejects_cpu_tensors_and_bad_modes():
    # the fused map-composition entry has no CPU path either, and says so before any kernel is involved
    import pwstablenet_b200 as pw
    frame = torch.zeros(1, 3, 8, 8)
    drift = torch.zeros(1, 4, 4, 2)
    with pytest.raises(RuntimeError, match="4-D CUDA tensor"):
        pw.warp_fused(frame, drift=drift, base="identity", upsample="aligned", out_size=(8, 8))
    with pytest.raises(RuntimeError, match="4-D CUDA tensor"):
        pw.warp_fused(torch.zeros(3, 8, 8), drift=drift)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pwstablenet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pwstablenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("DESIGN.md", ""), f"{fn} mentions the oracle"


def test_compiled_torch_shim_loads_and_binds_the_same_library():
    """csrc/torch_binding.cpp -> _pws_torch.so: importable without a GPU, linked against the libpwswarp.so next to it."""
    import importlib
    from pwstablenet_b200 import _build, _lib
    _build.build_torch_binding()
    m = importlib.import_module("pwstablenet_b200._pws_torch")
    assert m.abi_version() == _lib.ABI_VERSION
    for name in ("grid_sample", "warp2d_forward", "warp2d_backward"):
        assert callable(getattr(m, name))
    maps = open("/proc/self/maps").read()
    assert maps.count(os.path.join("pwstablenet_b200", "libpwswarp.so")) > 0
    assert len({l.split()[-1] for l in maps.splitlines() if l.endswith("libpwswarp.so")}) == 1   # one copy, not two


def test_aten_override_registers_for_cuda_only_and_uninstalls():
    """install(aten_override=True) replaces the CUDA kernels of the two operators; CPU tensors keep ATen's CPU kernels, the
    functional-level patch keeps refusing them, and uninstall() drops both."""
    import torch.nn.functional as F
    import pwstablenet_b200 as pw
    from pwstablenet_b200 import functional
    f, g = torch.rand(1, 1, 4, 4), torch.rand(1, 4, 4, 2) * 2 - 1
    ref = torch.grid_sampler(f, g, 0, 0, False)
    orig = F.grid_sample
    pw.install(aten_override=True)
    try:
        assert functional._aten_lib is not None and F.grid_sample is pw.grid_sample
        assert torch.equal(torch.grid_sampler(f, g, 0, 0, False), ref)       # CPU dispatch key untouched
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            F.grid_sample(f, g, align_corners=False)
        pw.install(aten_override=True)                                        # idempotent
    finally:
        pw.uninstall()
    assert functional._aten_lib is None and F.grid_sample is orig
    assert torch.equal(torch.grid_sampler(f, g, 0, 0, False), ref)
