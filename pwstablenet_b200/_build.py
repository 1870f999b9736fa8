"""Builds libpwswarp.so in-tree with nvcc for sm_100a (no JIT cache: the built
library travels with the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpwswarp.so")
SOURCES = ["capi.cu", "pws_launch.cu", "warp_fwd.cu", "warp_fwd_tile.cu", "warp_fwd_tma.cu", "warp_bwd.cu", "warp_bwd_lean.cu", "warp_bwd_tma.cu", "warp_fused.cu", "warp_stages.cu"]
HEADERS = ["pws_common.cuh", "pws_tile.cuh", "pws_tma.cuh", "pws_pipe.cuh", "pws_launch.cuh", "pws_f32x2.cuh", os.path.join("..", "..", "include", "pwswarp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-cudart", "shared",
] + os.environ.get("PWS_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (need CUDA 12.9 for sm_100a)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into pwstablenet_b200/libpwswarp.so. Returns the path."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    link = [nvcc, "-shared", "-cudart", "shared", "-o", LIB, *objs,
            "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(link)
    return LIB


TORCH_EXT = os.path.join(HERE, "_pws_torch.so")
TORCH_EXT_SRC = os.path.join(CSRC, "torch_binding.cpp")


def build_torch_binding(force: bool = False) -> str:
    """Compile csrc/torch_binding.cpp (host C++ only: the compiled twin of functional.py's ctypes shim, with a C++ autograd
    node) into pwstablenet_b200/_pws_torch.so, linked against libpwswarp.so next to it.  One g++ call, in-tree."""
    lib = build_library()
    deps = [TORCH_EXT_SRC, os.path.join(HERE, "..", "include", "pwswarp.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(TORCH_EXT) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_EXT) for d in deps):
        return TORCH_EXT
    import sysconfig
    import torch
    from torch.utils import cpp_extension as X
    inc = X.include_paths("cuda") if "device_type" in X.include_paths.__code__.co_varnames else X.include_paths(True)
    inc += [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wno-attributes",
           "-DTORCH_EXTENSION_NAME=_pws_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           *["-isystem" + i for i in inc], TORCH_EXT_SRC, "-o", TORCH_EXT,
           "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python",
           "-L" + HERE, "-l:" + os.path.basename(lib), "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return TORCH_EXT


if __name__ == "__main__":
    import sys
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_binding(force="-f" in sys.argv))
