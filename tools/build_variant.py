"""Development: build libpwswarp variants with extra -D flags for A/B runs on the GPU box.
usage: python tools/build_variant.py NAME file.cu[,file2.cu] -DFLAG=1 ...   -> pwstablenet_b200/var/libpwswarp_NAME.so
Select at run time with PWS_LIB_PATH."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwstablenet_b200 import _build as B

def main():
    name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
    B.build_library()
    out = os.path.join(B.HERE, "build", "var")
    libdir = os.path.join(B.HERE, "var")
    os.makedirs(libdir, exist_ok=True)
    os.makedirs(out, exist_ok=True)
    objs = []
    procs = []
    for s in B.SOURCES:
        obj = os.path.join(B.HERE, "build", s.replace(".cu", ".o"))
        if s in files:
            obj = os.path.join(out, name + "_" + s.replace(".cu", ".o"))
            procs.append(subprocess.Popen([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, s), "-o", obj]))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    lib = os.path.join(out, "libpwswarp_%s.so" % name)
    subprocess.check_call([B._nvcc(), "-shared", "-cudart", "shared", "-o", lib, *objs, "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    xz = os.path.join(libdir, os.path.basename(lib) + ".xz")
    subprocess.check_call("xz -T4 -3 -c %s > %s" % (lib, xz), shell=True)   # the push to the GPU box is slow: ship compressed
    print(xz)

main()
