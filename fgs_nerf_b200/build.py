"""In-tree build of libfgs_b200.so: plain nvcc on csrc/*.cu for sm_100a (no torch headers, so
each translation unit compiles in seconds; the .so travels to the GPU box with the snapshot)."""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libfgs_b200.so")
OBJ = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC", "-Xptxas", "-v",
    # IEEE div/sqrt, no fast-math: ids/masks must match the reference bit for bit
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(PKG, "..", "include", "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = ["nvcc", "-c", s, "-o", o] + NVCC_FLAGS
            log = open(o + ".log", "w")
            procs.append((s, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), o + ".log"))
    failed = False
    for s, p, log in procs:
        if p.wait() != 0:
            failed = True
            sys.stderr.write(open(log).read())
        elif verbose:
            sys.stdout.write(open(log).read())
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
