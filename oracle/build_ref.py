#!/usr/bin/env python
"""Build the reference's own CUDA extensions into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The three extensions the reference loads on its hot path
(/root/reference/model/cuda/{render_utils,adam_upd,total_variation}{.cpp,_kernel.cu};
JIT-loaded at model/nerf.py:14-19, model/grid.py:12-24, model/adam.py:10-13) are compiled
from the sources WHERE THEY LIE under /root/reference.  Nothing is copied into the repo: the
sources are staged in a throw-away temp dir only because they need one mechanical fix to
compile against torch 2.11 -- `X.type()` -> `X.scalar_type()` inside the AT_DISPATCH macros
(torch >= 2.x no longer converts DeprecatedTypeProperties to ScalarType).  No kernel body,
launch shape or arithmetic is touched.  Only the resulting .so files land in oracle/_ref/
(git-ignored; they travel to the GPU box with the gpurun snapshot like our own .so).

The .so files are CUDA-only pybind modules: they are *executed* only on the GPU box, by the
`-m gpu` parity tests (as the checker) -- never by the product path.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CUDA = "/root/reference/model/cuda"
EXTS = {
    "render_utils_cuda": ["render_utils.cpp", "render_utils_kernel.cu"],
    "adam_upd_cuda": ["adam_upd.cpp", "adam_upd_kernel.cu"],
    "total_variation_cuda": ["total_variation.cpp", "total_variation_kernel.cu"],
}


def _patch(text: str) -> str:
    # AT_DISPATCH_FLOATING_TYPES(x.type(), ...) -> x.scalar_type()
    return re.sub(r"AT_DISPATCH_FLOATING_TYPES\(\s*([A-Za-z_0-9]+)\.type\(\)",
                  r"AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()", text)


def build_one(name, srcs, force=False):
    so = os.path.join(OUT, name + ".so")
    if os.path.exists(so) and not force:
        print(f"[build_ref] {so} exists")
        return so
    import torch
    from torch.utils import cpp_extension as ce
    tmp = tempfile.mkdtemp(prefix="fgs_ref_")
    try:
        staged = []
        for s in srcs:
            with open(os.path.join(REF_CUDA, s)) as f:
                txt = _patch(f.read())
            p = os.path.join(tmp, s)
            with open(p, "w") as f:
                f.write(txt)
            staged.append(p)
        inc = []
        for d in ce.include_paths(device_type="cuda"):
            inc += ["-isystem", d]
        inc += ["-isystem", sysconfig.get_paths()["include"]]
        common = ["-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H",
                  "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
        objs = []
        procs = []
        for p in staged:
            o = p + ".o"
            objs.append(o)
            if p.endswith(".cu"):
                # default nvcc flags of torch's JIT build (no fast-math, fmad on), sm_100a
                cmd = ["nvcc", "-c", p, "-o", o, "-std=c++17", "-O3",
                       "-gencode", "arch=compute_100a,code=sm_100a",
                       "--compiler-options", "-fPIC",
                       "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                       "-D__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr"] + common + inc
            else:
                cmd = ["g++", "-c", p, "-o", o, "-std=c++17", "-O3", "-fPIC"] + common + inc
            procs.append(subprocess.Popen(cmd))
        for pr in procs:
            if pr.wait() != 0:
                raise RuntimeError("compile failed for " + name)
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        os.makedirs(OUT, exist_ok=True)
        cmd = ["g++", "-shared", "-o", so] + objs + [
            "-L" + libdir, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart",
            "-Wl,-rpath," + libdir]
        subprocess.check_call(cmd)
        print(f"[build_ref] built {so}")
        return so
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    force = "--force" in sys.argv
    if not os.path.isdir(REF_CUDA):
        print("[build_ref] /root/reference not present (GPU box): using prebuilt oracle/_ref")
        return 0
    for name, srcs in EXTS.items():
        build_one(name, srcs, force)
    return 0


if __name__ == "__main__":
    sys.exit(main())
