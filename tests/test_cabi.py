"""The C-ABI library loads and exports every symbol include/fgs_b200.h declares (no compute
calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            txt = open(os.path.join(inc, fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(fgs_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    path = os.path.join(ROOT, "fgs_nerf_b200", "libfgs_b200.so")
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    lib.fgs_abi_version.restype = ctypes.c_int
    assert lib.fgs_abi_version() == 1


def test_ctypes_signatures_cover_header():
    from fgs_nerf_b200 import _lib
    import fgs_nerf_b200.engine  # noqa: F401  (registers the fused-step entry points)
    import fgs_nerf_b200.coarse  # noqa: F401  (the coarse-stage entry points)
    undeclared = [n for n in _declared() if n not in _lib._SIGS]
    assert not undeclared, f"no ctypes signature for: {undeclared}"


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "fgs_nerf_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "fgs_oracle" in txt:
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
