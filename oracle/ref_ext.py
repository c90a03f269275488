"""TEST INFRASTRUCTURE ONLY -- loads the reference's own CUDA extensions from oracle/_ref/*.so
(built by oracle/build_ref.py from the sources under /root/reference; see that file).  Usable
only where a GPU is present; used by the `-m gpu` parity tests and tests/golden/make_*.py as the
checker, never by the product path."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cache = {}


def available(name="render_utils_cuda"):
    return os.path.exists(os.path.join(REF_DIR, name + ".so"))


def load(name):
    """name in {render_utils_cuda, adam_upd_cuda, total_variation_cuda} -> pybind module."""
    if name not in _cache:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        path = os.path.join(REF_DIR, name + ".so")
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]
