"""ctypes binding of libfgs_b200.so (the C-ABI declared in include/fgs_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing or a symbol
cannot be resolved, importing this module raises -- ops never silently route elsewhere.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfgs_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it first (python __graft_entry__.py build, or "
        f"python fgs_nerf_b200/build.py). fgs_b200 has no CPU / eager fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_f = C.c_float
_i = C.c_int
_l = C.c_int64

_SIGS = {
    "fgs_abi_version": (C.c_int, []),
    "fgs_last_error": (C.c_char_p, []),
    "fgs_launch_count": (_l, []),
    "fgs_reset_launch_count": (None, []),
    "fgs_set_device": (_i, [_i]),
    "fgs_device_sm_count": (_i, [_i, _p, _p, _p]),
    "fgs_prof_begin": (_i, [_p]),
    "fgs_prof_end": (_i, []),
    "fgs_prof_get": (_i, [_i, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    "fgs_memset_zero": (_i, [_p, _l, _p]),
    "fgs_prof_mark": (_i, [_i, _p]),
    "fgs_infer_t_minmax": (_i, [_p, _p, _p, _p, _f, _f, _i, _p, _p, _p]),
    "fgs_infer_n_samples": (_i, [_p, _p, _p, _f, _i, _p, _p]),
    "fgs_infer_ray_start_dir": (_i, [_p, _p, _p, _i, _p, _p, _p]),
    "fgs_sample_scratch_bytes": (_l, [_i]),
    "fgs_sample_pts_count": (_i, [_p, _p, _p, _p, _f, _f, _f, _i, _p, _p, _p, _p, _p, _p]),
    "fgs_sample_pts_fill": (_i, [_p, _p, _p, _p, _p, _p, _f, _i, _l, _p, _p, _p, _p, _p]),
    "fgs_sample_ndc_pts_on_rays": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "fgs_sample_bg_pts_on_rays": (_i, [_p, _p, _p, _f, _i, _i, _p, _p]),
    "fgs_maskcache_lookup": (_i, [_p, _p, _p, _p, _i, _i, _i, _l, _p, _p]),
    "fgs_raw2alpha": (_i, [_p, _f, _f, _p, _l, _p, _p, _p]),
    "fgs_raw2alpha_backward": (_i, [_p, _p, _f, _p, _l, _p, _p]),
    "fgs_alpha2weight": (_i, [_p, _p, _l, _i, _p, _p, _p, _p, _p, _p]),
    "fgs_alpha2weight_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _l, _p, _p, _p, _p]),
    "fgs_adam_upd": (_i, [_p, _p, _p, _p, _p, _l, _i, _f, _f, _f, _f, _i, _p]),
    "fgs_adam_upd_multi": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _p]),
    "fgs_adam_upd_multi_zero_grad": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _p]),
    "fgs_adam_upd_zero_grad": (_i, [_p, _p, _p, _p, _l, _i, _f, _f, _f, _f, _p]),
    "fgs_brick_list_ints": (_l, [_i, _i, _i]),
    "fgs_voxel_list_ints": (_l, [_l]),
    "fgs_voxel_list": (_i, [_p, _l, _p, _p]),
    "fgs_voxel_bits_pack": (_i, [_p, _p, _l, _p]),
    "fgs_voxel_bits_union_list": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "fgs_voxel_copy": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "fgs_brick_list": (_i, [_p, _i, _i, _i, _p, _p]),
    "fgs_brick_copy": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "fgs_adam_upd_bricks": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _p]),
    "fgs_adam_upd_brick_flags": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _p]),
    "fgs_total_variation_add_grad": (_i, [_p, _p, _p, _f, _f, _f, _i, _i, _i, _i, _i, _i, _p]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def declare(name, restype, argtypes):
    """Register a further entry point (used by the fused-step module)."""
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    _SIGS[name] = (restype, argtypes)
    return fn


def check(rc):
    if rc != 0:
        msg = lib.fgs_last_error()
        raise RuntimeError(msg.decode() if msg else f"fgs_b200 error {rc}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_cur_dev = [-1]


def stream():
    """Current torch stream handle; also keeps the library's device in step with torch's."""
    d = torch.cuda.current_device()
    if d != _cur_dev[0]:
        check(lib.fgs_set_device(d))
        _cur_dev[0] = d
    return torch.cuda.current_stream().cuda_stream


def check_input(t, name, dtype=None):
    """The reference's CHECK_INPUT (model/cuda/render_utils.cpp:46-48): CUDA + contiguous,
    raised as RuntimeError like TORCH_CHECK does.  The reference dispatches float/double but
    computes in fp32 temporaries; this implementation is fp32 only and says so."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype} (got {t.dtype}); fgs_b200 kernels are fp32-only")


def prof_begin():
    check(lib.fgs_prof_begin(stream()))


def prof_end():
    """-> list of (launch name, milliseconds), in launch order."""
    n = lib.fgs_prof_end()
    if n < 0:
        check(1)
    out = []
    name, ms = C.c_char_p(), C.c_float()
    for i in range(n):
        check(lib.fgs_prof_get(i, C.byref(name), C.byref(ms)))
        out.append((name.value.decode(), float(ms.value)))
    return out
