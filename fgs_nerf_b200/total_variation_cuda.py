"""Drop-in for the reference's `total_variation_cuda` extension module
(/root/reference model/cuda/total_variation.cpp:16-32): adds the Huber-TV gradient of a 5-D grid
into `grad` in place, returns None."""
import torch

from ._lib import check, lib, ptr, stream

_F32 = torch.float32


def _layout(t, name):
    """0 = contiguous NCDHW (what the reference requires), 1 = channels_last_3d storage of the
    same logical [1,C,X,Y,Z] tensor (this package's DenseGrid keeps multi-channel grids that way
    so a voxel's channels are one 128-bit-loadable run)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != _F32:
        raise RuntimeError(f"{name} must be torch.float32; fgs_b200 kernels are fp32-only")
    if t.dim() != 5 or t.size(0) != 1:
        raise RuntimeError(f"{name} must have shape [1,C,X,Y,Z]")
    if t.is_contiguous():
        return 0
    if t.is_contiguous(memory_format=torch.channels_last_3d):
        return 1
    raise RuntimeError(f"{name} must be contiguous")


def _run(param, grad, mask, wx, wy, wz, dense_mode):
    lay = _layout(param, "param")
    if _layout(grad, "grad") != lay or (mask is not None and _layout(mask, "mask") != lay):
        raise RuntimeError("param, grad and mask must share one memory layout")
    if grad.shape != param.shape or (mask is not None and mask.shape != param.shape):
        raise RuntimeError("param, grad and mask must have the same shape")
    _, C, X, Y, Z = param.shape
    with torch.cuda.device(param.device):
        check(lib.fgs_total_variation_add_grad(ptr(param), ptr(grad), ptr(mask), float(wx), float(wy),
                                               float(wz), C, X, Y, Z, int(bool(dense_mode)), lay,
                                               stream()))


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    """total_variation.cpp:16-21 (kernel :13-35 incl. its wz-for-i-axis quirk)"""
    _run(param, grad, None, wx, wy, wz, dense_mode)


def total_variation_add_grad_new(param, grad, mask, wx, wy, wz, dense_mode):
    """total_variation.cpp:23-27 (kernel :38-66): pairwise float mask"""
    _run(param, grad, mask, wx, wy, wz, dense_mode)
