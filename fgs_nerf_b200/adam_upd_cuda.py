"""Drop-in for the reference's `adam_upd_cuda` extension module
(/root/reference model/cuda/adam_upd.cpp:36-86): in-place Adam updates, return None."""
import ctypes as C

import torch

from ._lib import check, check_input, lib, ptr, stream

_F32 = torch.float32


def _upd(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, mode):
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        check_input(t, n, _F32)
    if perlr is not None:
        check_input(perlr, "perlr", _F32)
    with torch.cuda.device(param.device):
        check(lib.fgs_adam_upd(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(perlr),
                               param.numel(), int(step), float(beta1), float(beta2), float(lr),
                               float(eps), mode, stream()))


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:36-46"""
    _upd(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 0)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:48-58: p, m, v untouched where grad == 0"""
    _upd(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 1)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    """adam_upd.cpp:60-77"""
    _upd(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, 2)


def adam_upd_multi(params, grads, exp_avgs, exp_avg_sqs, steps, lrs, beta1, beta2, eps):
    """New (no reference equivalent): dense Adam over up to 32 small tensors in ONE launch;
    element-wise results are bit-identical to calling adam_upd per tensor."""
    n = len(params)
    if n == 0:
        return
    for t in list(params) + list(grads) + list(exp_avgs) + list(exp_avg_sqs):
        check_input(t, "tensor", _F32)
    PP = (C.c_void_p * n)(*[p.data_ptr() for p in params])
    GG = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
    MM = (C.c_void_p * n)(*[m.data_ptr() for m in exp_avgs])
    VV = (C.c_void_p * n)(*[v.data_ptr() for v in exp_avg_sqs])
    NN = (C.c_int64 * n)(*[p.numel() for p in params])
    SS = (C.c_int * n)(*[int(s) for s in steps])
    LL = (C.c_float * n)(*[float(l) for l in lrs])
    with torch.cuda.device(params[0].device):
        check(lib.fgs_adam_upd_multi(n, PP, GG, MM, VV, NN, SS, LL, float(beta1), float(beta2),
                                     float(eps), stream()))
