"""Drop-in for the reference's MaskedAdam (model/adam.py:167-221): same constructor, param-group keys
(`lr`, `betas`, `eps`, `skip_zero_grad`), lazy state (`step`, `exp_avg`, `exp_avg_sq`) and
`set_pervoxel_lr`; every update runs the sm_100a streaming kernels of csrc/adam.cu.  Dense groups of
small tensors (the 16 MLP tensors) are batched into one multi-tensor launch instead of one launch
per tensor."""
import ctypes as C

import torch

from . import adam_upd_cuda
from ._lib import check, lib, ptr, stream


class MaskedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps)
        self.per_lr = None
        super().__init__(params, defaults)

    def set_pervoxel_lr(self, count):
        self.per_lr = count.float() / count.max()

    @torch.no_grad()
    def step(self):
        small = []   # (param, grad, m, v, step, lr, betas, eps) of dense tensors below 2^20 elements
        for group in self.param_groups:
            lr = group['lr']
            beta1, beta2 = group['betas']
            eps = group['eps']
            skip_zero_grad = group.get('skip_zero_grad', False)
            for param in group['params']:
                if param.grad is None:
                    continue
                state = self.state[param]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state['step'] += 1
                if self.per_lr is not None and param.shape == self.per_lr.shape:
                    adam_upd_cuda.adam_upd_with_perlr(param, param.grad, state['exp_avg'], state['exp_avg_sq'],
                                                      self.per_lr, state['step'], beta1, beta2, lr, eps)
                elif skip_zero_grad:
                    adam_upd_cuda.masked_adam_upd(param, param.grad, state['exp_avg'], state['exp_avg_sq'],
                                                  state['step'], beta1, beta2, lr, eps)
                elif param.numel() < (1 << 20) and param.is_contiguous() and param.grad.is_contiguous():
                    small.append((param, param.grad, state['exp_avg'], state['exp_avg_sq'], state['step'], lr,
                                  beta1, beta2, eps))
                else:
                    adam_upd_cuda.adam_upd(param, param.grad, state['exp_avg'], state['exp_avg_sq'],
                                           state['step'], beta1, beta2, lr, eps)
        # one launch per (betas, eps) family, up to 32 tensors each
        while small:
            b1, b2, eps = small[0][6:9]
            take = [i for i, s in enumerate(small) if s[6:9] == (b1, b2, eps)][:32]
            batch = [small[i] for i in take]
            small = [s for i, s in enumerate(small) if i not in take]
            n = len(batch)
            vp = C.c_void_p
            with torch.cuda.device(batch[0][0].device):
                check(lib.fgs_adam_upd_multi(
                    n, (vp * n)(*[ptr(s[0]) for s in batch]), (vp * n)(*[ptr(s[1]) for s in batch]),
                    (vp * n)(*[ptr(s[2]) for s in batch]), (vp * n)(*[ptr(s[3]) for s in batch]),
                    (C.c_int64 * n)(*[s[0].numel() for s in batch]), (C.c_int * n)(*[s[4] for s in batch]),
                    (C.c_float * n)(*[s[5] for s in batch]), b1, b2, eps, stream()))
