"""Drop-in for the reference's MaskedAdam (model/adam.py:167-221): same constructor, param-group keys
(`lr`, `betas`, `eps`, `skip_zero_grad`), lazy state (`step`, `exp_avg`, `exp_avg_sq`) and
`set_pervoxel_lr`; every update runs the sm_100a streaming kernels of csrc/adam.cu.  Dense groups of
small tensors (the 16 MLP tensors) are batched into one multi-tensor launch instead of one launch
per tensor.  Shared with the reference: the public surface (names, defaults, state keys) that its training loop and its
checkpoints address; nothing else."""
import ctypes as C

import torch

from . import adam_upd_cuda
from ._lib import check, lib, ptr, stream


class MaskedAdam(torch.optim.Optimizer):
    """What callers and checkpoints depend on is kept exactly: constructor signature and defaults, the `per_lr`
    attribute, group keys and the state keys `step`, `exp_avg`, `exp_avg_sq` (in that order).  Everything else -- which
    kernel runs for which tensor, the batching of small tensors -- is this implementation's."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        for what, value, ok in (("learning rate", lr, lr >= 0.0), ("epsilon value", eps, eps >= 0.0),
                                ("beta parameter at index 0", betas[0], 0.0 <= betas[0] < 1.0),
                                ("beta parameter at index 1", betas[1], 0.0 <= betas[1] < 1.0)):
            if not ok:
                raise ValueError(f"Invalid {what}: {value}")
        self.per_lr = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def set_pervoxel_lr(self, count):
        self.per_lr = count.float() / count.max()

    def _moments(self, param):
        """(exp_avg, exp_avg_sq, step) of `param`, created on first use, step already advanced."""
        st = self.state[param]
        if not st:
            st.update(step=0, exp_avg=torch.zeros_like(param, memory_format=torch.preserve_format),
                      exp_avg_sq=torch.zeros_like(param, memory_format=torch.preserve_format))
        st['step'] += 1
        return st['exp_avg'], st['exp_avg_sq'], st['step']

    @torch.no_grad()
    def step(self):
        small = []   # (param, grad, m, v, step, lr, beta1, beta2, eps) of dense tensors below 2^20 elements
        for group in self.param_groups:
            lr, (beta1, beta2), eps = group['lr'], group['betas'], group['eps']
            masked = group.get('skip_zero_grad', False)
            for param in group['params']:
                grad = param.grad
                if grad is None:
                    continue
                m, v, t = self._moments(param)
                if self.per_lr is not None and param.shape == self.per_lr.shape:
                    adam_upd_cuda.adam_upd_with_perlr(param, grad, m, v, self.per_lr, t, beta1, beta2, lr, eps)
                elif masked:
                    adam_upd_cuda.masked_adam_upd(param, grad, m, v, t, beta1, beta2, lr, eps)
                elif param.numel() < (1 << 20) and param.is_contiguous() and grad.is_contiguous():
                    small.append((param, grad, m, v, t, lr, beta1, beta2, eps))
                else:
                    adam_upd_cuda.adam_upd(param, grad, m, v, t, beta1, beta2, lr, eps)
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
