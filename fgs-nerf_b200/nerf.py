"""Drop-ins for the operator-level pieces of the reference's scene model (model/nerf.py) that other
reference code imports by name: `Alphas2Weights` (:1173-1189, also model/dvgo.py:390-406) and
`MaskCache` (:1192-1209).  The fused fine-stage forward / training step itself is
fgs_nerf_b200.engine.FineStepEngine."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import render_utils_cuda
from .engine import grid_sample


class Alphas2Weights(torch.autograd.Function):
    """alpha [M], ray_id [M] (sorted), N -> weights [M], alphainv_last [N]   (model/nerf.py:1173-1189)"""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = render_utils_cuda.alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = render_utils_cuda.alpha2weight_backward(
            alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
            grad_weights.contiguous(), grad_last.contiguous())
        return grad, None, None


class MaskCache(nn.Module):
    """Known-free-space test on the max-pooled SDF mask of an earlier stage (model/nerf.py:1192-1209).
    Construct from a checkpoint path like the reference, or directly from tensors."""

    def __init__(self, path=None, mask_cache_thres=None, stage=None, ks=3, sdf_mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        self.mask_cache_thres = mask_cache_thres
        if path is not None:
            st = torch.load(path)
            xyz_min = st['MaskCache_kwargs']['xyz_min']
            xyz_max = st['MaskCache_kwargs']['xyz_max']
            sdf_mask = F.max_pool3d(st['model_state_dict']['sdf_mask.grid'], kernel_size=ks, padding=ks // 2, stride=1)
        self.register_buffer('xyz_min', torch.as_tensor(xyz_min, dtype=torch.float32))
        self.register_buffer('xyz_max', torch.as_tensor(xyz_max, dtype=torch.float32))
        self.register_buffer('sdf_mask', sdf_mask.float().contiguous())

    @torch.no_grad()
    def forward(self, xyz):
        v = grid_sample(self.sdf_mask, xyz.to(self.sdf_mask.device), self.xyz_min.cpu(), self.xyz_max.cpu())[..., 0]
        return v >= self.mask_cache_thres
