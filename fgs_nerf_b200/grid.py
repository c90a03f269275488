"""Drop-in for the reference's grid operator layer (model/grid.py): same classes, attributes and
method signatures, backed by the sm_100a kernels through the C-ABI.

  create_grid('DenseGrid', channels=, world_size=, xyz_min=, xyz_max=)      model/grid.py:27-33
  DenseGrid.grid  nn.Parameter [1,C,X,Y,Z]   (state_dict key '<name>.grid', as the reference)
  DenseGrid.forward(xyz[...,3]) -> [...,C] (squeezed when C == 1)           model/grid.py:49-68
  DenseGrid.scale_volume_grid / total_variation_add_grad / get_dense_grid   model/grid.py:101-122
  MaskGrid(path|mask, xyz_min, xyz_max).forward(xyz) -> bool                 model/grid.py:253-287

DenseGrid.forward is differentiable w.r.t. the grid (what the reference gets from F.grid_sample);
both directions run fgs_grid_sample_fwd / fgs_grid_sample_bwd -- ATen's grid_sampler_3d arithmetic,
bit-exact forward (tests/test_fine_gpu.py::test_grid_sample_matches_aten).  TensoRFGrid is not
provided: no config of the reference instantiates it.

What these classes share with the reference is its public surface only -- class / method / keyword names, the
parameter `grid`, the buffers `mask`, `xyz2ijk_scale`, `xyz2ijk_shift` -- because the reference's callers, configs and
checkpoints address them by those names; the bodies are written against this library's ops.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import render_utils_cuda, total_variation_cuda
from .engine import grid_resize, grid_sample, grid_sample_backward


def create_grid(type, **kwargs):
    if type == 'DenseGrid':
        return DenseGrid(**kwargs)
    raise NotImplementedError(f"grid type {type!r}: only DenseGrid is on the reference's hot path")


class _GridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, xyz, xyz_min, xyz_max):
        ctx.save_for_backward(xyz)
        ctx.meta = (tuple(grid.shape), grid.is_contiguous(memory_format=torch.channels_last_3d) and not grid.is_contiguous(),
                    xyz_min, xyz_max)
        return grid_sample(grid, xyz, xyz_min, xyz_max)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (xyz,) = ctx.saved_tensors
        shape, cl, xyz_min, xyz_max = ctx.meta
        g = grid_sample_backward(shape, cl, xyz, xyz_min, xyz_max, grad_out.contiguous())
        return g, None, None, None


class DenseGrid(nn.Module):
    def __init__(self, channels, world_size, xyz_min, xyz_max, **kwargs):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32).to('cuda')
        self.xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32).to('cuda')
        self.grid = nn.Parameter(torch.zeros([1, channels, *[int(w) for w in world_size]]))

    def forward(self, xyz):
        out = _GridSample.apply(self.grid, xyz, self.xyz_min.cpu(), self.xyz_max.cpu())
        if self.channels == 1:
            out = out.squeeze(-1)
        return out

    def scale_volume_grid(self, new_world_size):
        """model/grid.py:101-106: F.interpolate(grid, size, 'trilinear', align_corners=True) as one streaming
        kernel (fgs_grid_resize_trilinear)."""
        if self.channels == 0:
            self.grid = nn.Parameter(torch.zeros([1, self.channels, *new_world_size]))
        else:
            self.grid = nn.Parameter(grid_resize(self.grid.data, [int(w) for w in new_world_size]))
        self.world_size = new_world_size

    def total_variation_add_grad(self, wx, wy, wz, dense_mode, mask=None):
        '''Add gradients by total variation loss in-place (model/grid.py:108-119)'''
        if mask is None:
            total_variation_cuda.total_variation_add_grad(self.grid, self.grid.grad, wx, wy, wz, dense_mode)
        else:
            mask = mask.detach()
            if self.grid.size(1) > 1 and mask.size() != self.grid.size():
                mask = mask.repeat(1, self.grid.size(1), 1, 1, 1).contiguous()
            assert mask.size() == self.grid.size()
            total_variation_cuda.total_variation_add_grad_new(self.grid, self.grid.grad, mask.float(), wx, wy, wz,
                                                              dense_mode)

    def get_dense_grid(self):
        return self.grid

    @torch.no_grad()
    def __isub__(self, val):
        self.grid.data -= val
        return self

    def extra_repr(self):
        return f'channels={self.channels}, world_size={list(self.world_size)}'


class MaskGrid(nn.Module):
    """Known-free-space lookup on a boolean occupancy grid (model/grid.py:253-287).  Kept exactly, because state_dicts
    and callers name them: the constructor keywords and the buffers `mask`, `xyz2ijk_scale`, `xyz2ijk_shift` (registered
    in that order); the lookup itself is fgs_maskcache_lookup."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is None:
            occupied = mask.bool()
        else:
            # checkpoint of an earlier stage: occupied where the 3^3 max-pooled alpha reaches the threshold
            ckpt = torch.load(path)
            self.mask_cache_thres = mask_cache_thres
            pooled = F.max_pool3d(ckpt['model_state_dict']['alpha.grid'], kernel_size=3, padding=1, stride=1)
            occupied = (pooled >= mask_cache_thres)[0, 0]
            xyz_min, xyz_max = (ckpt['model_kwargs'][k] for k in ('xyz_min', 'xyz_max'))
        lo, hi = torch.Tensor(xyz_min), torch.Tensor(xyz_max)
        scale = (torch.Tensor(list(occupied.shape)) - 1) / (hi - lo)
        self.register_buffer('mask', occupied)
        self.register_buffer('xyz2ijk_scale', scale)
        self.register_buffer('xyz2ijk_shift', -lo * scale)

    @torch.no_grad()
    def forward(self, xyz):
        """xyz [..., 3] world coordinates -> bool [...]: True where the point's nearest cell is occupied."""
        pts = xyz.reshape(-1, 3).contiguous()
        hit = render_utils_cuda.maskcache_lookup(self.mask, pts, self.xyz2ijk_scale, self.xyz2ijk_shift)
        return hit.reshape(xyz.shape[:-1])

    def extra_repr(self):
        return f'mask.shape={list(self.mask.shape)}'
