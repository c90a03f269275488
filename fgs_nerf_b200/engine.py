"""Fused fine-stage step: host side of include/fgs_b200_fine.h.

`FineStepEngine` owns the parameters (sdf / k0 grids, rgbnet / refnet), the Adam state and every
workspace buffer as torch CUDA tensors, and drives one training iteration of the reference
(model/nerf_training.py:300-373, fine stage) as a fixed sequence of launches of our sm_100a
kernels through the C-ABI -- no host synchronisation inside a step, no torch compute ops on the
path (torch is used for allocation, memsets and, multi-GPU, the NCCL all-reduce).

  forward()      nerf.forward_fine                                  model/nerf.py:776-941
  train_step()   forward + loss + backward + TV + MaskedAdam        model/nerf_training.py:300-373
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from . import dist as fdist
from ._lib import check, lib, ptr, stream

MLP_LAYERS = 8
WIDTH = 256
X0_DIM, X0_LD = 106, 112
XR_DIM, XR_LD = 307, 320
K0_DIM = 12
NUM_COUNTERS = 8
CNT = dict(M1=0, M2=1, OVERFLOW=2, MA=3, TOTAL_LO=4, TOTAL_HI=5, M0=6, STICKY=7)
F_ALPHA, F_VISIT, F_L2 = 1, 2, 4

_fp = C.c_void_p


class FgsFineParams(C.Structure):
    """Mirror of `struct FgsFineParams` (include/fgs_b200_fine.h) -- keep field order in step."""
    _fields_ = [
        ("n_rays", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32),
        ("mX", C.c_int32), ("mY", C.c_int32), ("mZ", C.c_int32),
        ("cap_words", C.c_int32), ("cap_m1", C.c_int32), ("cap_m2", C.c_int32),
        ("render_depth", C.c_int32), ("render_normal", C.c_int32),
        ("skip_composite", C.c_int32), ("mlp_grads_zeroed", C.c_int32),
        ("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3),
        ("mask_xyz_min", C.c_float * 3), ("mask_xyz_max", C.c_float * 3),
        ("near_", C.c_float), ("far_", C.c_float), ("stepdist", C.c_float), ("dist", C.c_float),
        ("voxel_size", C.c_float), ("mask_thres", C.c_float), ("fast_color_thres", C.c_float),
        ("inv_s", C.c_float), ("bg", C.c_float), ("pad0_", C.c_float),
        ("rays_o", _fp), ("rays_d", _fp), ("viewdirs", _fp),
        ("sdf", _fp), ("k0", _fp), ("mask_grid", _fp),
        ("W", _fp * MLP_LAYERS), ("B", _fp * MLP_LAYERS),
        ("t_min", _fp), ("t_max", _fp),
        ("n_steps", _fp), ("cnt1", _fp), ("cnt2", _fp), ("off1", _fp), ("off2", _fp),
        ("bits", _fp), ("alphainv_last", _fp),
        ("ray_id1", _fp), ("step_id1", _fp), ("flag1", _fp),
        ("sdf1", _fp), ("grad1", _fp), ("alpha1", _fp), ("w1", _fp), ("T1", _fp),
        ("d_sdf1", _fp), ("d_grad1", _fp), ("d_w1", _fp),
        ("idx1_of2", _fp), ("ray_id2", _fp),
        ("X0", _fp), ("H", _fp), ("XR", _fp), ("logit", _fp), ("normal2", _fp),
        ("rgb_marched", _fp), ("sigmoid_rgb", _fp), ("cum_weights", _fp), ("depth", _fp),
        ("normal_marched", _fp), ("counters", _fp),
        ("g_sdf", _fp), ("g_k0", _fp), ("gW", _fp * MLP_LAYERS), ("gB", _fp * MLP_LAYERS),
        ("dA", _fp), ("dB", _fp), ("dXR", _fp), ("dX0", _fp), ("dlogit", _fp), ("dlogit_p", _fp), ("dXRp", _fp),
        ("w_split", _fp), ("hbits", _fp), ("g_mlp_flat", _fp), ("g_mlp_floats", C.c_int64), ("k0_brick_flags", _fp),
        ("dZ", _fp), ("k0_vox_flags", _fp), ("mask_occ", _fp), ("sdf_brick_flags", _fp),
        ("inc_mask", _fp), ("inc_scale", C.c_float * 3), ("inc_shift", C.c_float * 3),
        ("iX", C.c_int32), ("iY", C.c_int32), ("iZ", C.c_int32), ("pad2_", C.c_int32),
    ]


class FgsFineLoss(C.Structure):
    _fields_ = [
        ("target", _fp),
        ("weight_main", C.c_float), ("weight_entropy_last", C.c_float),
        ("weight_orientation", C.c_float), ("sigmoid_rgb_loss", C.c_float),
        ("n_rays_global", C.c_int32), ("has_last_ray", C.c_int32),
        ("g_alphainv_last", _fp), ("loss_out", _fp),
        ("weight_rgbper", C.c_float), ("rgbper_slot", C.c_int32),
    ]


class FgsFineStep(C.Structure):
    """include/fgs_b200_fine.h: FgsFineStep (same field order and types)."""
    _fields_ = [
        ("tv_smooth", C.c_int), ("smooth_weight", C.c_float), ("smooth_n_sel", C.c_double),
        ("nonempty", _fp), ("tv_scratch", _fp),
        ("tv_sdf", C.c_int), ("sdf_tv_weight", C.c_float), ("sdf_tv_dense", C.c_int),
        ("adam_step", C.c_int), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("lr_sdf", C.c_float), ("lr_k0", C.c_float),
        ("sdf_m", _fp), ("sdf_v", _fp), ("k0_m", _fp), ("k0_v", _fp),
        ("sparse_k0", C.c_int), ("brick_list", _fp),
        ("n_mlp", C.c_int), ("mlp_p", _fp), ("mlp_g", _fp), ("mlp_m", _fp), ("mlp_v", _fp), ("mlp_n", _fp),
        ("mlp_steps", _fp), ("mlp_lrs", _fp),
        ("side_stream", _fp),
    ]


_PP = C.POINTER(FgsFineParams)
_LP = C.POINTER(FgsFineLoss)
_lib.declare("fgs_fine_train_step", C.c_int, [_PP, _LP, C.POINTER(FgsFineStep), _fp])
_lib.declare("fgs_fine_train_step_graph", C.c_int, [_PP, _LP, C.POINTER(FgsFineStep), _fp, C.POINTER(C.c_int)])
_lib.declare("fgs_fine_forward", C.c_int, [_PP, _fp])
_lib.declare("fgs_fine_loss", C.c_int, [_PP, _LP, _fp])
_lib.declare("fgs_fine_backward", C.c_int, [_PP, _LP, _fp])
_lib.declare("fgs_fine_backward_phase", C.c_int, [_PP, _LP, C.c_int, _fp])
_lib.declare("fgs_set_dw_sm_reserve", C.c_int, [C.c_int])
_lib.declare("fgs_smooth_grad_tv_add_grad", C.c_int,
             [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_double, _fp, _fp])
_f3 = C.POINTER(C.c_float)
_lib.declare("fgs_grid_sample_fwd", C.c_int,
             [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int64, _f3, _f3, _fp, _fp])
_lib.declare("fgs_grid_sample_bwd", C.c_int,
             [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int64, _f3, _f3, _fp, _fp])
_lib.declare("fgs_fine_wsplit_floats", C.c_int64, [])
_lib.declare("fgs_mask_occupancy_words", C.c_int64, [C.c_int, C.c_int, C.c_int])
_lib.declare("fgs_mask_occupancy", C.c_int, [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp])
_lib.declare("fgs_gemm_tn", C.c_int, [C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, _fp, _fp])
_lib.declare("fgs_gemm_tn_bf16", C.c_int, [C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, _fp, _fp])
_lib.declare("fgs_gemm_tcw_bf16", C.c_int,
             [C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, _fp, C.c_int, _fp, _fp])
_lib.declare("fgs_debug_gemm_phases", C.c_int, [_fp])
_lib.declare("fgs_debug_feat_task", C.c_int, [C.c_int])
_i = C.c_int
_lib.declare("fgs_grid_resize_trilinear", C.c_int, [_fp, _i, _i, _i, _i, _fp, _i, _i, _i, _i, _fp])
_lib.declare("fgs_grid_sample_lattice", C.c_int,
             [_fp, _i, _i, _i, _fp, _fp, _fp, _i, _fp, _i, _fp, _i, C.c_float, _fp, _fp])
_lib.declare("fgs_conv3d_replicate", C.c_int, [_fp, _fp, _i, _i, _i, _fp, _i, C.c_float, C.c_float, _fp])
_lib.declare("fgs_gather_rays", C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp, _i, _fp, _fp, _fp, _fp, _fp, _fp])
_lib.declare("fgs_gemm_tcw", C.c_int,
             [C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int,
              _fp, _fp, _fp, _fp])


def _c3(t):
    a = (C.c_float * 3)()
    for i in range(3):
        a[i] = float(t[i])
    return a


def grid_sample(grid, xyz, xyz_min, xyz_max):
    """DenseGrid.forward (model/grid.py:49-59) / MaskCache sampling (model/nerf.py:1202-1207):
    grid [1,C,X,Y,Z] (contiguous or channels_last_3d), xyz [...,3] -> [...,C]."""
    assert grid.is_cuda and grid.dtype == torch.float32 and grid.dim() == 5 and grid.size(0) == 1
    Cn, X, Y, Z = grid.shape[1:]
    cl = 0
    if not grid.is_contiguous():
        if not grid.is_contiguous(memory_format=torch.channels_last_3d):
            raise RuntimeError("grid must be contiguous or channels_last_3d")
        cl = 1
    shape = xyz.shape[:-1]
    pts = xyz.reshape(-1, 3).contiguous()
    out = torch.empty((pts.shape[0], Cn), dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        check(lib.fgs_grid_sample_fwd(ptr(grid), Cn, X, Y, Z, cl, ptr(pts), pts.shape[0],
                                      _c3(xyz_min), _c3(xyz_max), ptr(out), stream()))
    return out.reshape(*shape, Cn)


def grid_sample_backward(grid_shape, channels_last, xyz, xyz_min, xyz_max, grad_out, grad_grid=None):
    Cn, X, Y, Z = grid_shape[1:]
    pts = xyz.reshape(-1, 3).contiguous()
    go = grad_out.reshape(-1, Cn).contiguous()
    if grad_grid is None:
        grad_grid = torch.zeros(grid_shape, dtype=torch.float32, device=xyz.device)
        if channels_last:
            grad_grid = grad_grid.contiguous(memory_format=torch.channels_last_3d)
    with torch.cuda.device(xyz.device):
        check(lib.fgs_grid_sample_bwd(ptr(grad_grid), Cn, X, Y, Z, int(channels_last), ptr(pts), pts.shape[0],
                                      _c3(xyz_min), _c3(xyz_max), ptr(go), stream()))
    return grad_grid


def grid_resize(grid, new_world_size):
    """DenseGrid.scale_volume_grid (model/grid.py:101-106): F.interpolate(grid [1,C,X,Y,Z], size, 'trilinear',
    align_corners=True) as one streaming kernel.  A channels_last_3d grid (the engine's k0) stays channels_last_3d."""
    _lib.check_input(grid if grid.is_contiguous() else grid.permute(0, 2, 3, 4, 1), "grid", torch.float32)
    if grid.dim() != 5 or grid.shape[0] != 1:
        raise RuntimeError(f"grid must be [1,C,X,Y,Z], got {tuple(grid.shape)}")
    Cn, X, Y, Z = (int(v) for v in grid.shape[1:])
    X2, Y2, Z2 = (int(v) for v in new_world_size)
    cl = int(not grid.is_contiguous())
    out = torch.empty((1, Cn, X2, Y2, Z2), dtype=torch.float32, device=grid.device,
                      memory_format=torch.channels_last_3d if cl else torch.contiguous_format)
    with torch.cuda.device(grid.device):
        check(lib.fgs_grid_resize_trilinear(ptr(grid), Cn, X, Y, Z, ptr(out), X2, Y2, Z2, cl, stream()))
    return out


def gaussian_3dconv_weights(ksize=3, sigma=1, device="cpu"):
    """The kernel of nerf._gaussian_3dconv (model/nerf.py:260-268), built step for step as the reference builds it:
    float64 exp on an integer lattice, cast to fp32 on the grid's device, divided by its fp32 sum there.
    Returns a CPU fp32 tensor [k,k,k]."""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    kernel = np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2))
    kernel = torch.from_numpy(kernel).to(device=device, dtype=torch.float32)
    return (kernel / kernel.sum()).cpu().contiguous()


def smooth_conv(grid, ksize=3, sigma=1, pre_div=1.0, post_div=1.0, weights=None):
    """nerf._gaussian_3dconv(ksize, sigma)(grid) for a [1,1,X,Y,Z] volume (Conv3d, replicate padding) as one
    shared-memory tiled launch; pre_div / post_div give m(grid / pre_div) / post_div in the same pass."""
    _lib.check_input(grid, "grid", torch.float32)
    if grid.dim() != 5 or grid.shape[0] != 1 or grid.shape[1] != 1:
        raise RuntimeError(f"grid must be [1,1,X,Y,Z], got {tuple(grid.shape)}")
    w = gaussian_3dconv_weights(ksize, sigma, grid.device) if weights is None else weights.float().cpu().contiguous()
    if tuple(w.shape) != (ksize,) * 3:
        raise RuntimeError(f"weights must be [{ksize},{ksize},{ksize}]")
    X, Y, Z = (int(v) for v in grid.shape[2:])
    out = torch.empty_like(grid)
    with torch.cuda.device(grid.device):
        check(lib.fgs_conv3d_replicate(ptr(grid), ptr(out), X, Y, Z, w.data_ptr(), int(ksize), float(pre_div),
                                       float(post_div), stream()))
    return out


def gather_rays(rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr, sel, out=None, bad=None):
    """The four `x_tr[sel_i]` selections of the training loop (model/nerf_training.py:253-258) in one launch.
    sel: int64 [n] on the device.  Returns rays_o, rays_d, viewdirs, target [n,3]."""
    n_total = rays_o_tr.shape[0]
    for t, name in ((rays_o_tr, "rays_o_tr"), (rays_d_tr, "rays_d_tr"), (viewdirs_tr, "viewdirs_tr"), (rgb_tr, "rgb_tr")):
        _lib.check_input(t, name, torch.float32)
        if tuple(t.shape) != (n_total, 3):
            raise RuntimeError(f"{name} must be [{n_total},3], got {tuple(t.shape)}")
    _lib.check_input(sel, "sel", torch.int64)
    n = sel.numel()
    if out is None:
        out = tuple(torch.empty((n, 3), dtype=torch.float32, device=sel.device) for _ in range(4))
    with torch.cuda.device(sel.device):
        check(lib.fgs_gather_rays(ptr(rays_o_tr), ptr(rays_d_tr), ptr(viewdirs_tr), ptr(rgb_tr), n_total, ptr(sel), n,
                                  ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), ptr(bad), stream()))
    return out


def gemm(mode, A, B, out, bias=None, relu=False, rows_dev=None, M=None, N=None, K=None, bits_in=None, bits_out=None):
    """fgs_gemm_tcw on 2-D row-major fp32 CUDA tensors (include/fgs_b200_fine.h): mode 0 = A B^T, 1 = A B."""
    if mode == 0:
        m, k = A.shape; n = B.shape[0]
    elif mode == 1:
        m, k = A.shape; n = B.shape[1]
    else:
        raise ValueError("mode must be 0 or 1 (the weight-gradient product is fgs_gemm_tn)")
    M, N, K = M or m, N or n, K or k
    check(lib.fgs_gemm_tcw(mode, M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(out), out.stride(0),
                           ptr(bias), int(relu), ptr(bits_in), ptr(bits_out), ptr(rows_dev), stream()))
    return out


class FineStepEngine:
    """State + workspace of the fine-stage model for batches of `n_rays` rays on one GPU."""

    def __init__(self, scene, n_rays, device="cuda", train_cfg=None, cap_m1_per_ray=192, cap_m2_per_ray=32,
                 world_size=1, rank=0, voxel_exchange=False):
        from .scene import FINE_TRAIN
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FineStepEngine needs a CUDA device: fgs_b200 has no CPU path")
        dev = self.device
        self.cfg = dict(scene.cfg)
        self.train_cfg = dict(train_cfg if train_cfg is not None else FINE_TRAIN)
        self.n_rays = int(n_rays)
        self.world_size, self.rank = int(world_size), int(rank)
        self._ctor = dict(n_rays=int(n_rays), device=device, train_cfg=self.train_cfg, cap_m1_per_ray=cap_m1_per_ray,
                          cap_m2_per_ray=cap_m2_per_ray, world_size=int(world_size), rank=int(rank),
                          voxel_exchange=voxel_exchange)
        # learning rates of the four param groups (python floats that the training loop multiplies step by step,
        # model/nerf_training.py:389-414; create_optimizer_or_freeze_model, :9-37)
        self.lr = {k: float(self.train_cfg["lrate_" + k]) for k in ("sdf", "k0", "rgbnet", "refnet")}
        f32 = torch.float32
        # ---- parameters ----
        self.xyz_min = scene.xyz_min.clone().float()
        self.xyz_max = scene.xyz_max.clone().float()
        self.mask_xyz_min = scene.mask_xyz_min.clone().float()
        self.mask_xyz_max = scene.mask_xyz_max.clone().float()
        self.voxel_size = np.float32(scene.voxel_size)
        self.sdf = scene.sdf.to(dev, f32).contiguous().clone()                       # [1,1,X,Y,Z]
        self.k0 = scene.k0.to(dev, f32).contiguous(memory_format=torch.channels_last_3d).clone()  # [1,12,X,Y,Z], C innermost
        assert self.k0.shape[1] == K0_DIM, "fine stage k0_dim must be 12"
        self.mask_grid = scene.mask_grid.to(dev, f32).contiguous()
        self.mask_thres = np.float32(scene.mask_thres)
        self.fast_color_thres = np.float32(self.cfg["fast_color_thres"])
        layers = list(scene.rgbnet) + list(scene.refnet)
        assert len(layers) == MLP_LAYERS
        dims = [(WIDTH, X0_DIM)] + [(WIDTH, WIDTH)] * 3 + [(WIDTH, XR_DIM)] + [(WIDTH, WIDTH)] * 2 + [(3, WIDTH)]
        lds = [X0_LD, WIDTH, WIDTH, WIDTH, XR_LD, WIDTH, WIDTH, WIDTH]
        # storage: input layers zero-padded to the activation row length ([256,112], [256,320]);
        # self.weights[i] are the nn.Linear-shaped views ([out,in]) of the same memory
        self.Wt, self.weights = [], []
        for (w, _), d, ld in zip(layers, dims, lds):
            assert tuple(w.shape) == d, (tuple(w.shape), d)
            st = torch.zeros(d[0], ld, dtype=f32, device=dev)
            st[:, :d[1]] = w.to(dev, f32)
            self.Wt.append(st)
            self.weights.append(st[:, :d[1]])
        self.Bt = [b.to(dev, f32).contiguous().clone() for _, b in layers]
        X, Y, Z = self.sdf.shape[2:]
        self.X, self.Y, self.Z = X, Y, Z
        self.num_voxels = X * Y * Z          # model.num_voxels (scale_volume_grid multiplies it by scale_ratio)
        self.num_voxels_base = self.num_voxels   # config value, unchanged by scale_volume_grid (model/nerf.py:67-68)
        # ---- gradients + Adam state ----
        self.g_k0 = torch.zeros_like(self.k0)       # keeps channels_last_3d
        # all 16 MLP gradient tensors are views of one flat buffer (one clear, one all-reduce bucket)
        sizes = [w.numel() for w in self.Wt] + [b.numel() for b in self.Bt]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)         # keep every view 16-byte aligned
        # one allocation [ g_sdf | MLP gradient bucket | packed touched k0 bricks ]: a data-parallel step reduces
        # everything with ONE all-reduce over a prefix of it (the packed part exists only when world_size > 1)
        V = self.sdf.numel()
        Vp = (V + 3) // 4 * 4
        nb = ((X + 3) // 4) * ((Y + 3) // 4) * ((Z + 3) // 4)
        self._xcap_bricks = min(nb, 32768) if self.world_size > 1 else 0
        # data parallel: g_sdf travels as packed 4^3 bricks too (64 floats each) behind the k0 bricks -- the dense grid
        # is 16 MB of a 44 MB all-reduce at 160^3 / 8 GPUs and 67 MB at 256^3
        self.sparse_sdf = self.world_size > 1 and (Z % 4 == 0) and not voxel_exchange
        self._xbuf = torch.zeros(Vp + offs[-1] + self._xcap_bricks * 768 + (nb * 64 if self.sparse_sdf else 0),
                                 dtype=f32, device=dev)
        self.g_sdf = self._xbuf[:V].view_as(self.sdf)
        self.g_mlp_flat = self._xbuf[Vp:Vp + offs[-1]]
        self._compact = self._xbuf[Vp + offs[-1]:]
        self._xdense = Vp + offs[-1]
        views = [self.g_mlp_flat[o:o + n] for o, n in zip(offs, sizes)]
        self.gW = [v.view_as(w) for v, w in zip(views[:MLP_LAYERS], self.Wt)]
        self.weight_grads = [g[:, :d[1]] for g, d in zip(self.gW, dims)]
        self.gB = [v.view_as(b) for v, b in zip(views[MLP_LAYERS:], self.Bt)]
        self.adam = {}
        self.adam_step = 0
        # ---- workspace ----
        N = self.n_rays
        diag = float((self.xyz_max - self.xyz_min).norm())
        stepdist = float(np.float32(self.cfg["stepsize"]) * self.voxel_size)
        self.cap_words = int(math.ceil((diag / stepdist + 2) / 32)) + 1
        self.cap_m1 = int(cap_m1_per_ray * N)
        self.cap_m2 = int((cap_m2_per_ray * N + 127) // 128 * 128)
        i32 = torch.int32
        e = lambda *s, dt=f32: torch.empty(s, dtype=dt, device=dev)
        ws = self.ws = {}
        for k in ("t_min", "t_max", "alphainv_last", "cum_weights", "depth", "g_alphainv_last"):
            ws[k] = e(N)
        for k in ("n_steps", "cnt1", "cnt2"):
            ws[k] = e(N, dt=i32)
        ws["off1"], ws["off2"] = e(N + 1, dt=i32), e(N + 1, dt=i32)
        ws["bits"] = e(N, self.cap_words, dt=i32)
        for k in ("ray_id1", "step_id1"):
            ws[k] = e(self.cap_m1, dt=i32)
        ws["flag1"] = e(self.cap_m1, dt=torch.uint8)
        for k in ("sdf1", "alpha1", "w1", "T1", "d_sdf1", "d_w1"):
            ws[k] = e(self.cap_m1)
        ws["grad1"], ws["d_grad1"] = e(3, self.cap_m1), e(3, self.cap_m1)
        ws["idx1_of2"], ws["ray_id2"] = e(self.cap_m2, dt=i32), e(self.cap_m2, dt=i32)
        ws["X0"] = e(self.cap_m2, X0_LD)
        ws["H"] = e(6, self.cap_m2, WIDTH)
        ws["XR"] = e(self.cap_m2, XR_LD)
        ws["logit"], ws["normal2"], ws["dlogit"] = e(self.cap_m2, 4), e(self.cap_m2, 4), e(self.cap_m2, 4)
        ws["dA"], ws["dB"] = e(self.cap_m2, WIDTH), e(self.cap_m2, WIDTH)
        # pre-activation gradients of the six hidden layers, each as two bf16 planes (hi, mid: the 3xBF16 operand format
        # of the backward GEMMs) in the bytes of one fp32 [cap_m2][256]; the same for the gradient at rgbnet's output
        # (dXR[:, :256]) and for dlogit.  ZERO-initialised: rows beyond the sample count are multiplied by zeros.
        ws["dZ"] = torch.zeros(6, self.cap_m2, WIDTH, dtype=f32, device=dev)
        ws["dXRp"] = torch.zeros(self.cap_m2, WIDTH, dtype=f32, device=dev)
        ws["dlogit_p"] = torch.zeros(2, self.cap_m2, 8, dtype=torch.int16, device=dev)
        ws["dXR"], ws["dX0"] = e(self.cap_m2, XR_LD), e(self.cap_m2, X0_LD)
        for k in ("rgb_marched", "sigmoid_rgb", "normal_marched"):
            ws[k] = e(N, 3)
        ws["w_split"] = torch.zeros(int(lib.fgs_fine_wsplit_floats()), dtype=f32, device=dev)
        ws["hbits"] = e(6, self.cap_m2, 8, dt=i32)
        nb = ((X + 3) // 4) * ((Y + 3) // 4) * ((Z + 3) // 4)
        # [k0 flags | sdf flags] in one allocation: the union over the ranks is ONE MAX all-reduce
        self._brick_flags2 = torch.zeros(2 * nb, dtype=torch.uint8, device=dev)
        ws["k0_brick_flags"], ws["sdf_brick_flags"] = self._brick_flags2[:nb], self._brick_flags2[nb:]
        ws["k0_brick_list"] = torch.zeros(int(lib.fgs_brick_list_ints(X, Y, Z)), dtype=i32, device=dev)
        ws["sdf_brick_list"] = torch.zeros(int(lib.fgs_brick_list_ints(X, Y, Z)) if self.sparse_sdf else 1, dtype=i32, device=dev)
        ws["k0_brick_ulist"] = torch.zeros(int(lib.fgs_brick_list_ints(X, Y, Z)) + 1, dtype=i32, device=dev)
        self.sparse_k0 = (Z % 4 == 0)      # brick-driven masked Adam (else the dense reference-style scan)
        # data parallel, optional: the k0 gradient travels at VOXEL granularity (one flag byte per voxel, ascending
        # voxel list: a third of the bytes of the brick exchange).  Off by default: the union of the voxel flags is a
        # 4 MB MAX all-reduce that, squeezed in beside the weight-gradient kernel, takes longer than the 19 MB it saves
        # (2 GPUs: 1.28 vs 1.17 ms per step); it needs bit-packed flags before it pays.
        self.voxel_exchange = bool(voxel_exchange) and self.sparse_k0 and self.world_size > 1
        if self.voxel_exchange:
            ws["k0_vox_flags"] = torch.zeros(X * Y * Z, dtype=torch.uint8, device=dev)
            ws["k0_vox_list"] = torch.zeros(int(lib.fgs_voxel_list_ints(X * Y * Z)), dtype=i32, device=dev)
            # bit masks of the touched voxels of every rank (all-gather target), when the grid allows it
            self.voxel_bits = (X * Y * Z) % 32 == 0
            if self.voxel_bits:
                ws["k0_vox_bits"] = torch.zeros(self.world_size, X * Y * Z // 32, dtype=i32, device=dev)
        self._grads_clean = True           # g_sdf / g_k0 are all-zero (fresh, or cleared by the optimiser)
        self._mlp_grads_clean = True       # the MLP gradient bucket is all-zero (fresh, or cleared by the native step's Adam)
        self._brick_list_ready = False
        self._count_host = torch.zeros(2, dtype=torch.int32).pin_memory()      # touched k0 bricks (or voxels), touched sdf bricks
        self._count_evt = torch.cuda.Event()
        self.exchange_bricks = 0
        self.exchange_sdf_bricks = 0
        self._side = None                  # side stream of the overlapped gradient exchange (data parallel)
        self.native_step = True            # single GPU: the iteration is one library call (fgs_fine_train_step)
        self.use_graph = True              # ... launched as one CUDA graph (fgs_fine_train_step_graph): 1.055 -> 1.017 ms per step
        self._graph_stream = None
        self.graph_steps = 0
        self._S = None
        self.dw_sm_reserve = 8             # SMs the weight-gradient kernel leaves to the overlapped exchange (2 GPUs: 1.220 ms per step with 8, 1.255 with 16, 1.230 with 0)
        ws["counters"] = torch.zeros(NUM_COUNTERS, dtype=i32, device=dev)
        ws["loss"] = torch.zeros(5, dtype=torch.float64, device=dev)   # mse, entropy, orientation, sigmoid_rgb, smooth_tv
        self.tv_scratch = None
        self.nonempty_mask = None
        self._keep = None
        self.P = FgsFineParams()
        self.L = FgsFineLoss()
        self._fill_static()

    # -------------------------------------------------------------------------------------
    def _fill_static(self):
        P, ws = self.P, self.ws
        P.n_rays = self.n_rays
        P.X, P.Y, P.Z = self.X, self.Y, self.Z
        P.mX, P.mY, P.mZ = self.mask_grid.shape[2:]
        P.cap_words, P.cap_m1, P.cap_m2 = self.cap_words, self.cap_m1, self.cap_m2
        for c in range(3):
            P.xyz_min[c] = float(self.xyz_min[c]); P.xyz_max[c] = float(self.xyz_max[c])
            P.mask_xyz_min[c] = float(self.mask_xyz_min[c]); P.mask_xyz_max[c] = float(self.mask_xyz_max[c])
        # model/nerf.py:689-691, 795: stepdist = dist = stepsize * voxel_size (fp32 tensor product)
        sd = np.float32(self.cfg["stepsize"]) * self.voxel_size
        P.stepdist = float(sd); P.dist = float(sd)
        P.voxel_size = float(self.voxel_size)
        P.mask_thres = float(self.mask_thres)
        P.fast_color_thres = float(self.fast_color_thres)
        P.far_ = 1e9                                      # model/nerf.py:686
        P.sdf, P.k0, P.mask_grid = ptr(self.sdf), ptr(self.k0), ptr(self.mask_grid)
        for i in range(MLP_LAYERS):
            P.W[i], P.B[i] = ptr(self.Wt[i]), ptr(self.Bt[i])
            P.gW[i], P.gB[i] = ptr(self.gW[i]), ptr(self.gB[i])
        P.g_sdf, P.g_k0 = ptr(self.g_sdf), ptr(self.g_k0)
        P.g_mlp_flat, P.g_mlp_floats = ptr(self.g_mlp_flat), self.g_mlp_flat.numel()
        for k in ("t_min", "t_max", "n_steps", "cnt1", "cnt2", "off1", "off2", "bits", "alphainv_last",
                  "ray_id1", "step_id1", "flag1", "sdf1", "grad1", "alpha1", "w1", "T1", "d_sdf1", "d_grad1",
                  "d_w1", "idx1_of2", "ray_id2", "X0", "H", "XR", "logit", "normal2", "rgb_marched",
                  "sigmoid_rgb", "cum_weights", "depth", "normal_marched", "counters", "dA", "dB", "dXR",
                  "dX0", "dlogit", "w_split", "hbits", "k0_brick_flags", "dZ", "dXRp", "dlogit_p"):
            setattr(P, k, ptr(ws[k]))
        P.k0_vox_flags = ptr(ws["k0_vox_flags"]) if self.voxel_exchange else None
        P.sdf_brick_flags = ptr(ws["sdf_brick_flags"]) if self.sparse_sdf else None
        self._build_mask_occupancy()

    def _build_mask_occupancy(self):
        """One bit per cell of the MaskCache grid (any corner non-zero): lets the march kernel skip the 8 gathers of a
        sample whose cell is empty -- exact, because such a sample interpolates to 0 < mask_thres.  Call again after
        changing self.mask_grid in place."""
        mX, mY, mZ = (int(v) for v in self.mask_grid.shape[2:])
        if float(self.mask_thres) <= 0.0 or min(mX, mY, mZ) < 2:
            self.P.mask_occ = None
            return
        self.mask_occ = torch.zeros(int(lib.fgs_mask_occupancy_words(mX, mY, mZ)), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.fgs_mask_occupancy(ptr(self.mask_grid), mX, mY, mZ, ptr(self.mask_occ), stream()))
        self.P.mask_occ = ptr(self.mask_occ)

    def s_val(self, global_step):
        """model/nerf.py:514 (python floats)."""
        c = self.cfg
        return 1.0 / (global_step + c["s_ratio"] / c["s_start"] - c["step_start"]) * c["s_ratio"]

    def _bind_inputs(self, rays_o, rays_d, viewdirs, global_step, near, bg, render_depth, render_grad):
        for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (viewdirs, "viewdirs")):
            _lib.check_input(t, n, torch.float32)
            if tuple(t.shape) != (self.n_rays, 3):
                raise RuntimeError(f"{n} must be [{self.n_rays},3], got {tuple(t.shape)}")
        P = self.P
        P.rays_o, P.rays_d, P.viewdirs = ptr(rays_o), ptr(rays_d), ptr(viewdirs)
        self._keep = (rays_o, rays_d, viewdirs)
        P.near_ = float(np.float32(near))
        P.bg = float(bg)
        # inv_s = ones(1) / (ones(1) * s_val)   (model/nerf.py:514-523), fp32
        P.inv_s = float(np.float32(1.0) / np.float32(self.s_val(global_step)))
        P.render_depth, P.render_normal = int(bool(render_depth)), int(bool(render_grad))

    # -------------------------------------------------------------------------------------
    def forward(self, rays_o, rays_d, viewdirs, global_step=20000, near=2.0, bg=1.0, render_depth=False,
                render_grad=False, _skip_composite=False):
        """nerf.forward_fine: enqueue only.  Results live in self.ws (see outputs())."""
        self._bind_inputs(rays_o, rays_d, viewdirs, global_step, near, bg, render_depth, render_grad)
        self.P.skip_composite = int(bool(_skip_composite))
        with torch.cuda.device(self.device):
            check(lib.fgs_fine_forward(C.byref(self.P), stream()))

    def counters(self):
        """Host copy of the device counters (synchronises)."""
        c = self.ws["counters"].cpu().numpy().astype(np.int64)
        out = {k: int(c[i]) for k, i in CNT.items() if not k.startswith("TOTAL") and k != "STICKY"}
        out["total_len"] = int((c[CNT["TOTAL_HI"]] << 32) | (c[CNT["TOTAL_LO"]] & 0xffffffff))
        if out["OVERFLOW"]:
            raise RuntimeError(f"fgs_b200 fine step: workspace overflow (bits {out['OVERFLOW']:#x}: "
                               f"1=steps/ray, 2=cap_m1 {self.cap_m1}, 4=cap_m2 {self.cap_m2}); counts {out}")
        return out

    def check_overflow(self, clear=True):
        """Synchronises.  Raises if ANY forward since the last check ran out of workspace: the per-forward overflow word
        is cleared by every forward, so an asynchronous loop that only looks every i_print iterations would miss the
        others; counters[STICKY] is only cleared here.  (An overflowing ray's samples are dropped by the kernels, i.e.
        the gradients of that iteration were wrong.)"""
        word = self.ws["counters"][CNT["STICKY"]:CNT["STICKY"] + 1]
        bits = int(word.item())
        if clear and bits:
            word.zero_()
        if bits:
            raise RuntimeError(f"fgs_b200 fine step: workspace overflow in an earlier iteration (bits {bits:#x}: "
                               f"1=steps/ray > {self.cap_words * 32}, 2=cap_m1 {self.cap_m1}, 4=cap_m2 {self.cap_m2}); "
                               f"raise cap_m1_per_ray / cap_m2_per_ray")

    def outputs(self):
        """The reference's render_result dict (model/nerf.py:922-939) from the workspace;
        exact-size tensors, so this synchronises once to read M2."""
        cnt = self.counters()
        ws = self.ws
        M1, M2 = cnt["M1"], cnt["M2"]
        idx1 = ws["idx1_of2"][:M2].long()
        ray_id = ws["ray_id2"][:M2].long()
        rgb = torch.sigmoid(ws["logit"][:M2, :3])
        flag = ws["flag1"][:M1]
        keep_a = (flag & F_ALPHA) != 0
        depth = ws["depth"] if self.P.render_depth else None
        return {
            "alphainv_cum": ws["alphainv_last"], "weights": ws["w1"][idx1], "ray_id": ray_id,
            "step_id": ws["step_id1"][idx1].long(), "viewdirs": self._keep[2][ray_id],
            "rgb_marched": ws["rgb_marched"], "sigmoid_rgb": ws["sigmoid_rgb"],
            "normal_marched": ws["normal_marched"] if self.P.render_normal else None,
            "normal": ws["normal2"][:M2, :3], "raw_alpha": ws["alpha1"][idx1], "raw_rgb": rgb,
            "depth": depth, "disp": (1 / depth) if depth is not None else None,
            "mask": ((flag & F_L2) != 0)[keep_a], "gradient": ws["grad1"][:, idx1].T,
            "s_val": None, "counts": cnt,
            "level1": dict(ray_id=ws["ray_id1"][:M1].long(), step_id=ws["step_id1"][:M1].long(),
                           sdf=ws["sdf1"][:M1], gradient=ws["grad1"][:, :M1].T, alpha=ws["alpha1"][:M1],
                           flag=flag),
        }

    # -------------------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, rays_o, rays_d, viewdirs, global_step=20000, near=2.0, bg=1.0, render_depth=True,
               render_grad=True, world_size=1, rank=0):
        """Full-image rendering (model/utils.py:249-441, render_viewpoints): the rays [R,3] are cut into
        chunks of self.n_rays (the reference uses 8192, utils.py:296-299) and each chunk runs the fused
        forward.  The per-ray outputs of a chunk are written by the kernels STRAIGHT into the image-sized
        result buffers (the output pointers of the parameter block are re-aimed per chunk: no copies).
        With world_size > 1 the chunks are dealt round-robin to the ranks and every result is
        all-gathered once per image -- each rank sends only the pixels it rendered (rendering needs no
        other exchange).
        Returns rgb_marched, sigmoid_rgb, depth, disp, normal_marched, alphainv_cum, cum_weights."""
        R = rays_o.shape[0]
        N = self.n_rays
        n_chunks = (R + N - 1) // N
        n_local = fdist.local_chunk_count(n_chunks, world_size)    # chunks per rank (the last ranks may idle once)
        dev = self.device
        keys = [("rgb_marched", 3), ("sigmoid_rgb", 3), ("alphainv_last", 1), ("cum_weights", 1)]
        if render_depth:
            keys.append(("depth", 1))
        if render_grad:
            keys.append(("normal_marched", 3))
        pad = n_chunks * N - R
        if pad:                                                     # last chunk: repeat the last ray
            rays_o, rays_d, viewdirs = (torch.cat([t, t[-1:].expand(pad, 3)]) for t in (rays_o, rays_d, viewdirs))
        rays_o, rays_d, viewdirs = (t.contiguous() for t in (rays_o, rays_d, viewdirs))
        out = {k: torch.zeros((n_local * N, c), dtype=torch.float32, device=dev) for k, c in keys}
        ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        P = self.P
        saved = {k: getattr(P, k) for k, _ in keys}
        try:
            for li, ci in enumerate(fdist.chunks_of_rank(n_chunks, world_size, rank)):
                a = ci * N
                for k, c in keys:
                    setattr(P, k, out[k].data_ptr() + li * N * c * 4)
                self.forward(rays_o[a:a + N], rays_d[a:a + N], viewdirs[a:a + N], global_step=global_step, near=near,
                             bg=bg, render_depth=render_depth, render_grad=render_grad)
            ovf |= self.ws["counters"][CNT["STICKY"]]
        finally:
            for k, v in saved.items():
                setattr(P, k, v)
        if world_size > 1:
            import torch.distributed as dist
            for k, c in keys:
                out[k] = fdist.gather_dealt_chunks(out[k], N, world_size)
            dist.all_reduce(ovf, op=dist.ReduceOp.MAX)
        bits = int(ovf.item())
        if bits:
            self.ws["counters"][CNT["STICKY"]] = 0
            raise RuntimeError(f"fgs_b200 render: workspace overflow (bits {bits:#x}); raise cap_m1_per_ray / cap_m2_per_ray")
        res = {k: v[:R] for k, v in out.items()}
        res["alphainv_cum"] = res.pop("alphainv_last")[:, 0]
        res["cum_weights"] = res["cum_weights"][:, 0]
        if render_depth:
            res["depth"] = res["depth"][:, 0]
            res["disp"] = 1 / res["depth"]
        return res

    def extract_fields(self, bound_min, bound_max, resolution=128):
        """The SDF field nerf.extract_geometry queries (model/nerf.py:1157-1170 -> extract_geometry.py:5-19):
        -trilinear(sdf) on a resolution^3 lattice, one launch, device tensor [res,res,res]."""
        from .extract_geometry import extract_sdf_field
        return extract_sdf_field(self.sdf, self.xyz_min, self.xyz_max, bound_min, bound_max, resolution)

    def set_nonempty_mask(self):
        """nerf._set_nonempty_mask (model/nerf.py:339-353): MaskCache at the voxel centres."""
        X, Y, Z = self.X, self.Y, self.Z
        g = torch.stack(torch.meshgrid(
            torch.linspace(float(self.xyz_min[0]), float(self.xyz_max[0]), X),
            torch.linspace(float(self.xyz_min[1]), float(self.xyz_max[1]), Y),
            torch.linspace(float(self.xyz_min[2]), float(self.xyz_max[2]), Z), indexing="ij"), -1).to(self.device)
        v = grid_sample(self.mask_grid, g, self.mask_xyz_min, self.mask_xyz_max)[..., 0]
        self.nonempty_mask = (v >= float(self.mask_thres)).contiguous()
        self.n_nonempty = int(self.nonempty_mask.sum().item())
        return self.nonempty_mask

    def is_tv_step(self, global_step):
        t = self.train_cfg
        return t["tv_from"] < global_step < t["tv_end"] and global_step % t["tv_every"] == 0

    def zero_grad(self):
        """optimizer.zero_grad(): a no-op when the optimiser kernels already cleared what they consumed."""
        if self._grads_clean:
            return
        with torch.cuda.device(self.device):
            st = stream()
            check(lib.fgs_memset_zero(ptr(self.g_sdf), self.g_sdf.numel() * 4, st))
            check(lib.fgs_memset_zero(ptr(self.g_k0), self.g_k0.numel() * 4, st))
            check(lib.fgs_memset_zero(ptr(self._brick_flags2), self._brick_flags2.numel(), st))
            if self.voxel_exchange:
                check(lib.fgs_memset_zero(ptr(self.ws["k0_vox_flags"]), self.ws["k0_vox_flags"].numel(), st))
        self._grads_clean = True

    def loss_backward(self, target, n_rays_global=None, has_last_ray=True, phase=0):
        """Loss block + backward (model/nerf_training.py:307-347); grads of the MLPs are overwritten,
        g_sdf / g_k0 accumulated into (call zero_grad() first).  phase=1 stops after the grid gradients are
        complete (the weight gradients follow with backward_weights())."""
        L = self._fill_loss(target, n_rays_global, has_last_ray)
        self.ws["loss"].zero_()
        self._grads_clean = False
        self._mlp_grads_clean = False
        self.P.mlp_grads_zeroed = 0
        with torch.cuda.device(self.device):
            st = stream()
            check(lib.fgs_fine_loss(C.byref(self.P), C.byref(L), st))
            check(lib.fgs_fine_backward_phase(C.byref(self.P), C.byref(L), int(phase), st))

    def _fill_loss(self, target, n_rays_global, has_last_ray):
        t = self.train_cfg
        L = self.L
        _lib.check_input(target, "target", torch.float32)
        L.target = ptr(target)
        self._keep_t = target
        L.weight_main, L.weight_entropy_last = t["weight_main"], t["weight_entropy_last"]
        L.weight_orientation, L.sigmoid_rgb_loss = t["weight_orientation"], t["sigmoid_rgb_loss"]
        L.n_rays_global = int(n_rays_global or self.n_rays)
        L.has_last_ray = int(bool(has_last_ray))
        L.g_alphainv_last = ptr(self.ws["g_alphainv_last"])
        L.loss_out = ptr(self.ws["loss"])
        return L

    def backward_weights(self):
        """Second phase of loss_backward(phase=1): the MLP weight / bias gradients."""
        self._mlp_grads_clean = False
        self.P.mlp_grads_zeroed = 0
        with torch.cuda.device(self.device):
            check(lib.fgs_fine_backward_phase(C.byref(self.P), C.byref(self.L), 2, stream()))

    def smooth_grad_tv_add_grad(self):
        """weight_tv_density * smooth_grad_tv term (model/nerf_training.py:330-336, nerf.py:438-446)."""
        t = self.train_cfg
        if self.tv_scratch is None:
            self.tv_scratch = torch.empty(6 * self.X * self.Y * self.Z, dtype=torch.float32, device=self.device)
        if self.nonempty_mask is not None:
            n_sel, m = 3.0 * self.n_nonempty, ptr(self.nonempty_mask)
        else:
            n_sel, m = 3.0 * self.X * self.Y * self.Z, None
        w = t["weight_tv_density"] * t["smooth_grad_tv"]
        with torch.cuda.device(self.device):
            check(lib.fgs_smooth_grad_tv_add_grad(ptr(self.sdf), ptr(self.g_sdf), m, ptr(self.tv_scratch), self.X,
                                                  self.Y, self.Z, float(self.voxel_size), float(w), float(n_sel),
                                                  self.ws["loss"].data_ptr() + 4 * 8, stream()))

    def sdf_tv_add_grad(self, global_step, n_rays_global=None):
        """model/nerf_training.py:353-360 + nerf.sdf_total_variation_add_grad (nerf.py:465-467)."""
        t = self.train_cfg
        w = t["weight_tv_density"] * t["sdf_tv"] / (n_rays_global or self.n_rays)
        w = float(w * max(self.X, self.Y, self.Z) / 128)
        dense = global_step < t["tv_dense_before"]
        with torch.cuda.device(self.device):
            check(lib.fgs_total_variation_add_grad(ptr(self.sdf), ptr(self.g_sdf), None, w, w, w, 1, self.X, self.Y,
                                                   self.Z, int(dense), 0, stream()))

    def reset_optimizer(self, global_step=0):
        """create_optimizer_or_freeze_model (model/nerf_training.py:9-37): fresh MaskedAdam state, learning rates
        back to lrate_* x 0.1^(global_step / (lrate_decay * 1000))."""
        t = self.train_cfg
        decay = 0.1 ** (global_step / (t.get("lrate_decay", 20) * 1000))
        self.lr = {k: float(t["lrate_" + k]) * decay for k in ("sdf", "k0", "rgbnet", "refnet")}
        self.adam = {}
        self.adam_step = 0
        self.ws["k0_brick_ulist"][:2].zero_()   # its two counters alternate with the parity of adam_step
        self._S = None                      # cached pointers of the Adam state

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels):
        """nerf.scale_volume_grid (model/nerf.py:369-381) with _set_grid_resolution (:298-303): new voxel size and
        world size from `num_voxels`, SDF and k0 resampled trilinearly (fgs_grid_resize_trilinear), non-empty mask
        recomputed.  Every grid-sized buffer changes shape, so this returns a NEW engine that shares the MLPs'
        values; the optimiser starts fresh like the reference's (nerf_training.py:254)."""
        from .scene import Scene
        ext = (self.xyz_max - self.xyz_min).float()
        voxel_size = (ext.prod() / num_voxels).pow(1 / 3)                 # fp32 tensor arithmetic, as the reference
        world_size = (ext / voxel_size).long().tolist()
        sdf = grid_resize(self.sdf, world_size)
        k0 = grid_resize(self.k0, world_size)
        sc = Scene(G=int(world_size[0]), xyz_min=self.xyz_min, xyz_max=self.xyz_max, voxel_size=float(voxel_size),
                   sdf=sdf, k0=k0, mask_grid=self.mask_grid, mask_xyz_min=self.mask_xyz_min,
                   mask_xyz_max=self.mask_xyz_max, mask_thres=float(self.mask_thres),
                   rgbnet=[(w, b) for w, b in zip(self.weights[:4], self.Bt[:4])],
                   refnet=[(w, b) for w, b in zip(self.weights[4:], self.Bt[4:])], cfg=dict(self.cfg))
        had_mask = self.nonempty_mask is not None
        ctor = dict(self._ctor)
        # drop this engine's grid-sized buffers before the new ones are allocated (256^3: 0.8 GB per k0-sized tensor)
        self.sdf = self.k0 = self.g_k0 = self._xbuf = self.g_sdf = self._compact = None
        self.adam, self.tv_scratch, self.nonempty_mask = {}, None, None
        new = FineStepEngine(sc, **ctor)
        new.num_voxels = num_voxels
        new.num_voxels_base = self.num_voxels_base
        if had_mask:
            new.set_nonempty_mask()
        new.reset_optimizer(0)
        return new

    @torch.no_grad()
    def init_sdf_from_sdf(self, sdf0, smooth=False, reduce=1., ksize=3, sigma=1., smooth_scale=True):
        """nerf.init_sdf_from_sdf (model/nerf.py:280-296), the stage hand-off of nerf_training.py:122-129: the earlier
        stage's SDF grid resampled to this resolution, divided by `reduce` (the reference divides twice when `smooth`),
        optionally smoothed, then -- `smooth_scale`, the reference's default -- smoothed with the 5^3, sigma 1 Gaussian.
        Writes into the existing self.sdf storage (the C structs keep their pointers)."""
        sdf0 = sdf0.to(self.device, torch.float32).contiguous()
        if tuple(sdf0.shape) != tuple(self.sdf.shape):
            sdf0 = grid_resize(sdf0, (self.X, self.Y, self.Z))
        if smooth:
            data = smooth_conv(sdf0, ksize, sigma, pre_div=reduce, post_div=reduce)
        else:
            data = smooth_conv(sdf0, 1, 1, pre_div=reduce, weights=torch.ones(1, 1, 1))     # sdf0 / reduce
        if self.nonempty_mask is not None:
            self.set_nonempty_mask()
        if smooth_scale:
            data = smooth_conv(data, 5, 1)
        self.sdf.copy_(data)

    def _ensure_adam(self):
        """Adam state (allocated and zeroed on the calling stream, before any side-stream use)."""
        if self.adam:
            return
        self.adam = {"sdf": (torch.zeros_like(self.sdf), torch.zeros_like(self.sdf)),
                     "k0": (torch.zeros_like(self.k0), torch.zeros_like(self.k0)),
                     "W": [(torch.zeros_like(w), torch.zeros_like(w)) for w in self.Wt],
                     "B": [(torch.zeros_like(b), torch.zeros_like(b)) for b in self.Bt]}
        n = 2 * MLP_LAYERS
        self._S = None
        self._mt = dict(
            p=(_fp * n)(*[ptr(x) for x in self.Wt + self.Bt]),
            g=(_fp * n)(*[ptr(x) for x in self.gW + self.gB]),
            m=(_fp * n)(*[ptr(x[0]) for x in self.adam["W"] + self.adam["B"]]),
            v=(_fp * n)(*[ptr(x[1]) for x in self.adam["W"] + self.adam["B"]]),
            n=(C.c_int64 * n)(*[x.numel() for x in self.Wt + self.Bt]))

    def optimizer_step(self, lr_scale=1.0, parts="all"):
        """MaskedAdam.step (model/adam.py:190-221) with the fine-stage param groups
        (model/nerf_training.py:9-37): k0 masked (skip_zero_grad), sdf and the 16 MLP tensors dense.
        parts: "all", or "grids" followed by "mlp" (train_step runs the grid part on a side stream while the
        weight gradients are still being computed)."""
        t = self.train_cfg
        if parts != "mlp":
            self.adam_step += 1
        step = self.adam_step
        self._ensure_adam()
        b1, b2, eps = 0.9, 0.99, 1e-8
        with torch.cuda.device(self.device):
            st = stream()
            if parts == "mlp":
                self._adam_mlp(step, lr_scale, b1, b2, eps, st)
                return
            m, v = self.adam["sdf"]
            # sdf: dense Adam (not in skip_zero_grad_fields), gradient cleared by the same kernel
            check(lib.fgs_adam_upd_zero_grad(ptr(self.sdf), ptr(self.g_sdf), ptr(m), ptr(v), self.sdf.numel(), step,
                                             b1, b2, self.lr["sdf"] * lr_scale, eps, st))
            m, v = self.adam["k0"]
            if self.sparse_k0:
                # k0: masked Adam over the touched bricks only; clears the gradients / flags it consumed
                if self._brick_list_ready:          # the multi-GPU exchange has built the (ordered) list of the union
                    check(lib.fgs_adam_upd_bricks(ptr(self.k0), ptr(self.g_k0), ptr(m), ptr(v), ptr(self.ws["k0_brick_list"]),
                                                  self.X, self.Y, self.Z, K0_DIM, step, b1, b2, self.lr["k0"] * lr_scale,
                                                  eps, st))
                else:                               # single GPU: straight from the flags, no list
                    check(lib.fgs_adam_upd_brick_flags(ptr(self.k0), ptr(self.g_k0), ptr(m), ptr(v),
                                                       ptr(self.ws["k0_brick_flags"]), ptr(self.ws["k0_brick_ulist"]),
                                                       self.X, self.Y, self.Z, K0_DIM, step, b1, b2,
                                                       self.lr["k0"] * lr_scale, eps, st))
                self._brick_list_ready = False
            else:
                check(lib.fgs_adam_upd(ptr(self.k0), ptr(self.g_k0), ptr(m), ptr(v), None, self.k0.numel(), step,
                                       b1, b2, self.lr["k0"] * lr_scale, eps, 1, st))
                check(lib.fgs_memset_zero(ptr(self.g_k0), self.g_k0.numel() * 4, st))
                check(lib.fgs_memset_zero(ptr(self.ws["k0_brick_flags"]), self.ws["k0_brick_flags"].numel(), st))
            self._grads_clean = True
            if parts == "all":
                self._adam_mlp(step, lr_scale, b1, b2, eps, st)

    def _adam_mlp(self, step, lr_scale, b1, b2, eps, st):
        t = self.train_cfg
        n = 2 * MLP_LAYERS
        lrs = [self.lr["rgbnet"]] * 4 + [self.lr["refnet"]] * 4
        lrs = (C.c_float * n)(*[x * lr_scale for x in lrs + lrs])
        steps = (C.c_int * n)(*([step] * n))
        mt = self._mt
        check(lib.fgs_adam_upd_multi(n, mt["p"], mt["g"], mt["m"], mt["v"], mt["n"], steps, lrs, b1, b2, eps, st))

    def all_reduce_grads(self):
        """Ray-sharded data parallelism: sum the per-rank gradients (NCCL over NVLink), not overlapped with anything
        (train_step overlaps the first two parts with the weight-gradient kernel).
        g_sdf and the MLP bucket are dense (16 MB + 1.8 MB at 160^3).  g_k0 (197 MB, ~2 % non-zero) is exchanged
        SPARSELY: union of the touched-brick flags (MAX all-reduce of one byte per 4^3 brick), ordered brick list,
        pack.  Then ONE all-reduce over [g_sdf | MLP bucket | packed bricks] and the unpack.  The brick count
        crosses to the host once per step to size the collective."""
        if self.world_size == 1:
            return
        self._exchange_begin()
        self._exchange_pack()
        self._exchange_reduce()

    def _exchange_begin(self):
        """Part 1 (needs only the grid gradients): union of the touched bricks, ordered brick list, its count on
        the way to the host."""
        import torch.distributed as dist
        ws = self.ws
        if self.voxel_exchange and self.voxel_bits:
            # union of the touched voxels as an all-gather of bit masks (512 KB per rank at 160^3); the union's
            # touched bricks (for the brick Adam) are derived from it: no other collective before the big one
            V = self.X * self.Y * self.Z
            bits = ws["k0_vox_bits"]
            with torch.cuda.device(self.device):
                check(lib.fgs_voxel_bits_pack(ptr(ws["k0_vox_flags"]), ptr(bits[self.rank]), V, stream()))
            dist.all_gather_into_tensor(bits.view(-1), bits[self.rank])
            with torch.cuda.device(self.device):
                st = stream()
                check(lib.fgs_memset_zero(ptr(ws["k0_brick_flags"]), ws["k0_brick_flags"].numel(), st))
                check(lib.fgs_voxel_bits_union_list(ptr(bits), self.world_size, self.X, self.Y, self.Z,
                                                    ptr(ws["k0_brick_flags"]), ptr(ws["k0_vox_list"]), st))
                check(lib.fgs_brick_list(ptr(ws["k0_brick_flags"]), self.X, self.Y, self.Z, ptr(ws["k0_brick_list"]), st))
                self._count_host[:1].copy_(ws["k0_vox_list"][:1], non_blocking=True)
                self._count_evt.record()
            return
        # union of the touched bricks (k0 and, when g_sdf travels sparsely, the SDF grid) over the ranks
        dist.all_reduce(self._brick_flags2 if self.sparse_sdf else ws["k0_brick_flags"], op=dist.ReduceOp.MAX)
        if not self.sparse_k0:
            return
        if self.voxel_exchange:
            dist.all_reduce(ws["k0_vox_flags"], op=dist.ReduceOp.MAX)          # union of the touched voxels
        with torch.cuda.device(self.device):
            st = stream()
            # the brick list drives the masked Adam; the voxel list (when enabled) the pack / unpack of the exchange
            check(lib.fgs_brick_list(ptr(ws["k0_brick_flags"]), self.X, self.Y, self.Z, ptr(ws["k0_brick_list"]), st))
            if self.voxel_exchange:
                check(lib.fgs_voxel_list(ptr(ws["k0_vox_flags"]), self.X * self.Y * self.Z, ptr(ws["k0_vox_list"]), st))
                self._count_host[:1].copy_(ws["k0_vox_list"][:1], non_blocking=True)
            else:
                self._count_host[:1].copy_(ws["k0_brick_list"][:1], non_blocking=True)
            if self.sparse_sdf:
                check(lib.fgs_brick_list(ptr(ws["sdf_brick_flags"]), self.X, self.Y, self.Z, ptr(ws["sdf_brick_list"]), st))
                self._count_host[1:].copy_(ws["sdf_brick_list"][:1], non_blocking=True)
            self._count_evt.record()

    def _exchange_pack(self):
        """Part 2: the one host sync of a DP step (brick count -> size of the collective) and the pack of the
        union bricks behind the dense gradients."""
        if not self.sparse_k0:
            return
        with torch.cuda.device(self.device):
            self._count_evt.synchronize()
            n = int(self._count_host[0])                            # voxels (12 floats each) or bricks (768 floats)
            ns = int(self._count_host[1]) if self.sparse_sdf else 0 # touched SDF bricks (64 floats each)
            self.exchange_bricks, self.exchange_sdf_bricks = n, ns
            per = K0_DIM if self.voxel_exchange else 768
            self._xbuf_overflow = None
            if n * per + ns * 64 > self._compact.numel():           # rare: more than the shared buffer holds
                self._xbuf_overflow = torch.empty(n * per + ns * 64, dtype=torch.float32, device=self.device)
            buf = self._xbuf_overflow if self._xbuf_overflow is not None else self._compact
            if n > 0:
                self._copy_touched(buf, n, 1)
            if ns > 0:
                self._copy_sdf_bricks(buf[n * per:], ns, 1)
        self._brick_list_ready = True

    def _exchange_reduce(self):
        """Part 3 (needs the MLP gradients too): one all-reduce over [g_sdf | MLP bucket | packed bricks], unpack."""
        import torch.distributed as dist
        if not self.sparse_k0:
            dist.all_reduce(self._xbuf[:self._xdense])
            g = self.g_k0
            dist.all_reduce(torch.as_strided(g, (g.numel(),), (1,), g.storage_offset()))
            lib.fgs_prof_mark(1, stream())
            return
        n, ns = self.exchange_bricks, self.exchange_sdf_bricks
        per = K0_DIM if self.voxel_exchange else 768
        # with g_sdf among the packed bricks the collective starts at the MLP bucket, not at the dense grid
        lo = self._xdense - self.g_mlp_flat.numel() if self.sparse_sdf else 0
        with torch.cuda.device(self.device):
            if self._xbuf_overflow is not None:
                dist.all_reduce(self._xbuf[lo:self._xdense])
                dist.all_reduce(self._xbuf_overflow)
                buf = self._xbuf_overflow
            else:
                dist.all_reduce(self._xbuf[lo:self._xdense + n * per + ns * 64])
                buf = self._compact
            lib.fgs_prof_mark(1, stream())
            if n > 0:
                self._copy_touched(buf, n, 0)
            if ns > 0:
                self._copy_sdf_bricks(buf[n * per:], ns, 0)

    def _native_train_step(self, rays_o, rays_d, viewdirs, target, global_step, near, bg, n_rays_global, has_last_ray, tv):
        """The single-GPU iteration as ONE library call (fgs_fine_train_step): same launches, order and streams as the
        host-sequenced branch of train_step below, without ~20 Python -> C round trips per iteration (host enqueue
        time 0.18 -> 0.10 ms per iteration; device time unchanged)."""
        t = self.train_cfg
        self._bind_inputs(rays_o, rays_d, viewdirs, global_step, near, bg, False, False)
        L = self._fill_loss(target, n_rays_global, has_last_ray)
        S = self._S
        if S is None:
            S = self._S = FgsFineStep()
            n = 2 * MLP_LAYERS
            self._S_steps, self._S_lrs = (C.c_int * n)(), (C.c_float * n)()
            mt = self._mt
            S.beta1, S.beta2, S.eps = 0.9, 0.99, 1e-8
            S.sdf_m, S.sdf_v = ptr(self.adam["sdf"][0]), ptr(self.adam["sdf"][1])
            S.k0_m, S.k0_v = ptr(self.adam["k0"][0]), ptr(self.adam["k0"][1])
            S.sparse_k0 = int(self.sparse_k0)
            S.brick_list = ptr(self.ws["k0_brick_ulist"])     # single GPU: the unordered list (two counters + ids)
            S.n_mlp = n
            S.mlp_p, S.mlp_g = C.cast(mt["p"], _fp), C.cast(mt["g"], _fp)
            S.mlp_m, S.mlp_v = C.cast(mt["m"], _fp), C.cast(mt["v"], _fp)
            S.mlp_n = C.cast(mt["n"], _fp)
            S.mlp_steps, S.mlp_lrs = C.cast(self._S_steps, _fp), C.cast(self._S_lrs, _fp)
            S.side_stream = self._side.cuda_stream
        S.tv_smooth = int(tv and t["smooth_grad_tv"] > 0)
        if S.tv_smooth:
            S.smooth_weight = t["weight_tv_density"] * t["smooth_grad_tv"]
            if self.nonempty_mask is not None:
                S.smooth_n_sel, S.nonempty = 3.0 * self.n_nonempty, ptr(self.nonempty_mask)
            else:
                S.smooth_n_sel, S.nonempty = 3.0 * self.X * self.Y * self.Z, None
            S.tv_scratch = ptr(self.tv_scratch)
        S.tv_sdf = int(tv and t["sdf_tv"] > 0)
        if S.tv_sdf:
            w = t["weight_tv_density"] * t["sdf_tv"] / (n_rays_global or self.n_rays)
            S.sdf_tv_weight = float(w * max(self.X, self.Y, self.Z) / 128)
            S.sdf_tv_dense = int(global_step < t["tv_dense_before"])
        self.adam_step += 1
        S.adam_step = self.adam_step
        S.lr_sdf, S.lr_k0 = self.lr["sdf"], self.lr["k0"]
        for i in range(MLP_LAYERS):
            lr = self.lr["rgbnet"] if i < 4 else self.lr["refnet"]
            self._S_lrs[i] = self._S_lrs[MLP_LAYERS + i] = lr
            self._S_steps[i] = self._S_steps[MLP_LAYERS + i] = self.adam_step
        self.P.mlp_grads_zeroed = int(self._mlp_grads_clean)     # its Adam clears the MLP gradients it consumes
        with torch.cuda.device(self.device):
            if self.use_graph and self.adam_step > 2:          # (the first iterations set function attributes, tensor maps ...)
                # one CUDA graph per iteration, on the engine's own stream (capture is illegal on the legacy default stream)
                cur = torch.cuda.current_stream(self.device)
                if self._graph_stream is None:
                    self._graph_stream = torch.cuda.Stream(device=self.device)
                gs = self._graph_stream
                gs.wait_stream(cur)
                with torch.cuda.stream(gs):
                    used = C.c_int(0)
                    check(lib.fgs_fine_train_step_graph(C.byref(self.P), C.byref(L), C.byref(S), stream(), C.byref(used)))
                    self.graph_steps += int(used.value)
                cur.wait_stream(gs)
            else:
                check(lib.fgs_fine_train_step(C.byref(self.P), C.byref(L), C.byref(S), stream()))
        self._mlp_grads_clean = True
        self._grads_clean = True
        self._brick_list_ready = False

    def _copy_touched(self, buf, n, pack):
        """Gather (pack=1) / scatter the touched part of g_k0 to / from `buf`: listed voxels or listed bricks."""
        if self.voxel_exchange:
            check(lib.fgs_voxel_copy(ptr(self.g_k0), ptr(buf), ptr(self.ws["k0_vox_list"]), n, K0_DIM, pack, stream()))
        else:
            check(lib.fgs_brick_copy(ptr(self.g_k0), ptr(buf), ptr(self.ws["k0_brick_list"]), n, self.X, self.Y,
                                     self.Z, K0_DIM, pack, stream()))

    def _copy_sdf_bricks(self, buf, n, pack):
        """Gather (pack=1) / scatter the touched 4^3 bricks of g_sdf (64 floats each) to / from `buf`."""
        check(lib.fgs_brick_copy(ptr(self.g_sdf), ptr(buf), ptr(self.ws["sdf_brick_list"]), n, self.X, self.Y, self.Z,
                                 1, pack, stream()))

    def train_step(self, rays_o, rays_d, viewdirs, target, global_step, near=2.0, bg=1.0, n_rays_global=None,
                   has_last_ray=True, _stop_before_optimizer=False):
        """One iteration of model/nerf_training.py:300-373 (fine stage).  Asynchronous.
        Data-parallel (world_size > 1): the backward runs in two phases -- the grid gradients first; the union of
        the touched k0 bricks, its count (the one host sync) and the pack run on a side stream while the
        weight-gradient kernel (which leaves a few SMs free for them) runs on the main one; then ONE all-reduce over
        [g_sdf | MLP bucket | packed bricks]."""
        native = self.world_size == 1 and self.native_step and not _stop_before_optimizer
        if not native:
            self.forward(rays_o, rays_d, viewdirs, global_step, near, bg, _skip_composite=True)   # fgs_fine_loss composites
        self.zero_grad()                                                   # optimizer.zero_grad(set_to_none)
        t = self.train_cfg
        tv = self.is_tv_step(global_step) and t["weight_tv_density"] > 0
        self._ensure_adam()
        if tv and self.tv_scratch is None:
            self.tv_scratch = torch.empty(6 * self.X * self.Y * self.Z, dtype=torch.float32, device=self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)   # its small CTAs go first when an SM has room
            self._ev_grids, self._ev_exch = torch.cuda.Event(), torch.cuda.Event()
            check(lib.fgs_set_dw_sm_reserve(self.dw_sm_reserve if self.world_size > 1 else 0))
        main = torch.cuda.current_stream(self.device)
        if native:
            self._native_train_step(rays_o, rays_d, viewdirs, target, global_step, near, bg, n_rays_global, has_last_ray, tv)
            return
        if self.world_size == 1:
            self.loss_backward(target, n_rays_global, has_last_ray, phase=1)     # the grid gradients are final after phase 1
            if tv and t["smooth_grad_tv"] > 0:
                self.smooth_grad_tv_add_grad()
            if tv and t["sdf_tv"] > 0:
                self.sdf_tv_add_grad(global_step, n_rays_global)
            if _stop_before_optimizer:                  # tests: the gradients an optimiser step would consume
                self.backward_weights()
                return
            # one stream: the optimiser kernels of the grids first (their gradients are final and still in L2), then the
            # weight gradients (a side stream cannot run beside the persistent weight-gradient CTAs, see fine_step.cu)
            self.optimizer_step(parts="grids")
            self.backward_weights()
            self.optimizer_step(parts="mlp")
            return
        else:
            self.loss_backward(target, n_rays_global, has_last_ray, phase=1)
            if tv and t["smooth_grad_tv"] > 0 and self.rank == 0 and not self.sparse_sdf:
                self.smooth_grad_tv_add_grad()                             # a pure function of the params:
            self._ev_grids.record(main)                                    # added once, not per rank
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._ev_grids)
                lib.fgs_prof_mark(4, stream())
                self._exchange_begin()                                     # flags, brick list, count -> host
            self.backward_weights()                                        # main stream, overlaps the exchange
            with torch.cuda.stream(self._side):
                self._exchange_pack()                                      # host sync + pack, still beside phase 2
                self._ev_exch.record(self._side)
            main.wait_event(self._ev_exch)
            lib.fgs_prof_mark(5, stream())
            self._exchange_reduce()                                        # ONE all-reduce of everything, unpack
            if tv and t["smooth_grad_tv"] > 0 and self.sparse_sdf:
                # g_sdf travelled as touched bricks: the (dense) smooth-TV gradient, a pure function of the replicated
                # parameters, is added by every rank AFTER the reduce (every rank waited for rank 0's copy before anyway)
                self.smooth_grad_tv_add_grad()
        if tv and t["sdf_tv"] > 0:
            self.sdf_tv_add_grad(global_step, n_rays_global)
        if _stop_before_optimizer:
            return
        self.optimizer_step()


def smoke():
    """Tiny fused step on cuda:0 checked against the oracle's forward (called by __graft_entry__)."""
    from . import scene
    dev = torch.device("cuda:0")
    sc = scene.make_scene(G=48, mask_G=40)
    ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=192, H=100, W=100, n_views=8)]
    eng = FineStepEngine(sc, 192, dev)
    eng.set_nonempty_mask()
    eng.train_step(ro, rd, vd, tgt, 5000)
    torch.cuda.synchronize()
    cnt = eng.counters()
    assert cnt["M2"] > 0 and bool(torch.isfinite(eng.ws["rgb_marched"]).all())
    print("smoke: fused fine step ok,", cnt, "loss", eng.ws["loss"].cpu().numpy().round(6).tolist())
    return eng
