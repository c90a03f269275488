"""Coarse stage of the per-ray path: host side of the "coarse stage" section of include/fgs_b200_fine.h.

`CoarseStepEngine` is the counterpart of engine.FineStepEngine for the two stages that run nerf.forward_coarse --
'coarse' and (cfg['stage'] = 'geometry_searching') the first stage of a run, which has no MaskCache, k0_dim 6 and the
Incremental-Voxel-Training box `inc_mask`:

  forward()      nerf.forward_coarse                               model/nerf.py:943-1075
  train_step()   forward + loss (incl. weight_rgbper, ori_tv) + backward + MaskedAdam
                                                                   model/nerf_training.py:300-373, coarse_train

It shares the sampling / MaskCache / compaction / compositing / loss / tensor-core MLP kernels with the fine stage
and adds the kernels of csrc/coarse.cu plus three grid operators: the Gaussian smoother of the SDF grid that every
forward applies (`fgs_conv3d_replicate`, nerf.py:969) and its transpose, the central-difference gradient grid
(`fgs_sdf_gradient_grid`, :972) and its transpose, and the autograd SDF TV of `ori_tv` (`fgs_l1_tv_add_grad`).
Single GPU (the data-parallel exchange is built for the fine stage only)."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .engine import (CNT, K0_DIM, MLP_LAYERS, NUM_COUNTERS, WIDTH, FgsFineLoss, FgsFineParams, FineStepEngine, _fp,
                     gaussian_3dconv_weights)

COARSE_MODEL = dict(  # config/shiny_blender.py `coarse_model`
    k0_dim=12, refnet_width=192, refnet_depth=3, posbase_pe=5, viewbase_pe=1, refbase_pe=5,
    smooth_ksize=5, smooth_sigma=0.8, s_ratio=50, s_start=0.2, step_start=0, stepsize=0.5,
    fast_color_thres=1e-4, mask_cache_thres=1e-3,
)
COARSE_TRAIN = dict(  # config/shiny_blender.py `coarse_train`
    lrate_k0=0.1, lrate_sdf=0.1, lrate_refnet=1e-3, lrate_decay=20,
    weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=0.2, weight_orientation=1e-4, sigmoid_rgb_loss=0.1,
    weight_tv_density=0.01, weight_tv_k0=0.0, tv_every=1, tv_from=0, tv_end=40000, tv_dense_before=40000,
    sdf_tv=0.1, smooth_grad_tv=0.05, ori_tv=True, N_iters=15000,
    skip_zero_grad_fields=("density", "k0", "sdf"),
)
GEOMETRY_MODEL = dict(  # config/shiny_blender.py `geometry_searching_model`
    stage="geometry_searching", k0_dim=6, refnet_width=128, refnet_depth=3, posbase_pe=5, viewbase_pe=1, refbase_pe=3,
    smooth_ksize=5, smooth_sigma=0.8, s_ratio=50, s_start=0.2, step_start=0, stepsize=0.5,
    fast_color_thres=1e-4, mask_cache_thres=1e-3,
)
GEOMETRY_TRAIN = dict(COARSE_TRAIN, N_iters=12000, lrate_density=0.1)    # config `geometry_searching` (same loss weights)
LAST = MLP_LAYERS - 1          # slot of the refnet's output layer in FgsFineParams.W / B / gW / gB


class FgsCoarse(C.Structure):
    """Mirror of `struct FgsCoarse` (include/fgs_b200_fine.h)."""
    _fields_ = [
        ("sdf_smooth", _fp), ("grad_grid", _fp), ("g_smooth", _fp), ("g_grad_grid", _fp),
        ("in_dim", C.c_int32), ("in_ld", C.c_int32),
        ("pos_pe", C.c_int32), ("view_pe", C.c_int32), ("ref_pe", C.c_int32), ("k0_dim", C.c_int32),
        ("width", C.c_int32), ("pad_", C.c_int32),
    ]


_PP, _LP, _CP = C.POINTER(FgsFineParams), C.POINTER(FgsFineLoss), C.POINTER(FgsCoarse)
_i = C.c_int
_lib.declare("fgs_coarse_wsplit_floats", C.c_int64, [_i])
_lib.declare("fgs_coarse_forward", _i, [_PP, _CP, _fp])
_lib.declare("fgs_coarse_backward", _i, [_PP, _LP, _CP, _fp])
_lib.declare("fgs_sdf_gradient_grid", _i, [_fp, _fp, _i, _i, _i, C.c_float, _fp])
_lib.declare("fgs_sdf_gradient_grid_bwd", _i, [_fp, _fp, _i, _i, _i, C.c_float, _fp])
_lib.declare("fgs_conv3d_replicate_bwd", _i, [_fp, _fp, _i, _i, _i, _fp, _i, _fp])
_lib.declare("fgs_l1_tv_add_grad", _i, [_fp, _fp, _fp, _i, _i, _i, C.c_double, C.c_double, _fp, _fp])
_lib.declare("fgs_l1_tv_unmasked_add_grad", _i, [_fp, _fp, _i, _i, _i, C.c_double, _fp, _fp, _fp])


def sdf_gradient_grid(sdf, voxel_size):
    """nerf.neus_sdf_gradient(mode='interpolate') (model/nerf.py:485-494): [1,1,X,Y,Z] -> [1,3,X,Y,Z]."""
    _lib.check_input(sdf, "sdf", torch.float32)
    X, Y, Z = (int(v) for v in sdf.shape[-3:])
    G = torch.empty((1, 3, X, Y, Z), dtype=torch.float32, device=sdf.device)
    with torch.cuda.device(sdf.device):
        check(lib.fgs_sdf_gradient_grid(ptr(sdf), ptr(G), X, Y, Z, float(voxel_size), stream()))
    return G


def sdf_gradient_grid_backward(g_G, voxel_size, g_sdf=None):
    """g_sdf += (d neus_sdf_gradient / d sdf)^T g_G."""
    _lib.check_input(g_G, "g_G", torch.float32)
    X, Y, Z = (int(v) for v in g_G.shape[-3:])
    if g_sdf is None:
        g_sdf = torch.zeros((1, 1, X, Y, Z), dtype=torch.float32, device=g_G.device)
    with torch.cuda.device(g_G.device):
        check(lib.fgs_sdf_gradient_grid_bwd(ptr(g_G), ptr(g_sdf), X, Y, Z, float(voxel_size), stream()))
    return g_sdf


def smooth_conv_backward(g_out, weights, g_in=None):
    """g_in += conv^T g_out for nerf._gaussian_3dconv's Conv3d (replicate padding); weights: CPU [k,k,k]."""
    _lib.check_input(g_out, "g_out", torch.float32)
    X, Y, Z = (int(v) for v in g_out.shape[-3:])
    w = weights.float().cpu().contiguous()
    if g_in is None:
        g_in = torch.zeros_like(g_out)
    with torch.cuda.device(g_out.device):
        check(lib.fgs_conv3d_replicate_bwd(ptr(g_out), ptr(g_in), X, Y, Z, w.data_ptr(), int(w.shape[0]), stream()))
    return g_in


def l1_tv_add_grad(sdf, mask, g_sdf, weight, loss_out=None):
    """weight * total_variation(sdf, mask) (model/nerf.py:1212-1221): gradient added into g_sdf, loss into loss_out[0]."""
    X, Y, Z = (int(v) for v in sdf.shape[-3:])
    cnt = float(mask.sum().item())
    with torch.cuda.device(sdf.device):
        check(lib.fgs_l1_tv_add_grad(ptr(sdf), ptr(mask), ptr(g_sdf), X, Y, Z, float(weight), cnt, ptr(loss_out), stream()))


def make_refnet(cfg=None, seed=790):
    """A coarse-stage refnet with nn.Linear's default initialisation under a fixed seed (synthetic scenes / benchmarks):
    [(W [w,in], b [w]), (W [w,w], b [w]), (W [3,w], b [3])]."""
    c = dict(COARSE_MODEL)
    c.update(cfg or {})
    dim = (3 + 6 * c["refbase_pe"]) + c["k0_dim"] + (3 + 6 * c["posbase_pe"]) + 3 + (3 + 6 * c["viewbase_pe"])
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        b = 1.0 / math.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b).float(), ((torch.rand(o, generator=g) * 2 - 1) * b).float()

    w = c["refnet_width"]
    return [lin(dim, w), lin(w, w), lin(w, 3)]


class CoarseStepEngine(FineStepEngine):
    """State + workspace of the coarse-stage model (`nerf(stage='coarse')`) for batches of `n_rays` rays on one GPU.
    `refnet`: [(W [w,in], b [w]), (W [w,w], b [w]), (W [3,w], b [3])], in = 12 + PE dims + 3 (90 for the shipped
    configs), w <= 256.  The refnet runs on the fine stage's 256-wide tensor-core kernels, zero-padded (the pad rows
    and columns provably stay 0 under Adam: see include/fgs_b200_fine.h)."""

    def __init__(self, scene, refnet, n_rays, device="cuda", cfg=None, train_cfg=None, cap_m1_per_ray=192,
                 cap_m2_per_ray=36):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CoarseStepEngine needs a CUDA device: fgs_b200 has no CPU path")
        dev = self.device
        self.cfg = dict(COARSE_MODEL)
        self.cfg.update(cfg or {})
        self.cfg.setdefault("step_start", 0)
        self.train_cfg = dict(COARSE_TRAIN)
        self.train_cfg.update(train_cfg or {})
        c = self.cfg
        self.stage = c.get("stage", "coarse")
        if self.stage not in ("coarse", "geometry_searching"):
            raise NotImplementedError(f"CoarseStepEngine: stage {self.stage!r} (forward_coarse stages: coarse, geometry_searching)")
        if not (0 < c["k0_dim"] <= K0_DIM):
            raise NotImplementedError(f"coarse stages: k0_dim must be 1..{K0_DIM} (got {c['k0_dim']})")
        if c["refnet_depth"] != 3 or len(refnet) != 3:
            raise NotImplementedError("coarse stage: refnet_depth must be 3 (in -> width -> width -> 3)")
        self.n_rays = int(n_rays)
        self.world_size, self.rank = 1, 0
        self.voxel_exchange = False
        f32, i32 = torch.float32, torch.int32
        self.lr = {k: float(self.train_cfg["lrate_" + k]) for k in ("sdf", "k0", "refnet")}
        # ---- parameters ----
        self.xyz_min, self.xyz_max = scene.xyz_min.clone().float(), scene.xyz_max.clone().float()
        self.mask_xyz_min, self.mask_xyz_max = scene.mask_xyz_min.clone().float(), scene.mask_xyz_max.clone().float()
        self.voxel_size = np.float32(scene.voxel_size)
        self.sdf = scene.sdf.to(dev, f32).contiguous().clone()
        # the feature grid is stored with 12 channels-last channels whatever k0_dim is (the gather / scatter / brick-Adam
        # kernels move three float4 per voxel); channels beyond k0_dim are 0 and stay 0 (their gradient is exactly 0, the
        # masked Adam skips them): self.k0 / self.g_k0 are the [1, k0_dim, X, Y, Z] views the reference's names refer to
        kd = int(c["k0_dim"])
        if tuple(scene.k0.shape[1:2]) != (kd,):
            raise RuntimeError(f"k0 grid has {scene.k0.shape[1]} channels, config says k0_dim {kd}")
        self._k0_store = torch.zeros((1, K0_DIM) + tuple(scene.k0.shape[2:]), dtype=f32, device=dev).contiguous(
            memory_format=torch.channels_last_3d)
        self._k0_store[:, :kd] = scene.k0.to(dev, f32)
        self.k0 = self._k0_store if kd == K0_DIM else self._k0_store[:, :kd]
        # geometry_searching has no MaskCache (model/nerf.py:158-161): every in-bbox sample passes
        self.mask_grid = scene.mask_grid.to(dev, f32).contiguous() if self.stage == "coarse" else None
        self.mask_thres = np.float32(c["mask_cache_thres"])
        self.fast_color_thres = np.float32(c["fast_color_thres"])
        X, Y, Z = self.sdf.shape[2:]
        self.X, self.Y, self.Z = X, Y, Z
        self.num_voxels = self.num_voxels_base = X * Y * Z
        self.width = int(c["refnet_width"])
        self.in_dim = kd + (3 + 6 * c["posbase_pe"]) + (3 + 6 * c["refbase_pe"]) + 3 + (3 + 6 * c["viewbase_pe"])
        self.in_ld = (self.in_dim + 31) // 32 * 32
        w = self.width
        if not (0 < w <= WIDTH) or self.in_ld > 256:
            raise NotImplementedError(f"coarse stage: refnet_width {w} / input {self.in_dim} beyond the 256-wide kernels")
        dims = [(w, self.in_dim), (w, w), (3, w)]
        pads = [(WIDTH, self.in_ld), (WIDTH, WIDTH), (3, WIDTH)]
        self.Wt, self.weights, self.Bt, self.biases = [], [], [], []
        for (wt, b), d, pd in zip(refnet, dims, pads):
            if tuple(wt.shape) != d:
                raise RuntimeError(f"refnet layer shape {tuple(wt.shape)}, expected {d}")
            st = torch.zeros(pd, dtype=f32, device=dev)
            st[:d[0], :d[1]] = wt.to(dev, f32)
            bt = torch.zeros(pd[0], dtype=f32, device=dev)
            bt[:d[0]] = b.to(dev, f32)
            self.Wt.append(st); self.weights.append(st[:d[0], :d[1]])
            self.Bt.append(bt); self.biases.append(bt[:d[0]])
        # ---- gradients ----
        self.g_sdf = torch.zeros_like(self.sdf)
        self._g_k0_store = torch.zeros_like(self._k0_store)
        self.g_k0 = self._g_k0_store if kd == K0_DIM else self._g_k0_store[:, :kd]
        self.inc_mask = None
        self._tv_sums = torch.zeros(2, dtype=torch.float64, device=dev)
        sizes = [t.numel() for t in self.Wt] + [t.numel() for t in self.Bt]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        self.g_mlp_flat = torch.zeros(offs[-1], dtype=f32, device=dev)
        views = [self.g_mlp_flat[o:o + n] for o, n in zip(offs, sizes)]
        self.gW = [v.view_as(t) for v, t in zip(views[:3], self.Wt)]
        self.gB = [v.view_as(t) for v, t in zip(views[3:], self.Bt)]
        self.weight_grads = [g[:d[0], :d[1]] for g, d in zip(self.gW, dims)]
        self.bias_grads = [g[:d[0]] for g, d in zip(self.gB, dims)]
        self.adam, self.adam_step = {}, 0
        # ---- the two derived grids of a forward and their gradients ----
        V = X * Y * Z
        self.sdf_smooth = torch.empty_like(self.sdf)
        self.grad_grid = torch.empty((1, 3, X, Y, Z), dtype=f32, device=dev)
        self._g_grids = torch.zeros(4 * V, dtype=f32, device=dev)            # [g_smooth | g_grad_grid], one memset
        self.g_smooth = self._g_grids[:V].view_as(self.sdf)
        self.g_grad_grid = self._g_grids[V:].view(1, 3, X, Y, Z)
        self.smooth_ksize = int(c.get("smooth_ksize", 0))
        self.smooth_weights = gaussian_3dconv_weights(self.smooth_ksize, c["smooth_sigma"], dev) \
            if self.smooth_ksize > 0 else None
        # ---- workspace ----
        N = self.n_rays
        diag = float((self.xyz_max - self.xyz_min).norm())
        stepdist = float(np.float32(c["stepsize"]) * self.voxel_size)
        self.cap_words = int(math.ceil((diag / stepdist + 2) / 32)) + 1
        self.cap_m1 = int(cap_m1_per_ray * N)
        self.cap_m2 = int((cap_m2_per_ray * N + 127) // 128 * 128)
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
        ws["X0"], ws["dX0"] = e(self.cap_m2, self.in_ld), e(self.cap_m2, self.in_ld)
        ws["H"] = torch.zeros(2, self.cap_m2, WIDTH, dtype=f32, device=dev)    # (columns beyond the GEMMs' N are never written)
        ws["logit"], ws["normal2"], ws["dlogit"] = e(self.cap_m2, 4), e(self.cap_m2, 4), e(self.cap_m2, 4)
        ws["dZ"] = torch.zeros(2, self.cap_m2, WIDTH, dtype=f32, device=dev)      # bf16 hi / mid planes, zero-initialised
        ws["dlogit_p"] = torch.zeros(2, self.cap_m2, 8, dtype=torch.int16, device=dev)
        for k in ("rgb_marched", "sigmoid_rgb", "normal_marched"):
            ws[k] = e(N, 3)
        ws["w_split"] = torch.zeros(int(lib.fgs_coarse_wsplit_floats(self.in_ld)), dtype=f32, device=dev)
        ws["hbits"] = e(2, self.cap_m2, 8, dt=i32)
        nb = ((X + 3) // 4) * ((Y + 3) // 4) * ((Z + 3) // 4)
        ws["k0_brick_flags"] = torch.zeros(nb, dtype=torch.uint8, device=dev)
        ws["k0_brick_ulist"] = torch.zeros(int(lib.fgs_brick_list_ints(X, Y, Z)) + 1, dtype=i32, device=dev)
        self.sparse_k0 = (Z % 4 == 0)
        ws["counters"] = torch.zeros(NUM_COUNTERS, dtype=i32, device=dev)
        ws["loss"] = torch.zeros(7, dtype=torch.float64, device=dev)   # mse, entropy, orientation, sigmoid_rgb, smooth_tv, sdf_tv, rgbper
        self._grads_clean = True
        self._mlp_grads_clean = True
        self._brick_list_ready = False
        self.tv_scratch = None
        self.nonempty_mask = None
        self._keep = None
        self.P, self.L, self.Cc = FgsFineParams(), FgsFineLoss(), FgsCoarse()
        self._fill_static()

    # -------------------------------------------------------------------------------------
    def _fill_static(self):
        P, ws, c = self.P, self.ws, self.cfg
        P.n_rays = self.n_rays
        P.X, P.Y, P.Z = self.X, self.Y, self.Z
        P.mX, P.mY, P.mZ = self.mask_grid.shape[2:] if self.mask_grid is not None else (0, 0, 0)
        P.cap_words, P.cap_m1, P.cap_m2 = self.cap_words, self.cap_m1, self.cap_m2
        for k in range(3):
            P.xyz_min[k] = float(self.xyz_min[k]); P.xyz_max[k] = float(self.xyz_max[k])
            P.mask_xyz_min[k] = float(self.mask_xyz_min[k]); P.mask_xyz_max[k] = float(self.mask_xyz_max[k])
        sd = np.float32(c["stepsize"]) * self.voxel_size
        P.stepdist = float(sd); P.dist = float(sd)
        P.voxel_size = float(self.voxel_size)
        P.mask_thres = float(self.mask_thres)
        P.fast_color_thres = float(self.fast_color_thres)
        P.far_ = 1e9
        P.sdf, P.k0, P.mask_grid = ptr(self.sdf), ptr(self._k0_store), ptr(self.mask_grid)
        for slot, i in ((0, 0), (1, 1), (LAST, 2)):
            P.W[slot], P.B[slot] = ptr(self.Wt[i]), ptr(self.Bt[i])
            P.gW[slot], P.gB[slot] = ptr(self.gW[i]), ptr(self.gB[i])
        P.g_sdf, P.g_k0 = ptr(self.g_sdf), ptr(self._g_k0_store)
        P.g_mlp_flat, P.g_mlp_floats = ptr(self.g_mlp_flat), self.g_mlp_flat.numel()
        for k in ("t_min", "t_max", "n_steps", "cnt1", "cnt2", "off1", "off2", "bits", "alphainv_last", "ray_id1",
                  "step_id1", "flag1", "sdf1", "grad1", "alpha1", "w1", "T1", "d_sdf1", "d_grad1", "d_w1", "idx1_of2",
                  "ray_id2", "X0", "H", "logit", "normal2", "rgb_marched", "sigmoid_rgb", "cum_weights", "depth",
                  "normal_marched", "counters", "dX0", "dlogit", "w_split", "hbits", "k0_brick_flags", "dZ", "dlogit_p"):
            setattr(P, k, ptr(ws[k]))
        if self.mask_grid is not None:
            self._build_mask_occupancy()
        Cc = self.Cc
        Cc.k0_dim = int(c["k0_dim"])
        Cc.width = int(self.width)
        # without a smoother (smooth_ksize 0) the SDF is sampled from the raw grid and its gradient lands in g_sdf directly
        Cc.sdf_smooth = ptr(self.sdf_smooth if self.smooth_ksize > 0 else self.sdf)
        Cc.g_smooth = ptr(self.g_smooth if self.smooth_ksize > 0 else self.g_sdf)
        Cc.grad_grid, Cc.g_grad_grid = ptr(self.grad_grid), ptr(self.g_grad_grid)
        Cc.in_dim, Cc.in_ld = self.in_dim, self.in_ld
        Cc.pos_pe, Cc.view_pe, Cc.ref_pe = c["posbase_pe"], c["viewbase_pe"], c["refbase_pe"]

    def init_smooth_conv(self, ksize=3, sigma=1):
        """nerf.init_smooth_conv (model/nerf.py:274-278; the `smooth_updates` schedule): a new Gaussian for the SDF
        smoothing of every forward.  ksize 0 switches it off."""
        had = self.smooth_ksize > 0
        self.smooth_ksize = int(ksize)
        self.cfg["smooth_ksize"], self.cfg["smooth_sigma"] = int(ksize), sigma
        self.smooth_weights = gaussian_3dconv_weights(self.smooth_ksize, sigma, self.device) if ksize > 0 else None
        if had != (ksize > 0):
            self.Cc.sdf_smooth = ptr(self.sdf_smooth if ksize > 0 else self.sdf)
            self.Cc.g_smooth = ptr(self.g_smooth if ksize > 0 else self.g_sdf)

    def set_nonempty_mask(self):
        """nerf._set_nonempty_mask (model/nerf.py:339-353): in the coarse stage the SDF outside the mask becomes 1.
        geometry_searching has neither a MaskCache nor a non-empty mask (:158-161)."""
        if self.stage != "coarse":
            self.nonempty_mask = None
            return None
        m = super().set_nonempty_mask()
        with torch.no_grad():
            self.sdf[0, 0][~m] = 1
        return m

    @torch.no_grad()
    def set_inc_mask(self, lower, upper):
        """nerf.set_inc_mask (model/nerf.py:1078-1088) + MaskGrid (grid.py:253-272): Incremental Voxel Training keeps only
        the samples whose nearest voxel lies in the box [lower, upper] of the unit-cube lattice.  The three 1-D lattices
        are torch.linspace(0, 1, n) evaluated on the CPU like the reference's (bit-identical comparisons), the box is
        assembled on the device."""
        if getattr(self, "_inc_lin", None) is None:
            self._inc_lin = [torch.linspace(0, 1, n).to(self.device) for n in (self.X, self.Y, self.Z)]
            ext = (self.xyz_max - self.xyz_min).float()
            scale = (torch.Tensor([self.X, self.Y, self.Z]) - 1) / ext
            shift = -self.xyz_min.float() * scale
            for k in range(3):
                self.P.inc_scale[k], self.P.inc_shift[k] = float(scale[k]), float(shift[k])
            self.P.iX, self.P.iY, self.P.iZ = self.X, self.Y, self.Z
        lo = torch.as_tensor(lower, dtype=torch.float32)
        hi = torch.as_tensor(upper, dtype=torch.float32)
        ax = [(l >= float(lo[k])) & (l <= float(hi[k])) for k, l in enumerate(self._inc_lin)]
        self.inc_mask = (ax[0][:, None, None] & ax[1][None, :, None] & ax[2][None, None, :]).to(torch.uint8).contiguous()
        self.P.inc_mask = ptr(self.inc_mask)

    def unset_inc_mask(self):
        self.inc_mask = None
        self.P.inc_mask = None

    @torch.no_grad()
    def reset_refnet(self):
        """nerf.reset_voxel_and_mlp (model/nerf.py:384-396; `reset_iter`): a freshly initialised refnet (nn.Linear's
        default init from torch's global RNG, like the reference's constructor); the optimiser is the caller's business."""
        import torch.nn as nn
        w = self.width
        for i, (n_in, n_out) in enumerate(((self.in_dim, w), (w, w), (w, 3))):
            lin = nn.Linear(n_in, n_out)
            self.Wt[i].zero_(); self.Bt[i].zero_()
            self.weights[i].copy_(lin.weight.data.to(self.device))
            self.biases[i].copy_(lin.bias.data.to(self.device))

    @torch.no_grad()
    def maskout_near_cam_vox(self, cam_o, near):
        """nerf.maskout_near_cam_vox (model/nerf.py:355-367): sdf = 5 at the voxels closer than `near` to any camera."""
        g = torch.stack(torch.meshgrid(
            torch.linspace(float(self.xyz_min[0]), float(self.xyz_max[0]), self.X),
            torch.linspace(float(self.xyz_min[1]), float(self.xyz_max[1]), self.Y),
            torch.linspace(float(self.xyz_min[2]), float(self.xyz_max[2]), self.Z), indexing="ij"), -1).to(self.device)
        nearest = torch.stack([(g.unsqueeze(-2) - co).pow(2).sum(-1).sqrt().amin(-1)
                               for co in cam_o.to(self.device).split(100)]).amin(0)
        self.sdf[0, 0][nearest <= near] = 5

    # -------------------------------------------------------------------------------------
    def forward(self, rays_o, rays_d, viewdirs, global_step=20000, near=2.0, bg=1.0, render_depth=False,
                render_grad=False, _skip_composite=False):
        """nerf.forward_coarse: enqueue only.  Results live in self.ws (see outputs())."""
        self._bind_inputs(rays_o, rays_d, viewdirs, global_step, near, bg, render_depth, render_grad)
        self.P.skip_composite = int(bool(_skip_composite))
        with torch.cuda.device(self.device):
            st = stream()
            if self.smooth_ksize > 0:                                  # sdf_grid = self.smooth_conv(self.sdf.grid)   (:969)
                check(lib.fgs_conv3d_replicate(ptr(self.sdf), ptr(self.sdf_smooth), self.X, self.Y, self.Z,
                                               self.smooth_weights.data_ptr(), self.smooth_ksize, 1.0, 1.0, st))
            check(lib.fgs_sdf_gradient_grid(ptr(self.sdf), ptr(self.grad_grid), self.X, self.Y, self.Z,
                                            float(self.voxel_size), st))   # self.gradient = neus_sdf_gradient()  (:972)
            check(lib.fgs_coarse_forward(C.byref(self.P), C.byref(self.Cc), st))

    def zero_grad(self):
        if not self._grads_clean:
            with torch.cuda.device(self.device):
                st = stream()
                check(lib.fgs_memset_zero(ptr(self.g_sdf), self.g_sdf.numel() * 4, st))
                check(lib.fgs_memset_zero(ptr(self._g_k0_store), self._g_k0_store.numel() * 4, st))
                check(lib.fgs_memset_zero(ptr(self.ws["k0_brick_flags"]), self.ws["k0_brick_flags"].numel(), st))
            self._grads_clean = True

    def _fill_loss(self, target, n_rays_global=None, has_last_ray=True):
        L = super()._fill_loss(target, n_rays_global, has_last_ray)
        L.weight_rgbper = float(self.train_cfg.get("weight_rgbper", 0.0))
        L.rgbper_slot = 6
        return L

    def loss_backward(self, target):
        """Loss block + backward (model/nerf_training.py:307-347): refnet gradients overwritten, g_sdf / g_k0
        accumulated into (call zero_grad() first)."""
        L = self._fill_loss(target)
        self.ws["loss"].zero_()
        self._grads_clean = False
        self._mlp_grads_clean = False
        self.P.mlp_grads_zeroed = 0
        with torch.cuda.device(self.device):
            st = stream()
            check(lib.fgs_memset_zero(ptr(self._g_grids), self._g_grids.numel() * 4, st))
            check(lib.fgs_fine_loss(C.byref(self.P), C.byref(L), st))                 # compositing + loss + first hop
            check(lib.fgs_coarse_backward(C.byref(self.P), C.byref(L), C.byref(self.Cc), st))
            # back through the two grids of the forward into sdf.grid.grad
            if self.smooth_ksize > 0:
                check(lib.fgs_conv3d_replicate_bwd(ptr(self.g_smooth), ptr(self.g_sdf), self.X, self.Y, self.Z,
                                                   self.smooth_weights.data_ptr(), self.smooth_ksize, st))
            check(lib.fgs_sdf_gradient_grid_bwd(ptr(self.g_grad_grid), ptr(self.g_sdf), self.X, self.Y, self.Z,
                                                float(self.voxel_size), st))

    def sdf_l1_tv_add_grad(self):
        """ori_tv: weight_tv_density * density_total_variation(sdf_tv) (model/nerf_training.py:337-338, nerf.py:430-437)."""
        t = self.train_cfg
        w = t["weight_tv_density"] * t["sdf_tv"] / 2.0 / float(self.voxel_size)
        if self.nonempty_mask is None:
            if self.stage == "coarse":
                raise RuntimeError("the SDF TV of the coarse stage needs the non-empty mask (set_nonempty_mask)")
            # geometry_searching: total_variation(sdf) without a mask divides by sdf.sum() (model/nerf.py:1221)
            with torch.cuda.device(self.device):
                check(lib.fgs_l1_tv_unmasked_add_grad(ptr(self.sdf), ptr(self.g_sdf), self.X, self.Y, self.Z, float(w),
                                                      ptr(self._tv_sums), self.ws["loss"].data_ptr() + 5 * 8, stream()))
            return
        with torch.cuda.device(self.device):
            check(lib.fgs_l1_tv_add_grad(ptr(self.sdf), ptr(self.nonempty_mask), ptr(self.g_sdf), self.X, self.Y, self.Z,
                                         float(w), float(self.n_nonempty), self.ws["loss"].data_ptr() + 5 * 8, stream()))

    def reset_optimizer(self, global_step=0):
        t = self.train_cfg
        decay = 0.1 ** (global_step / (t.get("lrate_decay", 20) * 1000))
        self.lr = {k: float(t["lrate_" + k]) * decay for k in ("sdf", "k0", "refnet")}
        self.adam, self.adam_step = {}, 0
        self.ws["k0_brick_ulist"][:2].zero_()

    def _ensure_adam(self):
        if self.adam:
            return
        z = torch.zeros_like
        self.adam = {"sdf": (z(self.sdf), z(self.sdf)), "k0": (z(self._k0_store), z(self._k0_store)),
                     "W": [(z(w), z(w)) for w in self.Wt], "B": [(z(b), z(b)) for b in self.Bt]}
        n = 6
        self._mt = dict(
            p=(_fp * n)(*[ptr(x) for x in self.Wt + self.Bt]), g=(_fp * n)(*[ptr(x) for x in self.gW + self.gB]),
            m=(_fp * n)(*[ptr(x[0]) for x in self.adam["W"] + self.adam["B"]]),
            v=(_fp * n)(*[ptr(x[1]) for x in self.adam["W"] + self.adam["B"]]),
            n=(C.c_int64 * n)(*[x.numel() for x in self.Wt + self.Bt]))

    def optimizer_step(self, lr_scale=1.0):
        """MaskedAdam.step (model/adam.py:190-221) with the coarse-stage param groups: sdf and k0 masked
        (skip_zero_grad_fields, config coarse_train), the refnet dense."""
        t = self.train_cfg
        self.adam_step += 1
        step = self.adam_step
        self._ensure_adam()
        b1, b2, eps = 0.9, 0.99, 1e-8
        skip = tuple(t.get("skip_zero_grad_fields", ()))
        with torch.cuda.device(self.device):
            st = stream()
            m, v = self.adam["sdf"]
            check(lib.fgs_adam_upd(ptr(self.sdf), ptr(self.g_sdf), ptr(m), ptr(v), None, self.sdf.numel(), step, b1, b2,
                                   self.lr["sdf"] * lr_scale, eps, 1 if "sdf" in skip else 0, st))
            check(lib.fgs_memset_zero(ptr(self.g_sdf), self.g_sdf.numel() * 4, st))
            m, v = self.adam["k0"]
            if self.sparse_k0 and "k0" in skip:
                check(lib.fgs_adam_upd_brick_flags(ptr(self._k0_store), ptr(self._g_k0_store), ptr(m), ptr(v), ptr(self.ws["k0_brick_flags"]),
                                                   ptr(self.ws["k0_brick_ulist"]), self.X, self.Y, self.Z, K0_DIM, step, b1, b2,
                                                   self.lr["k0"] * lr_scale, eps, st))
            else:
                check(lib.fgs_adam_upd(ptr(self._k0_store), ptr(self._g_k0_store), ptr(m), ptr(v), None, self._k0_store.numel(),
                                       step, b1, b2, self.lr["k0"] * lr_scale, eps, 1, st))   # (masked: the pad channels must stay 0)
                check(lib.fgs_memset_zero(ptr(self._g_k0_store), self._g_k0_store.numel() * 4, st))
                check(lib.fgs_memset_zero(ptr(self.ws["k0_brick_flags"]), self.ws["k0_brick_flags"].numel(), st))
            self._grads_clean = True
            n = 6
            lrs = (C.c_float * n)(*([self.lr["refnet"] * lr_scale] * n))
            steps = (C.c_int * n)(*([step] * n))
            mt = self._mt
            check(lib.fgs_adam_upd_multi(n, mt["p"], mt["g"], mt["m"], mt["v"], mt["n"], steps, lrs, b1, b2, eps, st))

    def train_step(self, rays_o, rays_d, viewdirs, target, global_step, near=2.0, bg=1.0, n_rays_global=None,
                   has_last_ray=True, _stop_before_optimizer=False):
        """One iteration of model/nerf_training.py:300-373 with a coarse_train-style config.  Asynchronous."""
        if (n_rays_global or self.n_rays) != self.n_rays or not has_last_ray:
            raise NotImplementedError("the coarse stage runs on one GPU (no ray sharding)")
        t = self.train_cfg
        self.forward(rays_o, rays_d, viewdirs, global_step, near, bg, _skip_composite=True)
        self.zero_grad()
        self._ensure_adam()
        self.loss_backward(target)
        if self.is_tv_step(global_step) and t["weight_tv_density"] > 0:
            if t["smooth_grad_tv"] > 0:
                self.smooth_grad_tv_add_grad()
            if t.get("ori_tv", False):
                if t["sdf_tv"] > 0:
                    self.sdf_l1_tv_add_grad()
            elif t["sdf_tv"] > 0:
                self.sdf_tv_add_grad(global_step)
        if _stop_before_optimizer:
            return
        self.optimizer_step()

    # the data-parallel exchange and the library-sequenced step exist for the fine stage only
    def all_reduce_grads(self):
        raise NotImplementedError("the coarse stage runs on one GPU")

    def scale_volume_grid(self, num_voxels):
        """nerf.scale_volume_grid (model/nerf.py:369-381): a NEW engine at the new resolution (MLP values shared)."""
        from .engine import grid_resize
        from .scene import Scene
        ext = (self.xyz_max - self.xyz_min).float()
        voxel_size = (ext.prod() / num_voxels).pow(1 / 3)
        world_size = (ext / voxel_size).long().tolist()
        sc = Scene(G=int(world_size[0]), xyz_min=self.xyz_min, xyz_max=self.xyz_max, voxel_size=float(voxel_size),
                   sdf=grid_resize(self.sdf, world_size), k0=grid_resize(self._k0_store, world_size)[:, :int(self.cfg["k0_dim"])],
                   mask_grid=self.mask_grid if self.mask_grid is not None else torch.zeros(1, 1, 2, 2, 2),
                   mask_xyz_min=self.mask_xyz_min, mask_xyz_max=self.mask_xyz_max, mask_thres=float(self.mask_thres))
        refnet = [(w.clone(), b.clone()) for w, b in zip(self.weights, self.biases)]
        had_mask = self.nonempty_mask is not None
        new = CoarseStepEngine(sc, refnet, self.n_rays, self.device, cfg=self.cfg, train_cfg=self.train_cfg,
                               cap_m1_per_ray=self.cap_m1 // self.n_rays, cap_m2_per_ray=self.cap_m2 // self.n_rays)
        new.num_voxels, new.num_voxels_base = num_voxels, self.num_voxels_base
        if had_mask:
            new.set_nonempty_mask()
        return new
