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


class _GridView:
    """`model.sdf.grid` / `model.k0.grid` of the reference: the engine owns the storage, this exposes it by name."""

    def __init__(self, get):
        self._get = get

    @property
    def grid(self):
        return self._get()


class nerf(nn.Module):
    """The reference's scene model class (model/nerf.py:20-1075) on top of FineStepEngine (stage='fine') or
    coarse.CoarseStepEngine (stage='coarse' | 'geometry_searching': forward_coarse, :943-1075; k0_dim 1..12, refnet
    in -> width -> width -> 3 with width <= 256, any PE counts, Gaussian SDF smoothing; the first stage also without a
    MaskCache and with the Incremental-Voxel-Training box, set_inc_mask / unset_inc_mask): same
    constructor keywords, `forward(rays_o, rays_d, viewdirs, global_step=..., **render_kwargs) -> dict` with the
    reference's keys (:922-939), `sdf.grid` / `k0.grid`, `state_dict()` in the reference's layout, `scale_volume_grid`,
    `extract_geometry`'s field.  What differs, stated loudly:
      * forward is inference only (no autograd graph): training is the fused `train_step` / training.FineTrainer
        (the reference's loss.backward() has no counterpart to hook into -- the backward is fused with the loss);
      * stage='fine' accepts only the architecture its kernels are specialised for (k0_dim 12, 4 x 256 rgbnet and
        refnet, PE 5 / 3 / 8, grad_feat = sdf_feat = (0.5, 1, 1.5, 2), ref, use_viewdir, no SDF smoothing) -- the
        `fine_model` block of every shipped config; anything else raises NotImplementedError;
      * a batch may have at most `n_rays` rays (the workspace is allocated once); shorter batches are padded."""

    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, nearest=False, mask_cache_path=None,
                 mask_cache_thres=1e-5, fast_color_thres=0, k0_dim=12, rgbnet_depth=4, rgbnet_width=256, ref=False,
                 refnet_width=256, refnet_depth=4, posbase_pe=5, viewbase_pe=3, refbase_pe=8, grad_feat=(), sdf_feat=(),
                 s_ratio=2000, s_start=0.05, step_start=0, smooth_ksize=0, smooth_sigma=1, stage='', use_viewdir=True, stepsize=0.5,
                 n_rays=8192, device="cuda", mask_cache=None, train_cfg=None, **kwargs):
        super().__init__()
        import numpy as np
        from .engine import FineStepEngine
        from .scene import Scene
        if stage in ("coarse", "geometry_searching"):
            self._init_coarse(stage, xyz_min, xyz_max, num_voxels, num_voxels_base, nearest, mask_cache_path, mask_cache_thres,
                              fast_color_thres, k0_dim, ref, refnet_width, refnet_depth, posbase_pe, viewbase_pe, refbase_pe,
                              s_ratio, s_start, step_start, smooth_ksize, smooth_sigma, use_viewdir, stepsize, n_rays, device,
                              mask_cache, train_cfg)
            return
        want = dict(stage="fine", ref=True, use_viewdir=True, nearest=False, k0_dim=12, rgbnet_depth=4, rgbnet_width=256,
                    refnet_depth=4, refnet_width=256, posbase_pe=5, viewbase_pe=3, refbase_pe=8,
                    grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0), smooth_ksize=0)
        got = dict(stage=stage, ref=ref, use_viewdir=use_viewdir, nearest=nearest, k0_dim=k0_dim, rgbnet_depth=rgbnet_depth,
                   rgbnet_width=rgbnet_width, refnet_depth=refnet_depth, refnet_width=refnet_width, posbase_pe=posbase_pe,
                   viewbase_pe=viewbase_pe, refbase_pe=refbase_pe, grad_feat=tuple(float(v) for v in grad_feat),
                   sdf_feat=tuple(float(v) for v in sdf_feat), smooth_ksize=smooth_ksize)
        bad = {k: got[k] for k in want if got[k] != want[k]}
        if bad:
            raise NotImplementedError(f"fgs_b200.nerf is the fine-stage model of config/*.py `fine_model`; unsupported: {bad}")
        self.stage, self.ref, self.use_viewdir = stage, ref, use_viewdir
        xyz_min = torch.as_tensor(np.asarray(xyz_min), dtype=torch.float32)
        xyz_max = torch.as_tensor(np.asarray(xyz_max), dtype=torch.float32)
        # _set_grid_resolution (model/nerf.py:298-303)
        voxel_size = ((xyz_max - xyz_min).prod() / num_voxels).pow(1 / 3)
        world_size = ((xyz_max - xyz_min) / voxel_size).long().tolist()
        # ball init (:77-82), k0 zeros (grid.py:40), nn.Linear default init in the reference's construction order
        # (refnet, then rgbnet: :125-142), so torch.manual_seed gives the reference's initial weights
        x, y, z = np.mgrid[-1.0:1.0:world_size[0] * 1j, -1.0:1.0:world_size[1] * 1j, -1.0:1.0:world_size[2] * 1j]
        sdf = torch.from_numpy((x ** 2 + y ** 2 + z ** 2) ** 0.5 - 1).float()[None, None, ...]
        k0 = torch.zeros([1, k0_dim, *world_size])
        n_feat = len(want["grad_feat"])
        rgb_dim = (3 + 3 * posbase_pe * 2) + k0_dim + 3 + n_feat * 3 + n_feat * 6 + 1 + (3 + 3 * viewbase_pe * 2)
        ref_dim = (3 + 3 * refbase_pe * 2) + refnet_width

        def mlp(i, width, depth, o):
            ls = [nn.Linear(i, width)] + [nn.Linear(width, width) for _ in range(depth - 2)] + [nn.Linear(width, o)]
            return [(l.weight.data.clone(), l.bias.data.clone()) for l in ls]

        refnet = mlp(ref_dim, refnet_width, refnet_depth, 3)
        rgbnet = mlp(rgb_dim, rgbnet_width, rgbnet_depth, rgbnet_width)
        if mask_cache is None:
            if mask_cache_path is None:
                raise ValueError("the fine stage needs the MaskCache of the earlier stage (mask_cache_path or mask_cache)")
            mask_cache = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres)
        self.mask_cache = mask_cache
        cfg = dict(k0_dim=k0_dim, rgbnet_width=rgbnet_width, rgbnet_depth=rgbnet_depth, refnet_width=refnet_width,
                   refnet_depth=refnet_depth, posbase_pe=posbase_pe, viewbase_pe=viewbase_pe, refbase_pe=refbase_pe,
                   grad_feat=want["grad_feat"], sdf_feat=want["sdf_feat"], s_ratio=s_ratio, s_start=s_start,
                   step_start=step_start, stepsize=stepsize, fast_color_thres=fast_color_thres,
                   mask_cache_thres=mask_cache.mask_cache_thres)
        sc = Scene(G=int(world_size[0]), xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=float(voxel_size), sdf=sdf, k0=k0,
                   mask_grid=mask_cache.sdf_mask, mask_xyz_min=mask_cache.xyz_min.cpu(), mask_xyz_max=mask_cache.xyz_max.cpu(),
                   mask_thres=mask_cache.mask_cache_thres, rgbnet=rgbnet, refnet=refnet, cfg=cfg)
        self.engine = FineStepEngine(sc, n_rays, device, train_cfg=train_cfg)
        self.engine.num_voxels = num_voxels
        self.num_voxels, self.num_voxels_base = num_voxels, num_voxels_base or num_voxels
        self.engine.set_nonempty_mask()                       # nerf.__init__ -> _set_nonempty_mask (:176-177)
        self.sdf = _GridView(lambda: self.engine.sdf)
        self.k0 = _GridView(lambda: self.engine.k0)

    def _init_coarse(self, stage, xyz_min, xyz_max, num_voxels, num_voxels_base, nearest, mask_cache_path, mask_cache_thres,
                     fast_color_thres, k0_dim, ref, refnet_width, refnet_depth, posbase_pe, viewbase_pe, refbase_pe, s_ratio,
                     s_start, step_start, smooth_ksize, smooth_sigma, use_viewdir, stepsize, n_rays, device, mask_cache,
                     train_cfg):
        """stage='coarse' (model/nerf.py:47-48, 113-133, 171-172): no rgbnet, refnet on [k0 | xyz_emb | reflect_emb | normal |
        viewdirs_emb], SDF sampled from the Gaussian-smoothed grid, sdf = 1 outside the non-empty mask."""
        import numpy as np
        from .coarse import CoarseStepEngine
        from .scene import Scene
        bad = {}
        if not (ref and use_viewdir) or nearest:
            bad.update(ref=ref, use_viewdir=use_viewdir, nearest=nearest)
        if not (0 < k0_dim <= 12) or refnet_depth != 3 or refnet_width > 256:
            bad.update(k0_dim=k0_dim, refnet_depth=refnet_depth, refnet_width=refnet_width)
        if bad:
            raise NotImplementedError(f"fgs_b200.nerf(stage={stage!r}): unsupported {bad}")
        self.stage, self.ref, self.use_viewdir = stage, ref, use_viewdir
        xyz_min = torch.as_tensor(np.asarray(xyz_min), dtype=torch.float32)
        xyz_max = torch.as_tensor(np.asarray(xyz_max), dtype=torch.float32)
        voxel_size = ((xyz_max - xyz_min).prod() / num_voxels).pow(1 / 3)
        world_size = ((xyz_max - xyz_min) / voxel_size).long().tolist()
        x, y, z = np.mgrid[-1.0:1.0:world_size[0] * 1j, -1.0:1.0:world_size[1] * 1j, -1.0:1.0:world_size[2] * 1j]
        # ball init (:77-82): the radius itself in the geometry_searching stage, radius - 1 otherwise
        sdf = torch.from_numpy((x ** 2 + y ** 2 + z ** 2) ** 0.5 - (0 if stage == "geometry_searching" else 1)).float()[None, None, ...]
        k0 = torch.zeros([1, k0_dim, *world_size])
        dim = (3 + 3 * refbase_pe * 2) + k0_dim + (3 + 3 * posbase_pe * 2) + 3 + (3 + 3 * viewbase_pe * 2)   # :116-122
        ls = [nn.Linear(dim, refnet_width), nn.Linear(refnet_width, refnet_width), nn.Linear(refnet_width, 3)]
        refnet = [(l.weight.data.clone(), l.bias.data.clone()) for l in ls]
        if stage == "geometry_searching":                     # no MaskCache, no non-empty mask (:158-161)
            mask_cache = None
            mgrid, mmin, mmax, mthres = torch.zeros(1, 1, 2, 2, 2), xyz_min, xyz_max, mask_cache_thres
        else:
            if mask_cache is None:
                if mask_cache_path is None:
                    raise ValueError("the coarse stage needs the MaskCache of the earlier stage (mask_cache_path or mask_cache)")
                mask_cache = MaskCache(path=mask_cache_path, mask_cache_thres=mask_cache_thres)
            mgrid, mmin, mmax, mthres = mask_cache.sdf_mask, mask_cache.xyz_min.cpu(), mask_cache.xyz_max.cpu(), mask_cache.mask_cache_thres
        self.mask_cache = mask_cache
        cfg = dict(stage=stage, k0_dim=k0_dim, refnet_width=refnet_width, refnet_depth=refnet_depth, posbase_pe=posbase_pe,
                   viewbase_pe=viewbase_pe, refbase_pe=refbase_pe, smooth_ksize=smooth_ksize, smooth_sigma=smooth_sigma,
                   s_ratio=s_ratio, s_start=s_start, step_start=step_start, stepsize=stepsize,
                   fast_color_thres=fast_color_thres, mask_cache_thres=mthres)
        sc = Scene(G=int(world_size[0]), xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=float(voxel_size), sdf=sdf, k0=k0,
                   mask_grid=mgrid, mask_xyz_min=mmin, mask_xyz_max=mmax, mask_thres=mthres, cfg=cfg)
        self.engine = CoarseStepEngine(sc, refnet, n_rays, device, cfg=cfg, train_cfg=train_cfg)
        self.engine.num_voxels = num_voxels
        self.engine.num_voxels_base = num_voxels_base or num_voxels
        self.num_voxels, self.num_voxels_base = num_voxels, num_voxels_base or num_voxels
        self.engine.set_nonempty_mask()                       # (:170-172: also sdf = 1 outside the mask)
        self.sdf = _GridView(lambda: self.engine.sdf)
        self.k0 = _GridView(lambda: self.engine.k0)

    # ---- attributes the reference's callers read ----
    @property
    def world_size(self):
        return torch.tensor([self.engine.X, self.engine.Y, self.engine.Z])

    @property
    def voxel_size(self):
        return torch.tensor(float(self.engine.voxel_size))

    @property
    def xyz_min(self):
        return self.engine.xyz_min

    @property
    def xyz_max(self):
        return self.engine.xyz_max

    @property
    def nonempty_mask(self):
        m = self.engine.nonempty_mask
        return None if m is None else m.reshape(self.engine.sdf.shape)

    def set_inc_mask(self, lower, upper):
        """nerf.set_inc_mask (model/nerf.py:1078-1088), coarse stages only."""
        self.engine.set_inc_mask(lower, upper)

    def unset_inc_mask(self):
        if hasattr(self.engine, "unset_inc_mask"):
            self.engine.unset_inc_mask()

    @torch.no_grad()
    def forward(self, rays_o, rays_d, viewdirs, global_step=20000, **render_kwargs):
        """nerf.forward_fine (model/nerf.py:776-941) or, for stage='coarse', nerf.forward_coarse (:943-1075).  Keys and
        per-key shapes of the reference's ret_dict."""
        eng = self.engine
        if "stepsize" in render_kwargs and float(render_kwargs["stepsize"]) != float(eng.cfg["stepsize"]):
            raise RuntimeError(f"stepsize {render_kwargs['stepsize']} differs from the one the model was built with "
                               f"({eng.cfg['stepsize']})")
        n, cap = rays_o.shape[0], eng.n_rays
        if n == 0 or n > cap:
            raise RuntimeError(f"batch of {n} rays: the model was built for 1..{cap} rays per call")
        pad = (lambda t: t.contiguous()) if n == cap else \
            (lambda t: torch.cat([t, t[-1:].expand(cap - n, 3)]).contiguous())
        ro, rd, vd = pad(rays_o), pad(rays_d), pad(viewdirs)
        eng.forward(ro, rd, vd, global_step=global_step if global_step is not None else 20000,
                    near=render_kwargs.get("near", 2.0), bg=render_kwargs.get("bg", 1.0),
                    render_depth=render_kwargs.get("render_depth", False), render_grad=render_kwargs.get("render_grad", False))
        out = eng.outputs()
        out["s_val"] = eng.s_val(global_step) if global_step is not None else 0
        # mask_outbbox [total_len] (bookkeeping only in the reference, :694-697, :789): True except at the samples that
        # passed the bbox test and the MaskCache, i.e. the level-1 samples; rebuilt from the per-ray step counts
        l1 = out["level1"]
        ns = eng.ws["n_steps"][:n].long()
        off = torch.cumsum(ns, 0) - ns
        total = int(ns.sum())
        real = l1["ray_id"] < n
        mo = torch.ones(total, dtype=torch.bool, device=ns.device)
        if total > 0 and bool(real.any()):
            mo[(off[l1["ray_id"][real]] + l1["step_id"][real]).clamp_(0, total - 1)] = False
        if getattr(eng, "inc_mask", None) is not None:
            # Voxel Inc drops samples WITHOUT touching mask_outbbox (:962-967): take the bbox test from the sampling op
            mo = render_utils_cuda.sample_pts_on_rays(ro[:n].contiguous(), rd[:n].contiguous(), eng.xyz_min.to(ro.device),
                                                      eng.xyz_max.to(ro.device), render_kwargs.get("near", 2.0), 1e9,
                                                      float(eng.P.stepdist))[1]
        out["mask_outbbox"] = mo
        if n == cap:
            return out
        keep = out["ray_id"] < n                                  # drop the samples of the padding rays
        per_ray = ("alphainv_cum", "rgb_marched", "sigmoid_rgb", "normal_marched", "depth", "disp")
        per_sample = ("weights", "ray_id", "step_id", "viewdirs", "normal", "raw_alpha", "raw_rgb", "gradient")
        res = dict(out)
        for k in per_ray:
            res[k] = out[k][:n] if out[k] is not None else None
        for k in per_sample:
            res[k] = out[k][keep]
        # `mask` (weights > thres over the alpha-pruned samples) keeps its full length: its samples are ray-major, so the
        # real rays' entries are a prefix; the padding rays' entries are dropped through level 1's ray ids
        res["mask"] = out["mask"][(l1["ray_id"][(l1["flag"] & 1) != 0]) < n]
        res.pop("level1", None)
        return res

    def train_step(self, rays_o, rays_d, viewdirs, target, global_step, **render_kwargs):
        """One iteration of the reference's loop body (model/nerf_training.py:300-373), fused."""
        self.engine.train_step(rays_o, rays_d, viewdirs, target, global_step, near=render_kwargs.get("near", 2.0),
                               bg=render_kwargs.get("bg", 1.0))

    def state_dict(self, *args, **kwargs):
        from .checkpoint import model_state_dict
        return model_state_dict(self.engine)

    @torch.no_grad()
    def load_state_dict(self, sd, strict=False):
        """Reference layout (tests/golden/ckpt_layout.json); unknown keys are ignored like strict=False does."""
        from .checkpoint import _LIN
        eng = self.engine
        for key, dst in (("sdf.grid", eng.sdf), ("k0.grid", eng.k0)):
            if key in sd:
                if tuple(sd[key].shape) != tuple(dst.shape):
                    raise RuntimeError(f"{key}: checkpoint {tuple(sd[key].shape)} vs model {tuple(dst.shape)} "
                                       f"(scale_volume_grid first, as utils.load_checkpoint does)")
                dst.copy_(sd[key].to(dst.device))
        if self.stage in ("coarse", "geometry_searching"):
            from .checkpoint import _LIN3
            for j, idx in enumerate(_LIN3):
                if f"refnet.{idx}.weight" in sd:
                    eng.weights[j].copy_(sd[f"refnet.{idx}.weight"].to(eng.device))
                    eng.biases[j].copy_(sd[f"refnet.{idx}.bias"].to(eng.device))
            return
        for name, base in (("rgbnet", 0), ("refnet", 4)):
            for j, idx in enumerate(_LIN):
                if f"{name}.{idx}.weight" in sd:
                    eng.weights[base + j].copy_(sd[f"{name}.{idx}.weight"].to(eng.device))
                    eng.Bt[base + j].copy_(sd[f"{name}.{idx}.bias"].to(eng.device))

    def scale_volume_grid(self, num_voxels):
        self.engine = self.engine.scale_volume_grid(num_voxels)
        self.num_voxels = num_voxels

    def get_kwargs(self):
        from .checkpoint import checkpoint_dict
        return checkpoint_dict(self.engine, 0)["model_kwargs"]

    def extract_fields(self, bound_min, bound_max, resolution=128):
        """The field nerf.extract_geometry hands to marching cubes (model/nerf.py:1157-1170)."""
        return self.engine.extract_fields(bound_min, bound_max, resolution)
