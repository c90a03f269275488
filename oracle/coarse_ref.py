"""TEST INFRASTRUCTURE (oracle): torch restatement of the reference's COARSE-stage forward and loss block
 -- `nerf.forward_coarse` (model/nerf.py:943-1075) with the helpers it calls, and the loss lines of
model/nerf_training.py:307-345 for a `coarse_train`-style config (weight_rgbper, ori_tv).  Same library ops as the
reference (F.grid_sample, nn.Conv3d semantics via F.conv3d + replicate pad, F.linear, index_add_ for
torch_scatter.segment_coo).  Only tests/ may import it; no product code path does (the kernels it checks: fgs_nerf_b200/csrc/coarse.cu).

PINNED against tests/golden/coarse_golden.npz, produced by importing the reference's real `nerf` (stage='coarse') in
this container on the CPU (tests/golden/make_coarse_golden.py): tests/test_oracle_coarse.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import nerf_ref as R


class CoarseModel:
    """Parameters of the coarse-stage `nerf` module as plain leaf tensors: sdf [1,1,X,Y,Z], k0 [1,C,X,Y,Z],
    refnet [(W, b)] (model/nerf.py:113-133: no rgbnet in this stage), the Gaussian smoother of init_smooth_conv."""

    def __init__(self, scene, refnet, cfg, device="cpu", requires_grad=True):
        dev = torch.device(device)
        self.device = dev
        self.cfg = dict(cfg)
        self.xyz_min, self.xyz_max = scene.xyz_min.to(dev), scene.xyz_max.to(dev)
        self.voxel_size = torch.tensor(scene.voxel_size, dtype=torch.float32, device=dev)
        self.world_size = torch.tensor(list(scene.sdf.shape[2:]), device=dev)
        self.sdf = scene.sdf.clone().to(dev).requires_grad_(requires_grad)
        self.k0 = scene.k0.clone().contiguous().to(dev).requires_grad_(requires_grad)
        self.refnet = [(w.clone().to(dev).requires_grad_(requires_grad), b.clone().to(dev).requires_grad_(requires_grad))
                       for w, b in refnet]
        self.mask_grid = scene.mask_grid.to(dev)
        self.mask_xyz_min, self.mask_xyz_max = scene.mask_xyz_min.to(dev), scene.mask_xyz_max.to(dev)
        self.mask_thres = cfg["mask_cache_thres"]
        self.posfreq = torch.tensor([2.0 ** i for i in range(cfg["posbase_pe"])], device=dev)
        self.viewfreq = torch.tensor([2.0 ** i for i in range(cfg["viewbase_pe"])], device=dev)
        self.reffreq = torch.tensor([2.0 ** i for i in range(cfg["refbase_pe"])], device=dev)
        self.fast_color_thres = cfg["fast_color_thres"]
        self.cfg.setdefault("step_start", 0)
        self.smooth_weight = gaussian_3dconv_weight(cfg["smooth_ksize"], cfg["smooth_sigma"], dev) \
            if cfg.get("smooth_ksize", 0) > 0 else None
        self.nonempty_mask = None
        self.stage = cfg.get("stage", "coarse")           # 'coarse' | 'geometry_searching' (model/nerf.py:47-50)
        self.inc_mask = None                              # (mask bool [X,Y,Z], xyz2ijk_scale, xyz2ijk_shift) of a MaskGrid

    def params(self):
        out = [("sdf", self.sdf), ("k0", self.k0)]
        for i, (w, b) in enumerate(self.refnet):
            out += [(f"refnet.{i}.weight", w), (f"refnet.{i}.bias", b)]
        return out

    @torch.no_grad()
    def set_nonempty_mask(self):
        """model/nerf.py:339-353; in the coarse stage the SDF outside the mask is overwritten with 1.  The
        geometry_searching stage has no MaskCache and no non-empty mask (:158-161)."""
        if self.stage != "coarse":
            return None
        R.FineModel.set_nonempty_mask(self)
        self.sdf[~self.nonempty_mask] = 1
        return self.nonempty_mask

    @torch.no_grad()
    def set_inc_mask(self, lower, upper):
        """nerf.set_inc_mask (model/nerf.py:1078-1088) + MaskGrid.__init__ (model/grid.py:253-272): the voxels of the
        unit-cube lattice inside [lower, upper] (Incremental Voxel Training), looked up with maskcache_lookup."""
        X, Y, Z = (int(v) for v in self.sdf.shape[2:])
        xyz = torch.stack(torch.meshgrid(torch.linspace(0, 1, X), torch.linspace(0, 1, Y), torch.linspace(0, 1, Z), indexing="ij"))
        mask = (xyz[0] >= lower[0]) & (xyz[0] <= upper[0]) & (xyz[1] >= lower[1]) & (xyz[1] <= upper[1]) & \
               (xyz[2] >= lower[2]) & (xyz[2] <= upper[2])
        xyz_min, xyz_max = self.xyz_min.cpu(), self.xyz_max.cpu()
        scale = (torch.Tensor(list(mask.shape)) - 1) / (xyz_max - xyz_min)
        shift = -xyz_min * scale
        dev = self.device
        self.inc_mask = (mask.to(dev).contiguous(), scale.to(dev), shift.to(dev))

    def unset_inc_mask(self):
        self.inc_mask = None


def gaussian_3dconv_weight(ksize, sigma, device):
    """model/nerf.py:260-268 -> Conv3d weight [1,1,k,k,k]."""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    kernel = torch.from_numpy(np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2))).to(device=device, dtype=torch.float32)
    return (kernel[None, None, ...] / kernel.sum())


def smooth_conv(m, grid):
    """nn.Conv3d(1, 1, k, padding=k//2, padding_mode='replicate'), zero bias (model/nerf.py:267-269)."""
    h = m.smooth_weight.shape[-1] // 2
    return F.conv3d(F.pad(grid, (h,) * 6, mode="replicate"), m.smooth_weight)


def sample_grid(m, xyz, grid):
    """nerf.grid_sampler with sample_ret only (model/nerf.py:639-659): any channel count; [M] for one channel."""
    shape = xyz.shape[:-1]
    xyz_ = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz_ - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
    return F.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True).reshape(
        grid.shape[1], -1).T.reshape(*shape, grid.shape[1]).squeeze(-1)


def forward_coarse(m, native, rays_o, rays_d, viewdirs, global_step, near, stepsize, bg, render_depth=True,
                   render_grad=True, stages=None):
    """model/nerf.py:943-1075 (stages 'coarse' and 'geometry_searching', optional inc_mask, use_viewdir)."""
    A2W = R.make_alphas2weights(native)
    N = len(rays_o)
    # sample_ray_cuda (:700-732): far = 1e9, masked to the bbox
    stepdist = stepsize * m.voxel_size
    ray_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max = native.sample_pts_on_rays(
        rays_o.contiguous(), rays_d.contiguous(), m.xyz_min, m.xyz_max, near, 1e9, float(stepdist))
    mask_inbbox = ~mask_outbbox
    ray_pts, ray_id, step_id = ray_pts[mask_inbbox], ray_id[mask_inbbox], step_id[mask_inbbox]
    viewdirs_pts = viewdirs[ray_id]
    if m.stage == "coarse":                                           # :952-960 (geometry_searching: no MaskCache)
        mask = R.mask_cache(m, ray_pts)
        ray_pts, ray_id, viewdirs_pts, step_id = ray_pts[mask], ray_id[mask], viewdirs_pts[mask], step_id[mask]
        mask_outbbox[~mask_outbbox] |= ~mask
    if m.inc_mask is not None:                                        # :962-967 Voxel Inc (MaskGrid.forward, grid.py:274-284)
        imask, iscale, ishift = m.inc_mask
        mask = native.maskcache_lookup(imask, ray_pts.reshape(-1, 3).contiguous(), iscale, ishift)
        ray_pts, ray_id, viewdirs_pts, step_id = ray_pts[mask], ray_id[mask], viewdirs_pts[mask], step_id[mask]
    sdf_grid = smooth_conv(m, m.sdf) if m.smooth_weight is not None else m.sdf     # :969
    sdf = sample_grid(m, ray_pts, sdf_grid)                            # :971
    grad_grid = R.neus_sdf_gradient(m)                                 # :972, from the RAW sdf grid ('interpolate')
    gradient = sample_grid(m, ray_pts, grad_grid)                      # :973
    dist = stepsize * m.voxel_size
    s_val, alpha = R.neus_alpha_from_sdf_scatter(m, viewdirs, ray_id, dist, sdf, gradient, global_step)
    if stages is not None:
        stages["M1"] = int(ray_id.numel())
    weights, alphainv_last = A2W.apply(alpha, ray_id, N)               # :978
    mask = None
    if m.fast_color_thres > 0:                                         # :980-988: ONE prune, on the weights
        mask = weights > m.fast_color_thres
        ray_pts, ray_id, viewdirs_pts, step_id = ray_pts[mask], ray_id[mask], viewdirs_pts[mask], step_id[mask]
        alpha, gradient = alpha[mask], gradient[mask]
    weights, alphainv_last = A2W.apply(alpha, ray_id, N)               # :990: recomputed on the pruned samples
    if stages is not None:
        stages["M2"] = int(ray_id.numel())
    normal = R.l2_normalize(gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-7))
    rays_xyz = (ray_pts - m.xyz_min) / (m.xyz_max - m.xyz_min)
    xyz_emb = (rays_xyz.unsqueeze(-1) * m.posfreq).flatten(-2)
    xyz_emb = torch.cat([rays_xyz, xyz_emb.sin(), xyz_emb.cos()], -1)
    k0 = R.dense_grid(m, m.k0, ray_pts)
    reflect_r = viewdirs_pts - 2. * torch.sum(viewdirs_pts * normal, dim=-1, keepdim=True) * normal
    reflect_emb = (reflect_r.unsqueeze(-1) * m.reffreq).flatten(-2)
    reflect_emb = torch.cat([reflect_r, reflect_emb.sin(), reflect_emb.cos()], -1)
    viewdirs_emb = (viewdirs.unsqueeze(-1) * m.viewfreq).flatten(-2)
    viewdirs_emb = torch.cat([viewdirs, viewdirs_emb.sin(), viewdirs_emb.cos()], -1)
    viewdirs_emb = viewdirs_emb.flatten(0, -2)[ray_id]
    ref_feat = torch.cat([k0, xyz_emb, reflect_emb, normal, viewdirs_emb], dim=-1)            # :1005
    if stages is not None:
        stages["ref_feat"] = ref_feat.detach()
    rgb = torch.sigmoid(R.mlp(ref_feat, m.refnet))
    sigmoid_rgb = torch.sigmoid(rgb)
    rgb_marched = R.segment_sum(weights.unsqueeze(-1) * rgb, ray_id, N)
    sigmoid_rgb = R.segment_sum(weights.unsqueeze(-1) * sigmoid_rgb, ray_id, N)
    cum_weights = R.segment_sum(weights.unsqueeze(-1), ray_id, N)
    rgb_marched = (rgb_marched + (1 - cum_weights) * bg).clamp(0, 1)
    sigmoid_rgb = (sigmoid_rgb + (1 - cum_weights) * bg).clamp(0, 1)
    normal_marched = R.segment_sum(weights.unsqueeze(-1) * normal, ray_id, N) if render_grad else None
    depth = disp = None
    if render_depth:
        with torch.no_grad():
            depth = R.segment_sum(weights * step_id * dist, ray_id, N)
            disp = 1 / depth
    return {
        'alphainv_cum': alphainv_last, 'weights': weights, 'ray_id': ray_id, 'step_id': step_id,
        'viewdirs': viewdirs[ray_id], 'rgb_marched': rgb_marched, 'sigmoid_rgb': sigmoid_rgb,
        'normal_marched': normal_marched, 'normal': normal, 'raw_alpha': alpha, 'raw_rgb': rgb, 'depth': depth,
        'disp': disp, 'mask': mask, 'mask_outbbox': mask_outbbox, 'gradient': gradient, 's_val': s_val,
    }


def total_variation(v, mask=None):
    """model/nerf.py:1212-1221."""
    tv2, tv3, tv4 = v.diff(dim=2).abs(), v.diff(dim=3).abs(), v.diff(dim=4).abs()
    if mask is not None:
        tv2 = tv2[mask[:, :, :-1] & mask[:, :, 1:]]
        tv3 = tv3[mask[:, :, :, :-1] & mask[:, :, :, 1:]]
        tv4 = tv4[mask[:, :, :, :, :-1] & mask[:, :, :, :, 1:]]
        return (tv2.sum() + tv3.sum() + tv4.sum()) / 3 / mask.sum()
    return (tv2.sum() + tv3.sum() + tv4.sum()) / 3 / v.sum()


def coarse_loss(m, out, target, t):
    """model/nerf_training.py:307-345 for a coarse_train-style config `t` (weight_rgbper > 0, ori_tv: the SDF TV goes
    through autograd, nerf.density_total_variation model/nerf.py:430-447).  Returns (loss, tv_smooth, tv_sdf)."""
    n = len(target)
    loss = t["weight_main"] * F.mse_loss(out["rgb_marched"], target)
    if t["weight_rgbper"] > 0:
        rgbper = (out["raw_rgb"] - target[out["ray_id"]]).pow(2).sum(-1)
        loss = loss + t["weight_rgbper"] * (rgbper * out["weights"].detach()).sum() / n
    if t["weight_entropy_last"] > 0:
        pout = out["alphainv_cum"][..., -1].clamp(1e-6, 1 - 1e-6)
        loss = loss + t["weight_entropy_last"] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    if t["weight_orientation"] > 0:
        zero = torch.tensor(0.0, dtype=torch.float32, device=target.device)
        w = out["weights"].detach()                        # nerf.orientation_loss, model/nerf.py:469-478
        v = -out["viewdirs"]
        n_dot_v = (out["normal"] * v).sum(dim=-1)
        ori = torch.mean((w * torch.fmin(zero, n_dot_v) ** 2).sum(dim=-1))
        loss = loss + t["weight_orientation"] * ori
    if t["sigmoid_rgb_loss"] > 0:
        loss = loss + t["sigmoid_rgb_loss"] * F.mse_loss(out["sigmoid_rgb"], target)
    tv_smooth = t["weight_tv_density"] * R.smooth_grad_tv(m, t["smooth_grad_tv"])
    tv_sdf = t["weight_tv_density"] * (total_variation(m.sdf, m.nonempty_mask) / 2 / m.voxel_size * t["sdf_tv"])
    return loss, tv_smooth, tv_sdf
