"""TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's fine-stage hot path:
`nerf.forward_fine` (model/nerf.py:776-941), the loss block of the training loop
(model/nerf_training.py:307-347), the TV add-grad (model/nerf_training.py:353-371) and
`MaskedAdam.step` (model/adam.py:190-221).

It uses the SAME library ops the reference uses (F.grid_sample, nn.functional.linear, sigmoid,
sin/cos, index_add_ for torch_scatter.segment_coo) in the same order, so on a given device it
is the reference's arithmetic.  The three native ops it needs (sample_pts_on_rays, alpha2weight,
alpha2weight_backward; plus adam / TV for the step) are injected through `native`:
  * on the CPU: oracle.native_ref (the C restatement)  -> CPU baseline and golden fixtures,
  * on the GPU box: the reference's own compiled kernels (oracle/_ref via oracle.ref_ext).
Pinned by tests/golden/fine_golden.npz, which tests/golden/make_fine_golden.py produces by
importing the REAL model/nerf.py from /root/reference (with stubs for its missing third-party
imports) and running it on the same inputs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# native-op adapters
# ------------------------------------------------------------------------------------------
class CpuNative:
    """oracle.native_ref behind the reference's torch-tensor calling convention."""

    def __init__(self):
        from oracle import native_ref
        self.o = native_ref

    @staticmethod
    def _t(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def sample_pts_on_rays(self, rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
        out = self.o.sample_pts_on_rays(rays_o.numpy(), rays_d.numpy(), xyz_min.numpy(), xyz_max.numpy(),
                                        float(near), float(far), float(stepdist))
        return [self._t(a) for a in out]

    def alpha2weight(self, alpha, ray_id, n_rays):
        return [self._t(a) for a in self.o.alpha2weight(alpha.detach().numpy(), ray_id.numpy(), n_rays)]

    def alpha2weight_backward(self, alpha, weight, T, alphainv_last, i_start, i_end, n_rays, gw, gl):
        return self._t(self.o.alpha2weight_backward(alpha.detach().numpy(), weight.numpy(), T.numpy(),
                                                    alphainv_last.numpy(), i_start.numpy(), i_end.numpy(),
                                                    n_rays, gw.contiguous().numpy(), gl.contiguous().numpy()))

    def maskcache_lookup(self, mask, xyz, scale, shift):
        return self._t(self.o.maskcache_lookup(mask.numpy(), xyz.numpy(), scale.numpy(), shift.numpy()))

    def adam(self, mode, param, grad, m, v, step, beta1, beta2, lr, eps):
        fn = self.o.masked_adam_upd if mode == "masked" else self.o.adam_upd
        p = param.detach().numpy()
        fn(p.reshape(-1), grad.contiguous().numpy().reshape(-1), m.numpy().reshape(-1), v.numpy().reshape(-1),
           step, beta1, beta2, lr, eps)

    def tv_add_grad(self, param, grad, wx, wy, wz, dense):
        self.o.total_variation_add_grad(param.detach().numpy(), grad.numpy(), wx, wy, wz, dense)


class RefCudaNative:
    """The reference's own compiled CUDA kernels (GPU box only)."""

    def __init__(self):
        from oracle import ref_ext
        self.ru = ref_ext.load("render_utils_cuda")
        self.ad = ref_ext.load("adam_upd_cuda")
        self.tv = ref_ext.load("total_variation_cuda")

    def sample_pts_on_rays(self, *a):
        a = list(a)
        return self.ru.sample_pts_on_rays(a[0], a[1], a[2], a[3], float(a[4]), float(a[5]), float(a[6]))

    def alpha2weight(self, alpha, ray_id, n_rays):
        return self.ru.alpha2weight(alpha.detach().contiguous(), ray_id.contiguous(), n_rays)

    def alpha2weight_backward(self, alpha, weight, T, last, i_s, i_e, n_rays, gw, gl):
        return self.ru.alpha2weight_backward(alpha.detach(), weight, T, last, i_s, i_e, n_rays,
                                             gw.contiguous(), gl.contiguous())

    def maskcache_lookup(self, mask, xyz, scale, shift):
        return self.ru.maskcache_lookup(mask, xyz, scale, shift)

    def adam(self, mode, param, grad, m, v, step, beta1, beta2, lr, eps):
        fn = self.ad.masked_adam_upd if mode == "masked" else self.ad.adam_upd
        fn(param.detach(), grad.contiguous(), m, v, step, beta1, beta2, lr, eps)

    def tv_add_grad(self, param, grad, wx, wy, wz, dense):
        self.tv.total_variation_add_grad(param.detach(), grad, wx, wy, wz, dense)


def make_alphas2weights(native):
    """model/nerf.py:1173-1189."""

    class Alphas2Weights(torch.autograd.Function):
        @staticmethod
        def forward(ctx, alpha, ray_id, N):
            weights, T, alphainv_last, i_start, i_end = native.alpha2weight(alpha, ray_id, N)
            if alpha.requires_grad:
                ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
                ctx.n_rays = N
            return weights, alphainv_last

        @staticmethod
        @torch.autograd.function.once_differentiable
        def backward(ctx, grad_weights, grad_last):
            alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
            grad = native.alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end,
                                                ctx.n_rays, grad_weights, grad_last)
            return grad, None, None

    return Alphas2Weights


# ------------------------------------------------------------------------------------------
# model state
# ------------------------------------------------------------------------------------------
class FineModel:
    """Parameters of the fine-stage `nerf` module as plain leaf tensors.
    sdf [1,1,X,Y,Z], k0 [1,C,X,Y,Z], rgbnet/refnet: list of (W [out,in], b [out])."""

    def __init__(self, scene, device="cpu", requires_grad=True):
        dev = torch.device(device)
        self.device = dev
        self.cfg = dict(scene.cfg)
        self.xyz_min = scene.xyz_min.to(dev)
        self.xyz_max = scene.xyz_max.to(dev)
        self.voxel_size = torch.tensor(scene.voxel_size, dtype=torch.float32, device=dev)
        self.world_size = torch.tensor(list(scene.sdf.shape[2:]), device=dev)
        self.sdf = scene.sdf.clone().to(dev).requires_grad_(requires_grad)
        self.k0 = scene.k0.clone().contiguous().to(dev).requires_grad_(requires_grad)
        self.rgbnet = [(w.clone().to(dev).requires_grad_(requires_grad), b.clone().to(dev).requires_grad_(requires_grad))
                       for w, b in scene.rgbnet]
        self.refnet = [(w.clone().to(dev).requires_grad_(requires_grad), b.clone().to(dev).requires_grad_(requires_grad))
                       for w, b in scene.refnet]
        self.mask_grid = scene.mask_grid.to(dev)
        self.mask_xyz_min = scene.mask_xyz_min.to(dev)
        self.mask_xyz_max = scene.mask_xyz_max.to(dev)
        self.mask_thres = scene.mask_thres
        c = self.cfg
        self.posfreq = torch.tensor([2.0 ** i for i in range(c["posbase_pe"])], device=dev)
        self.viewfreq = torch.tensor([2.0 ** i for i in range(c["viewbase_pe"])], device=dev)
        self.reffreq = torch.tensor([2.0 ** i for i in range(c["refbase_pe"])], device=dev)
        self.fast_color_thres = c["fast_color_thres"]
        self.nonempty_mask = None

    def params(self):
        out = [("sdf", self.sdf), ("k0", self.k0)]
        for i, (w, b) in enumerate(self.rgbnet):
            out += [(f"rgbnet.{i}.weight", w), (f"rgbnet.{i}.bias", b)]
        for i, (w, b) in enumerate(self.refnet):
            out += [(f"refnet.{i}.weight", w), (f"refnet.{i}.bias", b)]
        return out

    # model/nerf.py:339-353
    @torch.no_grad()
    def set_nonempty_mask(self):
        X, Y, Z = self.sdf.shape[2:]
        g = torch.stack(torch.meshgrid(
            torch.linspace(float(self.xyz_min[0]), float(self.xyz_max[0]), X),
            torch.linspace(float(self.xyz_min[1]), float(self.xyz_max[1]), Y),
            torch.linspace(float(self.xyz_min[2]), float(self.xyz_max[2]), Z), indexing="ij"), -1).to(self.device)
        self.nonempty_mask = mask_cache(self, g).reshape(self.sdf.shape).contiguous()
        return self.nonempty_mask


def mlp(x, layers):
    """nn.Sequential(Linear, ReLU, ..., Linear) of model/nerf.py:125-142."""
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        if i + 1 < len(layers):
            x = F.relu(x)
    return x


# model/nerf.py:1202-1209
@torch.no_grad()
def mask_cache(m, xyz):
    shape = xyz.shape[:-1]
    xyz = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - m.mask_xyz_min) / (m.mask_xyz_max - m.mask_xyz_min)).flip((-1,)) * 2 - 1
    v = F.grid_sample(m.mask_grid, ind_norm, align_corners=True)
    return v.reshape(*shape) >= m.mask_thres


# model/nerf.py:597-637
def sample_sdfs(m, xyz, grid, displace_list, use_grad_norm=False):
    shape = xyz.shape[:-1]
    xyz = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
    grid_size = grid.size()[-3:]
    size_factor_zyx = torch.tensor([grid_size[2], grid_size[1], grid_size[0]], device=xyz.device)
    ind = ((ind_norm + 1) / 2) * (size_factor_zyx - 1)
    offset = torch.tensor([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], device=xyz.device)
    displace = torch.tensor(displace_list, device=xyz.device)
    offset = offset[:, None, :] * displace[None, :, None]
    all_ind = ind.unsqueeze(-2) + offset.view(-1, 3)
    all_ind = all_ind.view(1, 1, 1, -1, 3)
    all_ind[..., 0] = all_ind[..., 0].clamp(min=0, max=size_factor_zyx[0] - 1)
    all_ind[..., 1] = all_ind[..., 1].clamp(min=0, max=size_factor_zyx[1] - 1)
    all_ind[..., 2] = all_ind[..., 2].clamp(min=0, max=size_factor_zyx[2] - 1)
    all_ind_norm = (all_ind / (size_factor_zyx - 1)) * 2 - 1
    feat = F.grid_sample(grid, all_ind_norm, mode="bilinear", align_corners=True)
    nd = len(displace_list)
    all_ind = all_ind.view(1, 1, 1, -1, 6, nd, 3)
    diff = all_ind[:, :, :, :, 1::2, :, :] - all_ind[:, :, :, :, 0::2, :, :]
    diff, _ = diff.max(dim=-1)
    feat_ = feat.view(1, 1, 1, -1, 6, nd)
    feat_diff = feat_[:, :, :, :, 1::2, :] - feat_[:, :, :, :, 0::2, :]
    grad = feat_diff / diff / m.voxel_size
    feat = feat.view(shape[-1], 6, nd)
    grad = grad.view(shape[-1], 3, nd)
    if use_grad_norm:
        grad = grad / (grad.norm(dim=1, keepdim=True) + 1e-5)
    feat = feat.view(shape[-1], 6 * nd)
    grad = grad.view(shape[-1], 3 * nd)
    return feat, grad


# model/nerf.py:639-672 (sample_ret=True, sample_grad=True)
def grid_sampler(m, xyz, grid):
    shape = xyz.shape[:-1]
    xyz_ = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz_ - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
    ret = F.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True).reshape(
        grid.shape[1], -1).T.reshape(*shape, grid.shape[1]).squeeze(-1)
    feat, grad = sample_sdfs(m, xyz_, grid, displace_list=[1.0], use_grad_norm=False)
    feat = torch.cat([feat[:, 4:6], feat[:, 2:4], feat[:, 0:2]], dim=-1)
    grad = torch.cat([grad[:, [2]], grad[:, [1]], grad[:, [0]]], dim=-1)
    return ret, grad, feat


# model/grid.py:49-59
def dense_grid(m, grid, xyz):
    shape = xyz.shape[:-1]
    xyz = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True)
    out = out.reshape(grid.shape[1], -1).T.reshape(*shape, grid.shape[1])
    if grid.shape[1] == 1:
        out = out.squeeze(-1)
    return out


# model/nerf.py:510-544
def neus_alpha_from_sdf_scatter(m, viewdirs, ray_id, dist, sdf, gradients, global_step):
    c = m.cfg
    s_val = 1. / (global_step + c["s_ratio"] / c["s_start"] - c["step_start"]) * c["s_ratio"]
    s_val_t = torch.ones(1, device=sdf.device) * s_val
    dirs = viewdirs[ray_id]
    inv_s = torch.ones(1, device=sdf.device) / s_val_t
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    cos_anneal_ratio = 1.0
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    sdf = sdf.unsqueeze(-1)
    estimated_next_sdf = sdf + iter_cos * dist.reshape(-1, 1) * 0.5
    estimated_prev_sdf = sdf - iter_cos * dist.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(estimated_prev_sdf * inv_s.reshape(-1, 1))
    next_cdf = torch.sigmoid(estimated_next_sdf * inv_s.reshape(-1, 1))
    p = prev_cdf - next_cdf
    cc = prev_cdf
    alpha = ((p + 1e-5) / (cc + 1e-5)).clip(0.0, 1.0).squeeze()
    return s_val, alpha


def l2_normalize(x, eps=torch.finfo(torch.float32).eps):
    eps = torch.tensor(eps, device=x.device)
    return x / torch.sqrt(torch.maximum(torch.sum(x ** 2, dim=-1, keepdims=True), eps))


def segment_sum(src, index, n):
    """torch_scatter.segment_coo(src, index, out=zeros, reduce='sum') (model/nerf.py:888-896):
    sorted-index segment sum; backward is a gather."""
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


# ------------------------------------------------------------------------------------------
# model/nerf.py:776-941
# ------------------------------------------------------------------------------------------
def forward_fine(m, native, rays_o, rays_d, viewdirs, global_step, near, far, stepsize, bg,
                 render_depth=False, render_grad=False, stages=None):
    A2W = make_alphas2weights(native)
    N = len(rays_o)
    c = m.cfg
    # sample_ray (674-698)
    stepdist = stepsize * m.voxel_size
    ray_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max = native.sample_pts_on_rays(
        rays_o.contiguous(), rays_d.contiguous(), m.xyz_min, m.xyz_max, near, 1e9, float(stepdist))
    mask_inbbox = ~mask_outbbox
    ray_pts = ray_pts[mask_inbbox]
    ray_id = ray_id[mask_inbbox]
    step_id = step_id[mask_inbbox]
    if stages is not None:
        stages["total_len"] = int(mask_outbbox.numel())
        stages["M0"] = int(ray_id.numel())
    # MaskCache (784-789)
    mask = mask_cache(m, ray_pts)
    ray_pts = ray_pts[mask]
    ray_id = ray_id[mask]
    step_id = step_id[mask]
    mask_outbbox[~mask_outbbox] |= ~mask
    if stages is not None:
        stages["M1"] = int(ray_id.numel())
        stages["ray_id_1"], stages["step_id_1"] = ray_id, step_id
    sdf_grid = m.sdf
    sdf, gradient, feat = grid_sampler(m, ray_pts, sdf_grid)
    dist = stepsize * m.voxel_size
    s_val, alpha = neus_alpha_from_sdf_scatter(m, viewdirs, ray_id, dist, sdf, gradient, global_step)
    if stages is not None:
        stages["sdf_1"], stages["gradient_1"], stages["alpha_1"] = sdf.detach(), gradient.detach(), alpha.detach()
    mask = None
    viewdirs_pts = viewdirs[ray_id]
    if m.fast_color_thres > 0:
        mask = (alpha > m.fast_color_thres)
        alpha = alpha[mask]
        ray_id = ray_id[mask]
        viewdirs_pts = viewdirs_pts[mask]
        ray_pts = ray_pts[mask]
        step_id = step_id[mask]
        gradient = gradient[mask]
        sdf = sdf[mask]
    if stages is not None:
        stages["Ma"] = int(ray_id.numel())
    weights, alphainv_last = A2W.apply(alpha, ray_id, N)
    if m.fast_color_thres > 0:
        mask = (weights > m.fast_color_thres)
        weights = weights[mask]
        alpha = alpha[mask]
        ray_pts = ray_pts[mask]
        viewdirs_pts = viewdirs_pts[mask]
        ray_id = ray_id[mask]
        step_id = step_id[mask]
        gradient = gradient[mask]
        sdf = sdf[mask]
    if stages is not None:
        stages["M2"] = int(ray_id.numel())
    normal = l2_normalize(gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-7))
    rays_xyz = (ray_pts - m.xyz_min) / (m.xyz_max - m.xyz_min)
    xyz_emb = (rays_xyz.unsqueeze(-1) * m.posfreq).flatten(-2)
    xyz_emb = torch.cat([rays_xyz, xyz_emb.sin(), xyz_emb.cos()], -1)
    k0 = dense_grid(m, m.k0, ray_pts)
    all_grad_inds = sorted(list(set(c["grad_feat"] + (1.0,))))
    all_feat, all_grad = sample_sdfs(m, ray_pts, sdf_grid, displace_list=list(all_grad_inds), use_grad_norm=True)
    hierarchical_feats = [sdf[:, None], all_feat, all_grad]
    all_feats_ = gradient
    viewdirs_emb = (viewdirs.unsqueeze(-1) * m.viewfreq).flatten(-2)
    viewdirs_emb = torch.cat([viewdirs, viewdirs_emb.sin(), viewdirs_emb.cos()], -1)
    viewdirs_emb = viewdirs_emb.flatten(0, -2)[ray_id]
    rgb_feat_in = torch.cat([k0, xyz_emb, viewdirs_emb, *hierarchical_feats, all_feats_], dim=-1)
    rgb_feat = mlp(rgb_feat_in, m.rgbnet)
    reflect_r = viewdirs_pts - 2. * torch.sum(viewdirs_pts * normal, dim=-1, keepdim=True) * normal
    reflect_emb = (reflect_r.unsqueeze(-1) * m.reffreq).flatten(-2)
    reflect_emb = torch.cat([reflect_r, reflect_emb.sin(), reflect_emb.cos()], -1)
    ref_feat = torch.cat([rgb_feat, reflect_emb], dim=-1)
    rgb = torch.sigmoid(mlp(ref_feat, m.refnet))
    sigmoid_rgb = torch.sigmoid(rgb)
    rgb_marched = segment_sum(weights.unsqueeze(-1) * rgb, ray_id, N)
    cum_weights = segment_sum(weights.unsqueeze(-1), ray_id, N)
    sigmoid_rgb = segment_sum(weights.unsqueeze(-1) * sigmoid_rgb, ray_id, N)
    rgb_marched = (rgb_marched + (1 - cum_weights) * bg).clamp(0, 1)
    sigmoid_rgb = (sigmoid_rgb + (1 - cum_weights) * bg).clamp(0, 1)
    normal_marched = segment_sum(weights.unsqueeze(-1) * normal, ray_id, N) if render_grad else None
    depth = disp = None
    if render_depth:
        with torch.no_grad():
            depth = segment_sum(weights * step_id * dist, ray_id, N)
            disp = 1 / depth
    if stages is not None:
        stages["rgb_feat_in"] = rgb_feat_in.detach()
        stages["reflect_emb"] = reflect_emb.detach()
    return {
        'alphainv_cum': alphainv_last, 'weights': weights, 'ray_id': ray_id, 'step_id': step_id,
        'viewdirs': viewdirs[ray_id], 'rgb_marched': rgb_marched, 'sigmoid_rgb': sigmoid_rgb,
        'normal_marched': normal_marched, 'normal': normal, 'raw_alpha': alpha, 'raw_rgb': rgb,
        'depth': depth, 'disp': disp, 'mask': mask, 'mask_outbbox': mask_outbbox, 'gradient': gradient,
        's_val': s_val,
    }


# model/nerf.py:485-494 ('interpolate' mode)
def neus_sdf_gradient(m):
    sdf = m.sdf
    gradient = torch.zeros([1, 3] + [*sdf.shape[-3:]], device=sdf.device)
    gradient[:, 0, 1:-1, :, :] = (sdf[:, 0, 2:, :, :] - sdf[:, 0, :-2, :, :]) / 2 / m.voxel_size
    gradient[:, 1, :, 1:-1, :] = (sdf[:, 0, :, 2:, :] - sdf[:, 0, :, :-2, :]) / 2 / m.voxel_size
    gradient[:, 2, :, :, 1:-1] = (sdf[:, 0, :, :, 2:] - sdf[:, 0, :, :, :-2]) / 2 / m.voxel_size
    return gradient


def tv_smooth_conv_weight(device):
    """model/nerf.py:226-258 (sigma=0): 3x3x3 binomial kernel / its sum."""
    k = np.asarray([[[1, 2, 1], [2, 4, 2], [1, 2, 1]], [[2, 4, 2], [4, 8, 4], [2, 4, 2]], [[1, 2, 1], [2, 4, 2], [1, 2, 1]]])
    w = torch.from_numpy(k / k.sum()).float()
    return w[None, None].to(device)


# model/nerf.py:438-447 (sdf_tv=0 branch of density_total_variation)
def smooth_grad_tv(m, smooth_grad_tv_w):
    G = neus_sdf_gradient(m).permute(1, 0, 2, 3, 4)
    w = tv_smooth_conv_weight(G.device)
    conv = F.conv3d(F.pad(G, (1, 1, 1, 1, 1, 1), mode="replicate"), w)
    err = conv.detach() - G
    if m.nonempty_mask is not None:
        err = err[m.nonempty_mask.repeat(3, 1, 1, 1, 1)] ** 2
    else:
        err = err ** 2
    return err.mean() * smooth_grad_tv_w


# model/nerf_training.py:307-345 + model/nerf.py:469-478
def fine_loss(m, out, target, train_cfg, global_step, with_tv_loss=True):
    loss = train_cfg["weight_main"] * F.mse_loss(out['rgb_marched'], target)
    if train_cfg["weight_entropy_last"] > 0:
        pout = out['alphainv_cum'][..., -1].clamp(1e-6, 1 - 1e-6)
        entropy_last_loss = -(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean()
        loss = loss + train_cfg["weight_entropy_last"] * entropy_last_loss
    if train_cfg["weight_orientation"] > 0:
        zero = torch.tensor(0.0, dtype=torch.float32, device=target.device)
        w = out['weights'].detach()
        n = out['normal']
        v = -out['viewdirs']
        n_dot_v = (n * v).sum(dim=-1)
        ori = torch.mean((w * torch.fmin(zero, n_dot_v) ** 2).sum(dim=-1))
        loss = loss + train_cfg["weight_orientation"] * ori
    if train_cfg["sigmoid_rgb_loss"] > 0:
        loss = loss + train_cfg["sigmoid_rgb_loss"] * F.mse_loss(out['sigmoid_rgb'], target)
    if with_tv_loss and is_tv_step(train_cfg, global_step) and train_cfg["weight_tv_density"] > 0 \
            and train_cfg["smooth_grad_tv"] > 0:
        loss = loss + train_cfg["weight_tv_density"] * smooth_grad_tv(m, train_cfg["smooth_grad_tv"])
    return loss


def is_tv_step(train_cfg, global_step):
    return train_cfg["tv_from"] < global_step < train_cfg["tv_end"] and global_step % train_cfg["tv_every"] == 0


class RefOptimizer:
    """MaskedAdam (model/adam.py:167-221) with the param groups of
    create_optimizer_or_freeze_model (model/nerf_training.py:9-37): k0 masked, everything else
    dense, betas (0.9, 0.99), eps 1e-8."""

    def __init__(self, m, train_cfg, native):
        self.native = native
        self.state = {}
        self.groups = []
        lr = {"sdf": train_cfg["lrate_sdf"], "k0": train_cfg["lrate_k0"],
              "rgbnet": train_cfg["lrate_rgbnet"], "refnet": train_cfg["lrate_refnet"]}
        for name, p in m.params():
            key = name.split(".")[0]
            self.groups.append(dict(name=name, param=p, lr=lr[key], skip_zero_grad=(key == "k0")))

    def step(self):
        for g in self.groups:
            p = g["param"]
            if p.grad is None:
                continue
            st = self.state.setdefault(g["name"], dict(step=0, m=torch.zeros_like(p.data), v=torch.zeros_like(p.data)))
            st["step"] += 1
            self.native.adam("masked" if g["skip_zero_grad"] else "dense", p, p.grad, st["m"], st["v"],
                             st["step"], 0.9, 0.99, g["lr"], 1e-8)


def train_step(m, native, opt, rays_o, rays_d, viewdirs, target, global_step, train_cfg, near, stepsize, bg,
               stages=None, render_depth=False, render_grad=False):
    """One iteration of model/nerf_training.py:300-373 (fine stage)."""
    out = forward_fine(m, native, rays_o, rays_d, viewdirs, global_step, near, 1e9, stepsize, bg, stages=stages,
                       render_depth=render_depth, render_grad=render_grad)
    for _, p in m.params():
        p.grad = None                                           # zero_grad(set_to_none=True)
    loss = fine_loss(m, out, target, train_cfg, global_step)
    loss.backward()
    if is_tv_step(train_cfg, global_step) and train_cfg["weight_tv_density"] > 0 and train_cfg["sdf_tv"] > 0:
        # nerf_training.py:353-360 + nerf.py:465-467
        w = train_cfg["weight_tv_density"] * train_cfg["sdf_tv"] / len(rays_o)
        w = float(w * m.world_size.max() / 128)
        native.tv_add_grad(m.sdf, m.sdf.grad, w, w, w, global_step < train_cfg["tv_dense_before"])
    if opt is not None:
        opt.step()
    return out, loss
