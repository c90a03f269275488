#!/usr/bin/env python
"""Golden fixture for the GEOMETRY_SEARCHING stage (the first stage of a run: forward_coarse without a MaskCache, with the
Incremental-Voxel-Training box `inc_mask`, k0_dim 6, refnet 72 -> 128 -> 128 -> 3), produced by THE
REFERENCE'S OWN PYTHON (/root/reference/model/nerf.py: class nerf with stage='geometry_searching') imported in this
container on the CPU with the stubs of make_fine_golden.py.

Model settings: config/shiny_blender.py `geometry_searching_model`; loss block of model/nerf_training.py:307-345 with
`geometry_searching`'s weights (weight_rgbper, ori_tv; the stage has no non-empty mask, so total_variation divides by
sdf.sum(), model/nerf.py:1212-1221).

    python tests/golden/make_geometry_golden.py     # needs /root/reference; writes geometry_golden.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"

GEOMETRY_MODEL = dict(stage="geometry_searching", k0_dim=6, refnet_width=128, refnet_depth=3, posbase_pe=5, viewbase_pe=1,
                      refbase_pe=3, smooth_ksize=5, smooth_sigma=0.8, s_ratio=50, s_start=0.2, fast_color_thres=1e-4,
                      mask_cache_thres=1e-3, stepsize=0.5)
GEOMETRY_TRAIN = dict(weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=0.2, weight_tv_density=0.01,
                      sigmoid_rgb_loss=0.1, weight_orientation=1e-4, sdf_tv=0.1, smooth_grad_tv=0.05)
INC_LOWER, INC_UPPER = (0.12, 0.2, 0.08), (0.9, 0.83, 0.88)       # a voxel_inc box part of the way through its schedule


def geometry_refnet(seed=791, cfg=GEOMETRY_MODEL):
    import math
    dim = (3 + 3 * cfg["refbase_pe"] * 2) + cfg["k0_dim"] + (3 + 3 * cfg["posbase_pe"] * 2) + 3 + (3 + 3 * cfg["viewbase_pe"] * 2)
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        b = 1.0 / math.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b).float(), ((torch.rand(o, generator=g) * 2 - 1) * b).float()

    W = cfg["refnet_width"]
    return [lin(dim, W), lin(W, W), lin(W, 3)]


def main():
    import make_fine_golden as mk
    mk.install_stubs()
    sys.path.insert(0, REF)
    from fgs_nerf_b200 import scene
    from model import nerf as ref_nerf
    import torch.nn.functional as F
    G, N, step = 40, 512, 600
    c = GEOMETRY_MODEL
    sc = scene.make_scene(G=G, mask_G=32, with_mlps=False, cfg=dict(scene.FINE_MODEL, k0_dim=c["k0_dim"]))
    ro, rd, vd, tgt = scene.make_rays(N=N, H=100, W=100, n_views=8)
    model = ref_nerf.nerf(
        xyz_min=sc.xyz_min.numpy(), xyz_max=sc.xyz_max.numpy(), num_voxels=G ** 3, num_voxels_base=G ** 3,
        mask_cache_path=None, mask_cache_thres=c["mask_cache_thres"], fast_color_thres=c["fast_color_thres"],
        k0_dim=c["k0_dim"], ref=True, refnet_width=c["refnet_width"], refnet_depth=c["refnet_depth"],
        posbase_pe=c["posbase_pe"], viewbase_pe=c["viewbase_pe"], refbase_pe=c["refbase_pe"],
        smooth_ksize=c["smooth_ksize"], smooth_sigma=c["smooth_sigma"], s_ratio=c["s_ratio"], s_start=c["s_start"],
        stage="geometry_searching", use_viewdir=True)
    assert tuple(model.world_size.tolist()) == (G, G, G) and model.mask_cache is None and model.nonempty_mask is None
    assert model.refnet_dim == 72 and model.rgbnet is None
    refnet = geometry_refnet()
    with torch.no_grad():
        model.sdf.grid.data = sc.sdf.clone()
        model.k0.grid.data = sc.k0.clone()
        lin = [m for m in model.refnet.modules() if isinstance(m, torch.nn.Linear)]
        for l, (w, b) in zip(lin, refnet):
            assert l.weight.shape == w.shape
            l.weight.data, l.bias.data = w.clone(), b.clone()
    model.set_inc_mask(torch.tensor(INC_LOWER), torch.tensor(INC_UPPER))
    out = model(ro, rd, vd, global_step=step, near=2.0, far=6.0, stepsize=c["stepsize"], bg=1.0, render_depth=True, render_grad=True)
    t = GEOMETRY_TRAIN
    loss = t["weight_main"] * F.mse_loss(out["rgb_marched"], tgt)
    rgbper = (out["raw_rgb"] - tgt[out["ray_id"]]).pow(2).sum(-1)
    loss = loss + t["weight_rgbper"] * (rgbper * out["weights"].detach()).sum() / len(ro)
    pout = out["alphainv_cum"][..., -1].clamp(1e-6, 1 - 1e-6)
    loss = loss + t["weight_entropy_last"] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    loss = loss + t["weight_orientation"] * model.orientation_loss(out)
    loss = loss + t["sigmoid_rgb_loss"] * F.mse_loss(out["sigmoid_rgb"], tgt)
    tv_smooth = t["weight_tv_density"] * model.density_total_variation(sdf_tv=0, smooth_grad_tv=t["smooth_grad_tv"])
    tv_sdf = t["weight_tv_density"] * model.density_total_variation(sdf_tv=t["sdf_tv"], smooth_grad_tv=0)      # ori_tv
    (loss + tv_smooth + tv_sdf).backward()
    fix = {"meta_G": G, "meta_N": N, "meta_mask_G": 32, "meta_global_step": step, "s_val": float(out["s_val"]),
           "loss": float(loss.detach()), "tv_smooth": float(tv_smooth.detach()), "tv_sdf": float(tv_sdf.detach()),
           "inc_mask": model.inc_mask.mask.numpy(), "grad_sdf": model.sdf.grid.grad.numpy()}
    for k in ("alphainv_cum", "weights", "ray_id", "rgb_marched", "sigmoid_rgb", "normal_marched", "normal", "raw_alpha",
              "raw_rgb", "depth", "mask", "mask_outbbox", "gradient"):
        fix["out_" + k] = out[k].detach().numpy()
    g = model.k0.grid.grad
    nz = torch.nonzero(g.reshape(-1)).reshape(-1)
    fix["grad_k0_idx"], fix["grad_k0_val"] = nz.numpy(), g.reshape(-1)[nz].numpy()
    for i, l in enumerate(m for m in model.refnet.modules() if isinstance(m, torch.nn.Linear)):
        fix[f"grad_refnet.{i}.weight"], fix[f"grad_refnet.{i}.bias"] = l.weight.grad.numpy(), l.bias.grad.numpy()
    path = os.path.join(HERE, "geometry_golden.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, "samples", out["weights"].numel(), "alpha samples", out["raw_alpha"].numel(), "loss", float(loss.detach()),
          "tv", float(tv_smooth.detach()), float(tv_sdf.detach()), "MB", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    main()
