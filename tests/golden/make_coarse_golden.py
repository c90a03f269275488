#!/usr/bin/env python
"""Golden fixture for the COARSE-stage forward + loss + backward, produced by THE REFERENCE'S OWN
PYTHON (/root/reference/model/nerf.py: class nerf with stage='coarse' -> forward_coarse, :943-1075) imported in this
container on the CPU, with the stubs of make_fine_golden.py (nothing arithmetic is stubbed except the three CUDA-only
ops, which run the C oracle that is itself pinned against the reference's kernels).

Model settings: config/shiny_blender.py `coarse_model` (refnet 3 x 192, PE 5 / 1 / 5, smooth_ksize 5, sigma 0.8,
s_ratio 50, s_start 0.2); loss block of model/nerf_training.py:307-345 with `coarse_train`'s weights (incl.
weight_rgbper and ori_tv: density_total_variation(sdf_tv) through autograd).

    python tests/golden/make_coarse_golden.py       # needs /root/reference; writes coarse_golden.npz
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"

COARSE_MODEL = dict(k0_dim=12, refnet_width=192, refnet_depth=3, posbase_pe=5, viewbase_pe=1, refbase_pe=5,
                    smooth_ksize=5, smooth_sigma=0.8, s_ratio=50, s_start=0.2, fast_color_thres=1e-4,
                    mask_cache_thres=1e-3, stepsize=0.5)
COARSE_TRAIN = dict(weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=0.2, weight_tv_density=0.01,
                    sigmoid_rgb_loss=0.1, weight_orientation=1e-4, sdf_tv=0.1, smooth_grad_tv=0.05)


def coarse_refnet(seed=790, cfg=COARSE_MODEL):
    """nn.Linear default init under a fixed seed: 90 -> 192 -> 192 -> 3."""
    import math
    dim = (3 + 3 * cfg["refbase_pe"] * 2) + cfg["k0_dim"] + (3 + 3 * cfg["posbase_pe"] * 2) + 3 + (3 + 3 * cfg["viewbase_pe"] * 2)
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        b = 1.0 / math.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b).float(), ((torch.rand(o, generator=g) * 2 - 1) * b).float()

    W = cfg["refnet_width"]
    return [lin(dim, W)] + [lin(W, W) for _ in range(cfg["refnet_depth"] - 2)] + [lin(W, 3)]


def main():
    import make_fine_golden as mk
    mk.install_stubs()
    sys.path.insert(0, REF)
    from fgs_nerf_b200 import scene
    from model import nerf as ref_nerf
    from model import grid as ref_grid  # noqa: F401
    G, N, step = 40, 512, 3000
    sc = scene.make_scene(G=G, mask_G=32, with_mlps=False)
    mlin = torch.linspace(-1.2, 1.2, 32)
    mx, my, mz = torch.meshgrid(mlin, mlin, mlin, indexing="ij")
    pre_pool = ((((mx * mx + my * my + mz * mz).sqrt() - 0.6).abs() < 0.15).float() * 1e-3)[None, None]
    ro, rd, vd, tgt = scene.make_rays(N=N, H=100, W=100, n_views=8)
    c = COARSE_MODEL
    with tempfile.TemporaryDirectory() as td:
        ckpt = os.path.join(td, "geometry_searching_last.tar")
        torch.save({"MaskCache_kwargs": {"xyz_min": sc.mask_xyz_min.tolist(), "xyz_max": sc.mask_xyz_max.tolist()},
                    "model_state_dict": {"sdf_mask.grid": pre_pool}}, ckpt)
        model = ref_nerf.nerf(
            xyz_min=sc.xyz_min.numpy(), xyz_max=sc.xyz_max.numpy(), num_voxels=G ** 3, num_voxels_base=G ** 3,
            mask_cache_path=ckpt, mask_cache_thres=c["mask_cache_thres"], fast_color_thres=c["fast_color_thres"],
            k0_dim=c["k0_dim"], ref=True, refnet_width=c["refnet_width"], refnet_depth=c["refnet_depth"],
            posbase_pe=c["posbase_pe"], viewbase_pe=c["viewbase_pe"], refbase_pe=c["refbase_pe"],
            smooth_ksize=c["smooth_ksize"], smooth_sigma=c["smooth_sigma"], s_ratio=c["s_ratio"], s_start=c["s_start"],
            stage="coarse", use_viewdir=True)
    assert tuple(model.world_size.tolist()) == (G, G, G), model.world_size
    assert model.refnet_dim == 90 and model.rgbnet is None and model.smooth_sdf
    refnet = coarse_refnet()
    with torch.no_grad():
        model.sdf.grid.data = sc.sdf.clone()
        model.k0.grid.data = sc.k0.clone()
        lin = [m for m in model.refnet.modules() if isinstance(m, torch.nn.Linear)]
        assert len(lin) == len(refnet)
        for l, (w, b) in zip(lin, refnet):
            assert l.weight.shape == w.shape
            l.weight.data, l.bias.data = w.clone(), b.clone()
    model._set_nonempty_mask()          # coarse stage: also writes sdf = 1 outside the mask (nerf.py:352-353)
    model.unset_inc_mask()              # nerf_training.py:295-296 (voxel_inc False)
    sdf_after_mask = model.sdf.grid.detach().clone()
    render_kwargs = dict(near=2.0, far=6.0, stepsize=c["stepsize"], bg=1.0, render_depth=True, render_grad=True)
    out = model(ro, rd, vd, global_step=step, **render_kwargs)
    import torch.nn.functional as F
    t = COARSE_TRAIN
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
           "sdf_after_mask": sdf_after_mask.numpy(), "nonempty_mask": model.nonempty_mask.numpy(),
           "smooth_weight": model.smooth_conv.weight.data[0, 0].numpy(),
           "grad_sdf": model.sdf.grid.grad.numpy()}
    for k in ("alphainv_cum", "weights", "ray_id", "rgb_marched", "sigmoid_rgb", "normal_marched", "normal", "raw_alpha",
              "raw_rgb", "depth", "mask", "mask_outbbox", "gradient"):
        fix["out_" + k] = out[k].detach().numpy()
    g = model.k0.grid.grad
    nz = torch.nonzero(g.reshape(-1)).reshape(-1)
    fix["grad_k0_idx"], fix["grad_k0_val"] = nz.numpy(), g.reshape(-1)[nz].numpy()
    for i, l in enumerate(m for m in model.refnet.modules() if isinstance(m, torch.nn.Linear)):
        fix[f"grad_refnet.{i}.weight"], fix[f"grad_refnet.{i}.bias"] = l.weight.grad.numpy(), l.bias.grad.numpy()
    path = os.path.join(HERE, "coarse_golden.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, "samples", out["weights"].numel(), "alpha samples", out["raw_alpha"].numel(), "loss", float(loss.detach()),
          "tv", float(tv_smooth.detach()), float(tv_sdf.detach()), "MB", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    main()
