#!/usr/bin/env python
"""Golden fixture for the fine-stage step, produced by THE REFERENCE'S OWN PYTHON
(/root/reference/model/nerf.py: class nerf, forward_fine; model/grid.py DenseGrid;
model/nerf.py MaskCache / Alphas2Weights) imported in this container on the CPU.

What is stubbed (none of it is arithmetic on the path):
  * third-party imports that are not installed and unused by forward_fine: mcubes, trimesh,
    imageio, lpips, cv2-free evaluation helpers ...                       -> empty modules
  * torch_scatter.segment_coo(src, index, out, reduce='sum')             -> out.index_add_(0, index, src)
    (torch_scatter 2.1.2 semantics: sorted-index segment sum)
  * torch.utils.cpp_extension.load(...) (the reference JIT-builds CUDA-only kernels)
      render_utils_cuda.{sample_pts_on_rays, alpha2weight, alpha2weight_backward}
                                                                          -> oracle/native_ref (C)
    the C restatement is itself pinned bit-exact against the reference kernels run on a B200
    (tests/golden/native_golden.npz)
  * hard-coded .cuda() / .to('cuda')                                      -> no-ops (CPU run)

Output: tests/golden/fine_golden.npz  (inputs are regenerated from seeds by fgs_nerf_b200.scene).
Run here:  python tests/golden/make_fine_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def install_stubs():
    from oracle import native_ref

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def segment_coo(src, index, out=None, reduce="sum"):
        assert reduce == "sum"
        return out.index_add_(0, index, src)

    stub("torch_scatter", segment_coo=segment_coo)
    for n in ("mcubes", "trimesh", "imageio", "lpips", "cv2", "icecream", "torch_efficient_distloss", "mmcv"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                stub(n)

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    class RU:
        @staticmethod
        def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
            return [t(a) for a in native_ref.sample_pts_on_rays(
                rays_o.numpy(), rays_d.numpy(), xyz_min.numpy(), xyz_max.numpy(), float(near), float(far), float(stepdist))]

        @staticmethod
        def alpha2weight(alpha, ray_id, n):
            return [t(a) for a in native_ref.alpha2weight(alpha.detach().numpy(), ray_id.numpy(), n)]

        @staticmethod
        def alpha2weight_backward(alpha, w, T, last, i_s, i_e, n, gw, gl):
            return t(native_ref.alpha2weight_backward(alpha.detach().numpy(), w.numpy(), T.numpy(), last.numpy(),
                                                      i_s.numpy(), i_e.numpy(), n, gw.contiguous().numpy(),
                                                      gl.contiguous().numpy()))

        @staticmethod
        def maskcache_lookup(world, xyz, scale, shift):
            return t(native_ref.maskcache_lookup(world.numpy(), xyz.numpy(), scale.numpy(), shift.numpy()))

    class TV:
        @staticmethod
        def total_variation_add_grad(param, grad, wx, wy, wz, dense):
            native_ref.total_variation_add_grad(param.detach().numpy(), grad.numpy(), wx, wy, wz, dense)

    class AD:
        @staticmethod
        def adam_upd(p, g, m, v, step, b1, b2, lr, eps):
            native_ref.adam_upd(p.detach().numpy().reshape(-1), g.contiguous().numpy().reshape(-1),
                                m.numpy().reshape(-1), v.numpy().reshape(-1), step, b1, b2, lr, eps)

        @staticmethod
        def masked_adam_upd(p, g, m, v, step, b1, b2, lr, eps):
            native_ref.masked_adam_upd(p.detach().numpy().reshape(-1), g.contiguous().numpy().reshape(-1),
                                       m.numpy().reshape(-1), v.numpy().reshape(-1), step, b1, b2, lr, eps)

    import torch.utils.cpp_extension as ce

    def fake_load(name, sources=None, **kw):
        return {"render_utils_cuda": RU, "total_variation_cuda": TV, "adam_upd_cuda": AD}[name]

    ce.load = fake_load
    import math
    if not hasattr(np, "math"):
        np.math = math   # numpy>=2 dropped np.math (model/utils.py:195 uses np.math.factorial)
    # hard-coded device moves -> CPU no-ops
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return _to(self, *a, **k)

    torch.Tensor.to = to


def build_reference_model(sc, tmpdir):
    """Instantiate the reference's `nerf` (fine stage) and load the synthetic scene into it."""
    from model import nerf as ref_nerf   # noqa  (the real reference module)
    ckpt = os.path.join(tmpdir, "geometry_searching_last.tar")
    torch.save({"MaskCache_kwargs": {"xyz_min": sc.mask_xyz_min.tolist(), "xyz_max": sc.mask_xyz_max.tolist()},
                "model_state_dict": {"sdf_mask.grid": sc.pre_pool_mask}}, ckpt)
    c = sc.cfg
    G = sc.G
    model = ref_nerf.nerf(
        xyz_min=sc.xyz_min.numpy(), xyz_max=sc.xyz_max.numpy(), num_voxels=G ** 3, num_voxels_base=G ** 3,
        mask_cache_path=ckpt, mask_cache_thres=c["mask_cache_thres"], fast_color_thres=c["fast_color_thres"],
        k0_dim=c["k0_dim"], rgbnet_depth=c["rgbnet_depth"], rgbnet_width=c["rgbnet_width"], ref=True,
        refnet_width=c["refnet_width"], refnet_depth=c["refnet_depth"], posbase_pe=c["posbase_pe"],
        viewbase_pe=c["viewbase_pe"], refbase_pe=c["refbase_pe"], grad_feat=c["grad_feat"], sdf_feat=c["sdf_feat"],
        s_ratio=c["s_ratio"], s_start=c["s_start"], stage="fine", use_viewdir=True, center_sdf=True)
    assert tuple(model.world_size.tolist()) == (G, G, G), model.world_size
    with torch.no_grad():
        model.sdf.grid.data = sc.sdf.clone()
        model.k0.grid.data = sc.k0.clone()
        lin = [m for m in model.rgbnet.modules() if isinstance(m, torch.nn.Linear)]
        for l, (w, b) in zip(lin, sc.rgbnet):
            l.weight.data, l.bias.data = w.clone(), b.clone()
        lin = [m for m in model.refnet.modules() if isinstance(m, torch.nn.Linear)]
        for l, (w, b) in zip(lin, sc.refnet):
            l.weight.data, l.bias.data = w.clone(), b.clone()
    model._set_nonempty_mask()
    return model, ref_nerf


def main():
    import tempfile
    install_stubs()
    sys.path.insert(0, REF)
    from fgs_nerf_b200 import scene
    G, N, step = 48, 192, 5000
    sc = scene.make_scene(G=G, mask_G=40)
    # the MaskCache max-pools the checkpointed grid itself (nerf.py:1199): give it the un-pooled one
    mlin = torch.linspace(-1.2, 1.2, 40)
    mx, my, mz = torch.meshgrid(mlin, mlin, mlin, indexing="ij")
    sc.pre_pool_mask = ((((mx * mx + my * my + mz * mz).sqrt() - 0.6).abs() < 0.15).float() * 1e-3)[None, None]
    ro, rd, vd, tgt = scene.make_rays(N=N, H=100, W=100, n_views=8)
    with tempfile.TemporaryDirectory() as td:
        model, ref_nerf = build_reference_model(sc, td)
    assert torch.equal(model.mask_cache.sdf_mask, sc.mask_grid), "scene.mask_grid != MaskCache pooling"
    tc = scene.FINE_TRAIN
    render_kwargs = dict(near=2.0, far=6.0, stepsize=0.5, bg=1.0, render_depth=True, render_grad=True)
    out = model(ro, rd, vd, global_step=step, **render_kwargs)
    # loss block of model/nerf_training.py:307-345 (verbatim semantics), then backward
    import torch.nn.functional as F
    loss = tc["weight_main"] * F.mse_loss(out["rgb_marched"], tgt)
    pout = out["alphainv_cum"][..., -1].clamp(1e-6, 1 - 1e-6)
    loss = loss + tc["weight_entropy_last"] * (-(pout * torch.log(pout) + (1 - pout) * torch.log(1 - pout)).mean())
    loss = loss + tc["weight_orientation"] * model.orientation_loss(out)
    loss = loss + tc["sigmoid_rgb_loss"] * F.mse_loss(out["sigmoid_rgb"], tgt)
    tv_loss = tc["weight_tv_density"] * model.density_total_variation(sdf_tv=0, smooth_grad_tv=tc["smooth_grad_tv"])
    (loss + tv_loss).backward()
    grads = {"sdf": model.sdf.grid.grad.clone(), "k0": model.k0.grid.grad.clone()}
    for name, net in (("rgbnet", model.rgbnet), ("refnet", model.refnet)):
        for i, l in enumerate(m for m in net.modules() if isinstance(m, torch.nn.Linear)):
            grads[f"{name}.{i}.weight"], grads[f"{name}.{i}.bias"] = l.weight.grad.clone(), l.bias.grad.clone()
    # TV add-grad exactly as nerf_training.py:353-360 / nerf.py:465-467
    sdf_grad_tv = model.sdf.grid.grad.clone()
    model.sdf.grid.grad = sdf_grad_tv
    model.sdf_total_variation_add_grad(tc["weight_tv_density"] * tc["sdf_tv"] / len(ro), step < tc["tv_dense_before"])
    keep = ["alphainv_cum", "weights", "ray_id", "rgb_marched", "sigmoid_rgb", "normal_marched", "normal",
            "raw_alpha", "raw_rgb", "depth", "mask", "mask_outbbox", "gradient"]
    fix = {"meta_G": G, "meta_N": N, "meta_mask_G": 40, "meta_global_step": step, "meta_H": 100, "meta_W": 100,
           "meta_n_views": 8, "s_val": float(out["s_val"]), "loss": float(loss.detach()), "tv_loss": float(tv_loss.detach()),
           "sdf_grad_after_tv": sdf_grad_tv.numpy(), "nonempty_mask": model.nonempty_mask.numpy()}
    for k in keep:
        fix["out_" + k] = out[k].detach().numpy()
    for k, g in grads.items():
        if k == "k0":   # sparse: only touched voxels are non-zero
            nz = torch.nonzero(g.reshape(-1)).reshape(-1)
            fix["grad_k0_idx"], fix["grad_k0_val"] = nz.numpy(), g.reshape(-1)[nz].numpy()
        else:
            fix["grad_" + k] = g.numpy()
    path = os.path.join(ROOT, "tests", "golden", "fine_golden.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, "M2 =", out["weights"].numel(), "loss =", float(loss.detach()), "tv =", float(tv_loss.detach()),
          "size MB =", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    main()
