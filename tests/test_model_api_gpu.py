"""GPU test of the reference's model-class API on top of the fused engine (fgs_nerf_b200.nerf.nerf; SURVEY.md 8b
"Python-class API to keep"): constructor keywords, `model(rays_o, rays_d, viewdirs, global_step=..., **render_kwargs)`
with the reference's result keys, `load_state_dict` / `state_dict` in the reference's key layout -- against the oracle's
forward_fine (oracle/nerf_ref.py + the reference's compiled kernels).
"""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import nerf_ref as R  # noqa: E402
from oracle import ref_ext  # noqa: E402

pytestmark = pytest.mark.gpu
LIN = ("0", "2.0", "3.0", "4")
REF_KEYS = ["alphainv_cum", "weights", "ray_id", "viewdirs", "rgb_marched", "sigmoid_rgb", "normal_marched", "normal",
            "raw_alpha", "raw_rgb", "depth", "disp", "mask", "mask_outbbox", "gradient", "s_val"]     # model/nerf.py:922-939


def _build(cuda, sc, n_rays):
    from fgs_nerf_b200 import nerf as N
    mc = N.MaskCache(mask_cache_thres=sc.mask_thres, sdf_mask=sc.mask_grid, xyz_min=sc.mask_xyz_min, xyz_max=sc.mask_xyz_max)
    c = sc.cfg
    model = N.nerf(xyz_min=sc.xyz_min.numpy(), xyz_max=sc.xyz_max.numpy(), num_voxels=sc.G ** 3, num_voxels_base=sc.G ** 3,
                   mask_cache=mc, fast_color_thres=c["fast_color_thres"], k0_dim=c["k0_dim"], rgbnet_depth=4, rgbnet_width=256,
                   ref=True, refnet_width=256, refnet_depth=4, posbase_pe=5, viewbase_pe=3, refbase_pe=8,
                   grad_feat=c["grad_feat"], sdf_feat=c["sdf_feat"], s_ratio=c["s_ratio"], s_start=c["s_start"], stage="fine",
                   use_viewdir=True, n_rays=n_rays, device=cuda)
    sd = {"sdf.grid": sc.sdf, "k0.grid": sc.k0}
    for name, layers in (("rgbnet", sc.rgbnet), ("refnet", sc.refnet)):
        for idx, (w, b) in zip(LIN, layers):
            sd[f"{name}.{idx}.weight"], sd[f"{name}.{idx}.bias"] = w, b
    model.load_state_dict(sd, strict=False)
    return model


def test_model_class_forward_matches_oracle(cuda):
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    from fgs_nerf_b200 import scene
    n_rays = 192
    sc = scene.make_scene(G=48, mask_G=40)
    ro, rd, vd, _ = [t.to(cuda) for t in scene.make_rays(N=n_rays, H=100, W=100, n_views=8)]
    model = _build(cuda, sc, n_rays)
    assert model.world_size.tolist() == [48, 48, 48] and tuple(model.k0.grid.shape) == (1, 12, 48, 48, 48)
    m = R.FineModel(sc, device=cuda, requires_grad=False)
    rk = dict(near=2.0, far=6.0, stepsize=0.5, bg=1.0, render_depth=True, render_grad=True)
    for n in (n_rays, 100):                                   # a full batch and a padded one
        got = model(ro[:n], rd[:n], vd[:n], global_step=5000, **rk)
        with torch.no_grad():
            want = R.forward_fine(m, R.RefCudaNative(), ro[:n].contiguous(), rd[:n].contiguous(), vd[:n].contiguous(), 5000,
                                  2.0, 1e9, 0.5, 1.0, render_depth=True, render_grad=True)
        assert [k for k in REF_KEYS if k not in got] == []
        assert torch.equal(got["ray_id"], want["ray_id"]) and torch.equal(got["mask"], want["mask"])
        assert torch.equal(got["mask_outbbox"], want["mask_outbbox"])
        assert got["rgb_marched"].shape == (n, 3) and got["alphainv_cum"].shape == (n,)
        for k in ("rgb_marched", "sigmoid_rgb", "weights", "raw_alpha", "raw_rgb", "normal", "alphainv_cum", "depth"):
            a, b = got[k].float(), want[k].float()
            assert a.shape == b.shape, k
            assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max())), k
        assert abs(got["s_val"] - want["s_val"]) < 1e-12
    with pytest.raises(RuntimeError):
        model(ro, rd, vd, global_step=5000, **dict(rk, stepsize=0.25))


def test_model_class_state_dict_and_refusals(cuda):
    from fgs_nerf_b200 import nerf as N, scene
    sc = scene.make_scene(G=24, mask_G=20)
    model = _build(cuda, sc, 64)
    layout = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_layout.json")))
    want = [(k, s) for k, s, _ in layout["model_state_dict"] if not k.startswith(("grad_conv.", "tv_smooth_conv."))]
    sd = model.state_dict()
    assert [(k, list(v.shape)) for k, v in sd.items() if k != "sdf_mask.grid"] == want
    assert torch.equal(sd["k0.grid"].cpu(), sc.k0) and torch.equal(sd["refnet.4.weight"].cpu(), sc.refnet[3][0])
    assert list(model.get_kwargs().keys()) == list(layout["model_kwargs"].keys())
    mc = N.MaskCache(mask_cache_thres=sc.mask_thres, sdf_mask=sc.mask_grid, xyz_min=sc.mask_xyz_min, xyz_max=sc.mask_xyz_max)
    for bad in (dict(stage="coarse"), dict(refnet_width=192), dict(smooth_ksize=5), dict(k0_dim=6)):
        kw = dict(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=24 ** 3, mask_cache=mc, ref=True,
                  grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0), stage="fine", n_rays=64, device=cuda)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            N.nerf(**kw)
