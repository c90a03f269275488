"""CPU tests (no GPU): the torch restatement of the COARSE-stage forward + loss (oracle/coarse_ref.py, SURVEY.md 8f
row 1) against the fixture produced by the reference's OWN model/nerf.py with stage='coarse'
(tests/golden/coarse_golden.npz, made by tests/golden/make_coarse_golden.py).  No kernel exists for this row yet: this
pins the checker the kernels will be held to."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from oracle import coarse_ref as Cr  # noqa: E402
from oracle import nerf_ref as R  # noqa: E402


@pytest.fixture(scope="module")
def run():
    import make_coarse_golden as mk            # the model / loss settings the fixture was made with (no reference import)
    from fgs_nerf_b200 import scene
    gold = np.load(os.path.join(HERE, "golden", "coarse_golden.npz"))
    G, N, step = int(gold["meta_G"]), int(gold["meta_N"]), int(gold["meta_global_step"])
    sc = scene.make_scene(G=G, mask_G=int(gold["meta_mask_G"]), with_mlps=False)
    ro, rd, vd, tgt = scene.make_rays(N=N, H=100, W=100, n_views=8)
    m = Cr.CoarseModel(sc, mk.coarse_refnet(), mk.COARSE_MODEL)
    m.set_nonempty_mask()
    out = Cr.forward_coarse(m, R.CpuNative(), ro, rd, vd, step, near=2.0, stepsize=0.5, bg=1.0)
    loss, tv_smooth, tv_sdf = Cr.coarse_loss(m, out, tgt, mk.COARSE_TRAIN)
    (loss + tv_smooth + tv_sdf).backward()
    return gold, m, out, (loss, tv_smooth, tv_sdf)


def test_coarse_setup_matches_reference(run):
    gold, m, out, _ = run
    assert np.array_equal(m.nonempty_mask.numpy(), gold["nonempty_mask"])
    assert np.array_equal(m.sdf.detach().numpy(), gold["sdf_after_mask"])          # sdf = 1 outside the mask
    np.testing.assert_allclose(m.smooth_weight[0, 0].numpy(), gold["smooth_weight"], rtol=1e-6, atol=0)


def test_coarse_forward_matches_reference_python(run):
    gold, m, out, _ = run
    assert np.array_equal(out["ray_id"].numpy(), gold["out_ray_id"])
    assert np.array_equal(out["mask"].numpy(), gold["out_mask"])
    assert np.array_equal(out["mask_outbbox"].numpy(), gold["out_mask_outbbox"])
    for k in ("weights", "raw_alpha", "alphainv_cum", "gradient", "normal", "raw_rgb", "rgb_marched", "sigmoid_rgb",
              "normal_marched", "depth"):
        np.testing.assert_allclose(out[k].detach().numpy(), gold["out_" + k], rtol=1e-5, atol=1e-7, err_msg=k)
    assert abs(out["s_val"] - float(gold["s_val"])) < 1e-12
    assert out["weights"].numel() > 500


def test_coarse_loss_and_grads_match_reference_python(run):
    gold, m, out, (loss, tv_smooth, tv_sdf) = run
    assert abs(float(loss.detach()) - float(gold["loss"])) <= 1e-6 * abs(float(gold["loss"]))
    assert abs(float(tv_smooth.detach()) - float(gold["tv_smooth"])) <= 1e-6 * abs(float(gold["tv_smooth"]))
    assert abs(float(tv_sdf.detach()) - float(gold["tv_sdf"])) <= 1e-6 * abs(float(gold["tv_sdf"]))
    got = dict(m.params())
    for i in range(3):
        for kind in ("weight", "bias"):
            name = f"refnet.{i}.{kind}"
            np.testing.assert_allclose(got[name].grad.numpy(), gold["grad_" + name], rtol=1e-4, atol=1e-9, err_msg=name)
    dense = np.zeros(int(np.prod(got["k0"].shape)), dtype=np.float32)
    dense[gold["grad_k0_idx"]] = gold["grad_k0_val"]
    np.testing.assert_allclose(got["k0"].grad.numpy().reshape(-1), dense, rtol=1e-4, atol=1e-10)
    # through the Gaussian smoother, the gradient grid and the autograd SDF TV
    np.testing.assert_allclose(got["sdf"].grad.numpy(), gold["grad_sdf"], rtol=1e-4, atol=1e-9)


# ---- the geometry_searching stage (no MaskCache, inc_mask, k0_dim 6): the same restatement, pinned on its own fixture ----
@pytest.fixture(scope="module")
def run_geometry():
    import make_geometry_golden as mg
    from fgs_nerf_b200 import scene
    gold = np.load(os.path.join(HERE, "golden", "geometry_golden.npz"))
    G, N, step = int(gold["meta_G"]), int(gold["meta_N"]), int(gold["meta_global_step"])
    sc = scene.make_scene(G=G, mask_G=int(gold["meta_mask_G"]), with_mlps=False, cfg=dict(scene.FINE_MODEL, k0_dim=6))
    ro, rd, vd, tgt = scene.make_rays(N=N, H=100, W=100, n_views=8)
    m = Cr.CoarseModel(sc, mg.geometry_refnet(), mg.GEOMETRY_MODEL)
    assert m.set_nonempty_mask() is None and m.nonempty_mask is None
    m.set_inc_mask(mg.INC_LOWER, mg.INC_UPPER)
    out = Cr.forward_coarse(m, R.CpuNative(), ro, rd, vd, step, near=2.0, stepsize=0.5, bg=1.0)
    loss, tv_smooth, tv_sdf = Cr.coarse_loss(m, out, tgt, mg.GEOMETRY_TRAIN)
    (loss + tv_smooth + tv_sdf).backward()
    return gold, m, out, (loss, tv_smooth, tv_sdf)


def test_geometry_forward_matches_reference_python(run_geometry):
    gold, m, out, _ = run_geometry
    assert np.array_equal(m.inc_mask[0].numpy(), gold["inc_mask"])
    assert np.array_equal(out["ray_id"].numpy(), gold["out_ray_id"])
    assert np.array_equal(out["mask"].numpy(), gold["out_mask"])
    assert np.array_equal(out["mask_outbbox"].numpy(), gold["out_mask_outbbox"])
    for k in ("weights", "raw_alpha", "alphainv_cum", "gradient", "normal", "raw_rgb", "rgb_marched", "sigmoid_rgb",
              "normal_marched", "depth"):
        np.testing.assert_allclose(out[k].detach().numpy(), gold["out_" + k], rtol=1e-5, atol=1e-7, err_msg=k)
    assert out["weights"].numel() > 500


def test_geometry_loss_and_grads_match_reference_python(run_geometry):
    gold, m, out, (loss, tv_smooth, tv_sdf) = run_geometry
    assert abs(float(loss.detach()) - float(gold["loss"])) <= 1e-6 * abs(float(gold["loss"]))
    assert abs(float(tv_smooth.detach()) - float(gold["tv_smooth"])) <= 1e-6 * abs(float(gold["tv_smooth"]))
    assert abs(float(tv_sdf.detach()) - float(gold["tv_sdf"])) <= 1e-6 * abs(float(gold["tv_sdf"]))     # divides by sdf.sum()
    got = dict(m.params())
    for i in range(3):
        for kind in ("weight", "bias"):
            name = f"refnet.{i}.{kind}"
            np.testing.assert_allclose(got[name].grad.numpy(), gold["grad_" + name], rtol=1e-4, atol=1e-9, err_msg=name)
    dense = np.zeros(int(np.prod(got["k0"].shape)), dtype=np.float32)
    dense[gold["grad_k0_idx"]] = gold["grad_k0_val"]
    np.testing.assert_allclose(got["k0"].grad.numpy().reshape(-1), dense, rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(got["sdf"].grad.numpy(), gold["grad_sdf"], rtol=1e-4, atol=1e-9)
