"""CPU tests (no GPU): the torch restatement oracle/nerf_ref.py against the fixture produced by
the reference's OWN model/nerf.py (tests/golden/fine_golden.npz, made by make_fine_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fine_case  # noqa: E402
from oracle import nerf_ref as R  # noqa: E402


@pytest.fixture(scope="module")
def run():
    from fgs_nerf_b200 import scene
    gold = fine_case.load_golden()
    sc, (ro, rd, vd, tgt), step = fine_case.golden_inputs(gold)
    m = R.FineModel(sc)
    m.set_nonempty_mask()
    native = R.CpuNative()
    stages = {}
    out, loss = R.train_step(m, native, None, ro, rd, vd, tgt, step, _cfg(scene), near=2.0, stepsize=0.5, bg=1.0,
                             stages=stages)
    return gold, m, out, loss, stages


def _cfg(scene):
    # the fixture applies both TV terms at global_step 5000 (make_fine_golden.py): tv_every=1
    return dict(scene.FINE_TRAIN, tv_every=1)


def test_forward_matches_reference_python(run):
    gold, m, out, loss, stages = run
    # same torch ops on the same device: ids/masks identical, floats identical up to op-order noise
    assert np.array_equal(out["ray_id"].numpy(), gold["out_ray_id"])
    assert np.array_equal(out["mask"].numpy(), gold["out_mask"])
    assert np.array_equal(out["mask_outbbox"].numpy(), gold["out_mask_outbbox"])
    for k in ("weights", "raw_alpha", "alphainv_cum", "gradient", "normal", "raw_rgb", "rgb_marched", "sigmoid_rgb"):
        np.testing.assert_allclose(out[k].detach().numpy(), gold["out_" + k], rtol=1e-5, atol=1e-7, err_msg=k)
    assert abs(out["s_val"] - float(gold["s_val"])) < 1e-12
    assert np.array_equal(m.nonempty_mask.numpy(), gold["nonempty_mask"])


def test_loss_and_grads_match_reference_python(run):
    gold, m, out, loss, stages = run
    total = float(gold["loss"]) + float(gold["tv_loss"])
    assert abs(float(loss.detach()) - total) <= 1e-6 * abs(total)
    got = dict(m.params())
    for name in ("rgbnet.0.weight", "rgbnet.3.bias", "refnet.0.weight", "refnet.3.weight", "refnet.3.bias"):
        np.testing.assert_allclose(got[name].grad.numpy(), gold["grad_" + name], rtol=1e-4, atol=1e-9, err_msg=name)
    np.testing.assert_allclose(got["k0"].grad.numpy(), fine_case.dense_k0_grad(gold, got["k0"].shape), rtol=1e-4, atol=1e-10)
    # sdf.grad after the TV add-grad of nerf_training.py:353-360
    np.testing.assert_allclose(got["sdf"].grad.numpy(), gold["sdf_grad_after_tv"], rtol=1e-4, atol=1e-9)


def test_stage_counts_are_sane(run):
    gold, m, out, loss, stages = run
    assert stages["total_len"] >= stages["M0"] >= stages["M1"] >= stages["Ma"] >= stages["M2"] > 0
    assert stages["M2"] == gold["out_weights"].shape[0]
