"""CPU tests of the training-loop host side and the stage operators' oracle:
  * the oracle's trilinear resize against torch's F.interpolate -- the call the reference makes
    (model/grid.py:106);
  * the oracle's and the PRODUCT's schedules against vectors produced by executing the reference's own
    schedule block (tests/golden/make_schedule_golden.py);
  * the batch index stream against the reference's generator (same golden file).
The kernels themselves are tested on the GPU (tests/test_training_gpu.py)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "schedule_golden.npz"))
CASES = ["shiny_fine", "shiny_coarse", "dtu_fine"]


def _cfg(case):
    t = json.loads(str(GOLD[case + "_train_json"]))
    m = json.loads(str(GOLD[case + "_model_json"]))
    for k in ("decay_step_module", "tv_updates"):
        if k in t:
            t[k] = {int(a): b for a, b in t[k].items()}
    if "s_updates" in m:
        m["s_updates"] = {int(a): b for a, b in m["s_updates"].items()}
    return t, m


@pytest.mark.parametrize("C,src,dst", [(1, (5, 6, 7), (9, 11, 13)), (12, (8, 8, 8), (20, 20, 20)),
                                       (3, (6, 5, 4), (6, 5, 4)), (2, (7, 7, 7), (1, 4, 16)), (1, (16, 16, 16), (5, 9, 12))])
def test_oracle_resize_matches_torch_interpolate(C, src, dst):
    from oracle import stage_ref
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, C, *src, generator=g)
    want = F.interpolate(x, size=dst, mode="trilinear", align_corners=True)[0].numpy()
    got = stage_ref.grid_resize_trilinear(x[0].numpy(), dst)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
    if src == dst:
        assert np.array_equal(got, x[0].numpy())          # same size: the identity, bit for bit


@pytest.mark.parametrize("case", CASES)
def test_oracle_schedule_matches_reference_block(case):
    from oracle.stage_ref import RefSchedule
    t, m = _cfg(case)
    sch = RefSchedule(t, m)
    names = [str(n) for n in GOLD[case + "_names"]]
    assert names == sch.names
    steps = GOLD[case + "_steps"].tolist()
    n = json.loads(str(GOLD[case + "_cfg"]))["n"]
    k = 0
    for gs in range(1, n + 1):
        sch.iteration(gs)
        if k < len(steps) and gs == steps[k]:
            want = GOLD[case + "_lr"][k]
            got = np.array([sch.lr[x] for x in names])
            assert np.allclose(got, want, rtol=1e-13, atol=0), (gs, got, want)
            assert [sch.tv.get("sdf_tv", 0.0), sch.tv.get("smooth_grad_tv", 0.0)] == GOLD[case + "_tv"][k].tolist()
            assert [sch.s.get("s_ratio", -1.0), sch.s.get("s_start", -1.0)] == GOLD[case + "_s"][k].tolist()
            k += 1
    assert k == len(steps)


@pytest.mark.parametrize("case", CASES)
def test_product_schedule_matches_reference_block(case):
    """fgs_nerf_b200.training.end_of_iteration_updates (what FineTrainer.step applies) + the optimiser reset at a
    pg_scale iteration, driven over the whole run, against the reference's block."""
    from fgs_nerf_b200 import training
    t, m = _cfg(case)
    flat = dict(t)
    flat.update(flat.pop("tv_terms", {}))
    flat["s_updates"] = m.get("s_updates", {})
    names = [str(n) for n in GOLD[case + "_names"]]
    lr = {x: flat["lrate_" + x] for x in names}
    eng_train, eng_model = {}, {}
    steps = GOLD[case + "_steps"].tolist()
    n = json.loads(str(GOLD[case + "_cfg"]))["n"]
    k = 0
    for gs in range(1, n + 1):
        if gs in flat.get("pg_scale", []):
            lr = {x: flat["lrate_" + x] * 0.1 ** (0 / (flat["lrate_decay"] * 1000)) for x in names}  # reset_optimizer(0)
        training.end_of_iteration_updates(flat, lr, gs, eng_train, eng_model)
        if k < len(steps) and gs == steps[k]:
            got = np.array([lr[x] for x in names])
            assert np.allclose(got, GOLD[case + "_lr"][k], rtol=1e-13, atol=0), (gs, got)
            assert [flat.get("sdf_tv", 0.0), flat.get("smooth_grad_tv", 0.0)] == GOLD[case + "_tv"][k].tolist()
            assert [eng_model.get("s_ratio", -1.0), eng_model.get("s_start", -1.0)] == GOLD[case + "_s"][k].tolist()
            k += 1
    assert k == len(steps)


def test_batch_index_stream_matches_reference_generator():
    from fgs_nerf_b200 import training
    want = GOLD["batches_N1000_BS96_seed123"]
    np.random.seed(123)
    gen = training.batch_indices_generator(1000, 96)                # numpy's global state, like the reference
    got = np.stack([next(gen) for _ in range(len(want))])
    assert np.array_equal(got, want)
    smp = training.DeviceBatchSampler(1000, 96, "cpu", rng=np.random.RandomState(123))
    got = np.stack([next(smp).numpy().copy() for _ in range(len(want))])    # views of the live permutation
    assert np.array_equal(got, want)                                # epoch roll-over included (10 batches per epoch)
    with pytest.raises(ValueError):
        training.DeviceBatchSampler(10, 11, "cpu")


def test_unsupported_loop_switches_fail_loudly():
    from fgs_nerf_b200 import training
    from fgs_nerf_b200.scene import FINE_TRAIN
    ok = training.flatten_train_cfg(dict(FINE_TRAIN, tv_terms=dict(sdf_tv=0.2, smooth_grad_tv=0.0)))
    assert ok["sdf_tv"] == 0.2 and "tv_terms" not in ok
    for bad in (dict(weight_rgbper=0.1), dict(voxel_inc=True), dict(ray_sampler="patch"), dict(weight_tv_k0=1e-3),
                dict(reset_iter=[100])):
        with pytest.raises(NotImplementedError):
            training.flatten_train_cfg(dict(FINE_TRAIN, **bad))


def test_cosine_factor_known_values():
    from fgs_nerf_b200 import training
    t = dict(cosine_lr=True, cosine_lr_cfg=dict(warm_up_iters=0, const_warm_up=True, warm_up_min_ratio=1.0), N_iters=20000)
    assert training.lr_decay_factor(t, 1) == 1.0                    # cos(0) / warm-up value 1.0
    prod = 1.0
    for gs in range(1, 10002):
        prod *= training.lr_decay_factor(t, gs)
    assert abs(prod - 0.5) < 1e-9                                   # half-way down the cosine after 10001 iterations
    assert training.lr_decay_factor(dict(lrate_decay=20), 7) == 0.1 ** (1 / 20000)


STAGE = np.load(os.path.join(HERE, "golden", "stage_golden.npz"))


@pytest.mark.parametrize("k,s", [(3, 1), (5, 1), (5, 0.8), (7, 2)])
def test_oracle_gaussian_conv_matches_reference_method(k, s):
    """oracle gaussian_3dconv against the output of the reference's own nerf._gaussian_3dconv (executed from its
    source by tests/golden/make_stage_golden.py): weights bit-exact or 1 ulp, volume to fp32 rounding."""
    from oracle import stage_ref
    w = stage_ref.gaussian_3dconv_weights(k, s)
    assert np.abs(w - STAGE[f"weight_k{k}_s{s}"]).max() <= 1.2e-7 * w.max()
    got = stage_ref.gaussian_3dconv(STAGE["vol"], k, s)
    want = STAGE[f"conv_k{k}_s{s}"]
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_oracle_resize_matches_golden_and_handoff_composes():
    from oracle import stage_ref
    got = stage_ref.grid_resize_trilinear(STAGE["vol"][None], (20, 17, 23))[0]
    assert np.abs(got - STAGE["resize_20_17_23"]).max() <= 2e-6 * np.abs(STAGE["resize_20_17_23"]).max()
    # init_sdf_from_sdf (model/nerf.py:280-296) = resize, / reduce, 5^3 sigma-1 smoothing -- against the same steps
    # made with torch's ops
    vol = torch.from_numpy(STAGE["vol"])[None, None]
    x = F.interpolate(vol, size=(20, 17, 23), mode="trilinear", align_corners=True) / 0.3
    w = torch.from_numpy(STAGE["weight_k5_s1"])[None, None]
    want = F.conv3d(F.pad(x, (2,) * 6, mode="replicate"), w)[0, 0].numpy()
    got = stage_ref.init_sdf_from_sdf(STAGE["vol"], (20, 17, 23), smooth=False, reduce=0.3)
    assert np.abs(got - want).max() <= 3e-6 * np.abs(want).max()


def test_ranks_draw_the_same_global_batch_and_shard_it():
    """Data parallel: every rank runs the same index stream (same seed) and keeps its contiguous slice of the global
    batch (dist.ray_shard), so the union over ranks is exactly the reference's single-process batch."""
    from fgs_nerf_b200 import training
    from fgs_nerf_b200.dist import ray_shard
    want = GOLD["batches_N1000_BS96_seed123"]
    world = 4
    samplers = [training.DeviceBatchSampler(1000, 96, "cpu", rng=np.random.RandomState(123)) for _ in range(world)]
    for b in range(len(want)):
        parts = []
        for r, smp in enumerate(samplers):
            lo, hi = ray_shard(96, world, r)
            parts.append(next(smp)[lo:hi].numpy().copy())
        assert np.array_equal(np.concatenate(parts), want[b])
