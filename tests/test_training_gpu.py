"""GPU tests of the between-step operators and the training-loop host side,
through the C-ABI of include/fgs_b200_train.h:
  * fgs_grid_resize_trilinear against torch's CUDA F.interpolate (the reference's call, model/grid.py:106)
    and against the numpy oracle (oracle/stage_ref.py);
  * fgs_gather_rays against torch indexing (bit-exact: it moves bytes);
  * FineTrainer against the same iterations issued by hand with the ORACLE's batches and learning rates,
    and its second iteration against the oracle's own step (oracle/nerf_ref.py + the reference's kernels);
  * a pg_scale iteration: grids resampled as the reference does, optimiser re-created."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import nerf_ref as R  # noqa: E402
from oracle import ref_ext, stage_ref  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,cl,src,dst", [
    (1, False, (48, 48, 48), (77, 77, 77)),        # odd Z2: scalar stores
    (1, False, (40, 40, 40), (64, 64, 64)),        # 128-bit stores
    (12, True, (20, 20, 20), (52, 52, 52)),        # the engine's k0 layout
    (12, False, (16, 12, 20), (24, 30, 28)),       # reference layout, non-cubic
    (3, False, (9, 9, 9), (1, 5, 9)),              # an axis of output size 1 (scale 0), an identity axis
    (4, True, (10, 10, 10), (7, 6, 5)),            # shrinking
])
def test_grid_resize_matches_aten_and_oracle(cuda, C, cl, src, dst):
    from fgs_nerf_b200 import engine
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, C, *src, generator=g)
    xd = x.to(cuda)
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last_3d)
    got = engine.grid_resize(xd, dst)
    assert tuple(got.shape) == (1, C, *dst)
    assert got.is_contiguous() != cl or C == 1
    want = F.interpolate(x.to(cuda), size=dst, mode="trilinear", align_corners=True)
    scale = max(1.0, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert err <= 2e-6 * scale, f"vs torch CUDA F.interpolate: {err:.3e}"
    orc = torch.from_numpy(stage_ref.grid_resize_trilinear(x[0].numpy(), dst))[None]
    err = float((got.cpu() - orc).abs().max())
    assert err <= 2e-6 * scale, f"vs the numpy oracle: {err:.3e}"
    print("bit-identical to torch CUDA:", float((got == want).float().mean()))


def test_grid_resize_full_size_properties(cuda):
    """160^3 -> 256^3 (the pg_scale step of config/shiny_blender.py:203-204) on the 12-channel k0 layout:
    the first corner is copied exactly, a linear ramp stays a linear ramp, the range does not widen."""
    from fgs_nerf_b200 import engine
    G, G2 = 160, 256
    lin = torch.linspace(-1, 1, G, device=cuda)
    ramp = (0.25 * lin[:, None, None] - 0.5 * lin[None, :, None] + lin[None, None, :])
    x = torch.randn(1, 12, G, G, G, device=cuda, generator=torch.Generator(device=cuda).manual_seed(3))
    x[0, 0] = ramp
    xc = x.contiguous(memory_format=torch.channels_last_3d)
    y = engine.grid_resize(xc, (G2, G2, G2))
    assert torch.equal(y[0, :, 0, 0, 0], x[0, :, 0, 0, 0])      # index 0: weights exactly (1, 0) on every axis
    assert float((y[0, :, -1, -1, -1] - x[0, :, -1, -1, -1]).abs().max()) < 1e-4   # fl(159/255)*255 need not be 159
    lin2 = torch.linspace(-1, 1, G2, device=cuda)
    ramp2 = (0.25 * lin2[:, None, None] - 0.5 * lin2[None, :, None] + lin2[None, None, :])
    assert float((y[0, 0] - ramp2).abs().max()) < 2e-6
    assert float(y.max()) <= float(x.max()) + 1e-5 and float(y.min()) >= float(x.min()) - 1e-5
    want = F.interpolate(x[:, 1:3], size=(G2, G2, G2), mode="trilinear", align_corners=True)
    assert float((y[:, 1:3] - want).abs().max()) <= 2e-6 * float(want.abs().max())


def test_grid_resize_rejects_bad_arguments(cuda):
    from fgs_nerf_b200 import engine
    with pytest.raises(RuntimeError):
        engine.grid_resize(torch.zeros(1, 1, 4, 4, 4), (8, 8, 8))                     # not a CUDA tensor
    with pytest.raises(RuntimeError):
        engine.grid_resize(torch.zeros(2, 1, 4, 4, 4, device=cuda), (8, 8, 8))       # batch must be 1
    with pytest.raises(RuntimeError):
        engine.grid_resize(torch.zeros(1, 1, 4, 4, 4, device=cuda), (8, 0, 8))       # empty output


def test_gather_rays_matches_indexing(cuda):
    from fgs_nerf_b200 import engine
    g = torch.Generator().manual_seed(21)
    n_total, n = 5000, 777
    arrs = [torch.randn(n_total, 3, generator=g).to(cuda) for _ in range(4)]
    sel = torch.randint(0, n_total, (n,), generator=g).to(cuda)
    bad = torch.zeros(1, dtype=torch.int32, device=cuda)
    out = engine.gather_rays(*arrs, sel, bad=bad)
    for o, a in zip(out, arrs):
        assert torch.equal(o, a[sel])
    assert int(bad.item()) == 0
    want = stage_ref.gather_rays(*[a.cpu().numpy() for a in arrs], sel.cpu().numpy())
    for o, w in zip(out, want):
        assert np.array_equal(o.cpu().numpy(), w)
    sel2 = sel.clone()
    sel2[5] = n_total                                   # one past the end: flagged, row zero-filled, rest intact
    out2 = engine.gather_rays(*arrs, sel2, bad=bad)
    assert int(bad.item()) == 1
    assert float(out2[0][5].abs().sum()) == 0.0 and torch.equal(out2[0][6:], arrs[0][sel[6:]])
    empty = engine.gather_rays(*arrs, sel[:0])
    assert empty[0].shape == (0, 3)


def _training_rays(cuda, n_total):
    from fgs_nerf_b200 import scene
    return [t.to(cuda) for t in scene.make_rays(N=n_total, H=100, W=100, n_views=8)]


def _frac_differing(a, b, before, lr):
    """Share of parameters whose accumulated update differs by more than Adam's noise floor (an entry whose gradient
    is at rounding-noise level moves by ~lr in a direction rounding decides; see test_fine_gpu)."""
    d_a, d_b = (a.detach() - before).float(), (b.detach() - before).float()
    return float(((d_a - d_b).abs() > 1e-3 * lr + 1e-4 * d_b.abs()).float().mean())


def test_trainer_equals_hand_issued_iterations(cuda):
    """FineTrainer.run(3) against three engine.train_step calls fed by the ORACLE: batches from torch indexing with the
    reference generator's indices, learning rates from oracle RefSchedule (pinned to the reference's block on the CPU)."""
    from fgs_nerf_b200 import engine, scene, training
    N, n_total, K = 192, 1500, 3
    cfg = dict(scene.FINE_TRAIN, tv_every=2, N_rand=N, lrate_decay=20, N_iters=20000, cosine_lr=True,
               cosine_lr_cfg=dict(warm_up_iters=0, const_warm_up=True, warm_up_min_ratio=1.0),
               decay_step_module={4999: dict(sdf=0.1)}, pg_scale=[])     # keyed on global_step - 1: after iteration 5000
    rays = _training_rays(cuda, n_total)
    sc = scene.make_scene(G=48, mask_G=40)
    start = 4999

    eng = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN))
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=start, rng=np.random.RandomState(7))

    ref_t = dict(cfg, tv_terms=dict(sdf_tv=cfg["sdf_tv"], smooth_grad_tv=cfg["smooth_grad_tv"]))
    sch = stage_ref.RefSchedule(ref_t)
    sch._fresh_optimizer(0)                            # create_optimizer_or_freeze_model(global_step=0), nerf_training.py:117
    eng2 = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(cfg))
    eng2.set_nonempty_mask()
    eng2._ensure_adam()
    gen = training.batch_indices_generator(n_total, N, rng=np.random.RandomState(7))
    losses = []
    # the scatter-adds are float atomics, so two runs agree to rounding (amplified by Adam where a gradient is itself
    # rounding noise), not bit for bit.  Lockstep: before every iteration the hand-issued engine takes over the
    # trainer's parameters and Adam moments, so that one decision taken differently after a last-bit difference is not
    # integrated over the following iterations (see test_graph_launched_step_equals_directly_enqueued_step).
    for i in range(K):
        gs = start + 1 + i
        with torch.no_grad():
            for x, y in zip(_engine_state(eng2), _engine_state(tr.engine)):
                x.copy_(y)
        before = {"sdf": eng2.sdf.clone(), "k0": eng2.k0.clone(), "W0": eng2.weights[0].clone(), "W7": eng2.weights[7].clone()}
        assert tr.run(1) == gs
        sel = torch.from_numpy(next(gen)).to(cuda)
        ro, rd, vd, tgt = [r[sel].contiguous() for r in rays]
        eng2.lr = dict(sch.lr)
        eng2.train_step(ro, rd, vd, tgt, gs)
        torch.cuda.synchronize()
        losses.append(eng2.ws["loss"].cpu().numpy().copy())
        for name, a, b in (("sdf", tr.engine.sdf, eng2.sdf), ("k0", tr.engine.k0, eng2.k0),
                           ("W0", tr.engine.weights[0], eng2.weights[0]), ("W7", tr.engine.weights[7], eng2.weights[7])):
            lr = {"sdf": sch.lr["sdf"], "k0": sch.lr["k0"], "W0": sch.lr["rgbnet"], "W7": sch.lr["refnet"]}[name]
            frac = _frac_differing(a, b, before[name], lr)
            assert frac < 1e-2, f"{name}: {frac:.5f} of the parameters differ after iteration {gs}"
        sch.iteration(gs)
    st = tr.stats()
    assert st["steps"] == [start + 1, start + 2, start + 3] and np.isfinite(st["psnr"]) and st["terms"].shape == (K, 5)
    assert tr.engine.lr == pytest.approx(sch.lr, rel=1e-13)
    assert sch.lr["sdf"] < 0.11 * sch.lr["k0"] * cfg["lrate_sdf"] / cfg["lrate_k0"]      # decay_step_module applied
    assert np.allclose(st["terms"], np.stack(losses), rtol=1e-6, atol=1e-12)


def test_trainer_second_iteration_against_oracle_step(cuda):
    """Two iterations of the trainer against two of the oracle's steps (oracle/nerf_ref.py with the reference's
    compiled kernels), the second one with a learning rate the schedule has changed (sdf x0.1)."""
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    from fgs_nerf_b200 import engine, scene, training
    N, n_total = 192, 1500
    cfg = dict(scene.FINE_TRAIN, tv_every=1, N_rand=N, lrate_decay=20, decay_step_module={5000: dict(sdf=0.1)}, pg_scale=[])
    rays = _training_rays(cuda, n_total)
    sc = scene.make_scene(G=48, mask_G=40)
    start = 5000
    eng = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN))
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=start, rng=np.random.RandomState(9))
    tr.run(2)
    torch.cuda.synchronize()

    native = R.RefCudaNative()
    m = R.FineModel(sc, device=cuda)
    m.set_nonempty_mask()
    opt = R.RefOptimizer(m, cfg, native)
    sch = stage_ref.RefSchedule(dict(cfg, tv_terms=dict(sdf_tv=cfg["sdf_tv"], smooth_grad_tv=cfg["smooth_grad_tv"])))
    sch._fresh_optimizer(0)
    gen = training.batch_indices_generator(n_total, N, rng=np.random.RandomState(9))
    before = {k: v.detach().clone() for k, v in m.params()}
    for gs in (start + 1, start + 2):
        sel = torch.from_numpy(next(gen)).to(cuda)
        ro, rd, vd, tgt = [r[sel].contiguous() for r in rays]
        for g in opt.groups:
            g["lr"] = sch.lr[g["name"].split(".")[0]]
        R.train_step(m, native, opt, ro, rd, vd, tgt, gs, cfg, near=2.0, stepsize=0.5, bg=1.0)
        sch.iteration(gs)
    ref = dict(m.params())
    assert sch.lr["sdf"] == pytest.approx(tr.engine.lr["sdf"], rel=1e-13)
    for name, mine, lr in (("sdf", tr.engine.sdf, cfg["lrate_sdf"]), ("k0", tr.engine.k0, cfg["lrate_k0"]),
                           ("rgbnet.0.weight", tr.engine.weights[0], 1e-3), ("refnet.3.weight", tr.engine.weights[7], 1e-3)):
        frac = _frac_differing(mine, ref[name], before[name], lr)
        print(f"  {name}: {frac:.6f} of the two-iteration updates differ from the oracle's")
        # observed on a B200: sdf 6e-5 (entries whose gradient is rounding noise: Adam moves them by ~lr in a direction
        # the last bit decides), k0 / MLP weights 0; the bound leaves an order of magnitude
        assert frac < 1e-3, f"{name}: {frac:.5f} of the two-iteration updates differ from the oracle's"
    # the sdf step of iteration 2 used a tenth of the rate: |update| <= lr1 + 1.001 lr2 (Adam's bias-corrected ratio
    # is at most 1.001 at step 2), where two full-rate steps would reach 2 lr1
    lr1 = cfg["lrate_sdf"]                       # a fresh optimiser starts at the base rate (global_step=0)
    d = (tr.engine.sdf - before["sdf"]).abs().max()
    assert float(d) <= lr1 * (1.0 + 0.1 * 1.001) * 1.02
    assert float(d) >= lr1 * 0.99


def test_pg_scale_iteration_rescales_grids_and_restarts_optimizer(cuda):
    from fgs_nerf_b200 import engine, scene, training
    N, n_total = 192, 1500
    ratio = (64 / 48) ** 3
    cfg = dict(scene.FINE_TRAIN, tv_every=1, N_rand=N, lrate_decay=20, pg_scale=[5002], scale_ratio=ratio)
    rays = _training_rays(cuda, n_total)
    sc = scene.make_scene(G=48, mask_G=40)
    eng = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN))
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=5000, rng=np.random.RandomState(1))
    tr.step()
    torch.cuda.synchronize()
    sdf_before, k0_before = tr.engine.sdf.clone(), tr.engine.k0.clone()
    assert tr.engine.adam_step == 1
    # what the reference computes (model/nerf.py:298-303, 369-381; model/grid.py:101-106)
    ext = (sc.xyz_max - sc.xyz_min).float()
    vs = (ext.prod() / (48 ** 3 * ratio)).pow(1 / 3)
    ws = (ext / vs).long().tolist()
    want_sdf = F.interpolate(sdf_before, size=tuple(ws), mode="trilinear", align_corners=True)
    want_k0 = F.interpolate(k0_before.contiguous(), size=tuple(ws), mode="trilinear", align_corners=True)
    new = tr.engine.scale_volume_grid(tr.engine.num_voxels * ratio)       # the call step() makes at a pg_scale iteration
    assert [new.X, new.Y, new.Z] == ws and abs(float(new.voxel_size) - float(vs)) < 1e-9
    assert float((new.sdf - want_sdf).abs().max()) <= 2e-6 * float(want_sdf.abs().max())
    assert float((new.k0 - want_k0).abs().max()) <= 2e-6 * float(want_k0.abs().max())
    assert new.nonempty_mask is not None and tuple(new.nonempty_mask.shape) == tuple(ws)
    assert new.adam_step == 0 and not new.adam and new.lr == {k: cfg["lrate_" + k] for k in new.lr}
    for a, b in zip(new.weights, tr.engine.weights):
        assert torch.equal(a, b)
    # ... and through the loop: iteration 5002 rescales, trains on, with a fresh optimiser
    tr2 = training.FineTrainer(engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN)), *rays, cfg_train=cfg,
                               start=5000, rng=np.random.RandomState(1))
    tr2.engine.set_nonempty_mask()
    tr2.run(3)
    st = tr2.stats()
    assert [tr2.engine.X, tr2.engine.Y, tr2.engine.Z] == ws
    assert tr2.engine.adam_step == 2 and np.isfinite(st["terms"]).all() and st["terms"][:, 0].min() > 0
    f = 0.1 ** (1 / 20000)
    assert tr2.engine.lr["k0"] == pytest.approx(cfg["lrate_k0"] * f * f, rel=1e-12)
    assert tr2.engine.counters()["M2"] > 0


def test_resume_keeps_optimizer_state_and_rates(cuda):
    """utils.load_checkpoint -> optimizer.load_state_dict (model/utils.py:49-56): a trainer built on a reloaded engine
    continues with the saved moments, step count and per-group learning rates -- the 4th iteration after a resume at 3
    equals the 4th iteration of the uninterrupted run (up to the float atomics of the scatters)."""
    from fgs_nerf_b200 import checkpoint, engine, scene, training
    N, n_total = 192, 1500
    cfg = dict(scene.FINE_TRAIN, tv_every=1, N_rand=N, lrate_decay=20, pg_scale=[])
    rays = _training_rays(cuda, n_total)
    sc = scene.make_scene(G=48, mask_G=40)
    eng = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN))
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=0, rng=np.random.RandomState(3))
    tr.run(3)
    torch.cuda.synchronize()
    ck = checkpoint.checkpoint_dict(tr.engine, 3)
    before = {"sdf": tr.engine.sdf.clone(), "k0": tr.engine.k0.clone(), "W0": tr.engine.weights[0].clone()}
    lr_at_3 = dict(tr.engine.lr)
    tr.run(1)
    torch.cuda.synchronize()

    eng2, start = checkpoint.engine_from_checkpoint(ck, N, cuda, train_cfg=dict(scene.FINE_TRAIN))
    assert start == 3 and eng2.adam_step == 3
    tr2 = training.FineTrainer(eng2, *rays, cfg_train=cfg, start=start, rng=np.random.RandomState(3))
    assert eng2.adam_step == 3 and eng2.adam, "the loaded optimiser state was thrown away"
    assert eng2.lr == pytest.approx(lr_at_3, rel=1e-12)
    for _ in range(3):
        next(tr2.sampler)                                  # the index stream continues where the first run stood
    tr2.run(1)
    torch.cuda.synchronize()
    assert tr2.engine.adam_step == 4 and tr2.engine.lr == pytest.approx(tr.engine.lr, rel=1e-12)
    for name, a, b, lr in (("sdf", tr2.engine.sdf, tr.engine.sdf, cfg["lrate_sdf"]),
                           ("k0", tr2.engine.k0, tr.engine.k0, cfg["lrate_k0"]),
                           ("W0", tr2.engine.weights[0], tr.engine.weights[0], 1e-3)):
        frac = _frac_differing(a, b, before[name], lr)
        assert frac < 1e-2, f"{name}: {frac:.5f} of the 4th-iteration updates differ after the resume"
        # a restarted Adam (zero moments, step 1) would move every touched entry by the full lr: measurably different
        moved = (b - before[name]).abs()
        assert float((a - b).abs().max()) < 0.5 * float(moved.max())
    # --no_reload_optimizer: fresh state, base rates
    eng3, _ = checkpoint.engine_from_checkpoint(ck, N, cuda, train_cfg=dict(scene.FINE_TRAIN), no_reload_optimizer=True)
    tr3 = training.FineTrainer(eng3, *rays, cfg_train=cfg, start=3, rng=np.random.RandomState(3))
    assert tr3.engine.adam_step == 0 and tr3.engine.lr == {k: cfg["lrate_" + k] for k in tr3.engine.lr}


def test_workspace_overflow_is_sticky_across_iterations(cuda):
    """An iteration that runs out of level-2 workspace must not be forgotten when a later one fits: run() with
    i_print=0 never reads the per-forward counters, the sticky word still raises."""
    from fgs_nerf_b200 import engine, scene, training
    N, n_total = 192, 1500
    cfg = dict(scene.FINE_TRAIN, tv_every=1, N_rand=N, lrate_decay=20, pg_scale=[])
    rays = _training_rays(cuda, n_total)
    sc = scene.make_scene(G=48, mask_G=40)
    eng = engine.FineStepEngine(sc, N, cuda, train_cfg=dict(scene.FINE_TRAIN), cap_m2_per_ray=1)
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=5000, rng=np.random.RandomState(1))
    with pytest.raises(RuntimeError, match="overflow"):
        tr.run(2)
    eng.check_overflow()                                   # cleared by the check that raised
    # a forward that fits leaves the per-forward word clean while the sticky one still remembers the earlier overflow
    ro, rd, vd, tgt = [r[:N].contiguous() for r in rays]
    eng.train_step(ro, rd, vd, tgt, 5001)                  # overflows (cap_m2 = 1 per ray)
    far = ro + 100.0                                       # rays that miss the box: no samples, no overflow
    eng.forward(far, rd, vd, 5001)
    torch.cuda.synchronize()
    assert eng.counters()["OVERFLOW"] == 0
    with pytest.raises(RuntimeError, match="earlier iteration"):
        eng.check_overflow()


@pytest.mark.parametrize("res,bound", [(33, 1.0), (40, 1.1), (64, 0.7)])
def test_sdf_field_query_matches_reference_blockwise_query(cuda, res, bound):
    """extract_fields (model/extract_geometry.py:5-19) with nerf.extract_geometry's query_func (model/nerf.py:1163:
    grid_sampler(pts, -sdf_grid)): the one-launch lattice query against the oracle's F.grid_sample on the
    materialised points, block by block like the reference; bit-exact (same ATen arithmetic), including the zero
    padding outside the bbox (bound 1.1)."""
    from fgs_nerf_b200 import engine, extract_geometry, scene
    sc = scene.make_scene(G=48, mask_G=40)
    m = R.FineModel(sc, device=cuda, requires_grad=False)
    bmin, bmax = torch.tensor([-bound] * 3), torch.tensor([bound, bound * 0.9, bound])
    with torch.no_grad():
        query = lambda pts: R.dense_grid(m, -m.sdf, pts.to(cuda))          # noqa: E731
        want = extract_geometry.extract_fields(bmin, bmax, res, query, N=16)      # the reference's block loop
    got = extract_geometry.extract_sdf_field(sc.sdf.to(cuda), sc.xyz_min, sc.xyz_max, bmin, bmax, res)
    assert tuple(got.shape) == (res, res, res)
    got = got.cpu().numpy()
    assert np.array_equal(got, want), f"max abs diff {np.abs(got - want).max():.3e}"
    eng = engine.FineStepEngine(sc, 64, cuda)
    assert np.array_equal(eng.extract_fields(bmin, bmax, res).cpu().numpy(), want)


@pytest.mark.parametrize("k,s,shape", [(3, 1, (12, 10, 14)), (5, 1, (33, 20, 70)), (5, 0.8, (8, 8, 32)), (7, 2, (16, 9, 40)),
                                       (1, 1, (5, 6, 7))])
def test_gaussian_conv_matches_torch_conv3d_and_oracle(cuda, k, s, shape):
    """fgs_conv3d_replicate against nn.Conv3d(1, 1, k, padding=k//2, padding_mode='replicate') with the reference's
    kernel (model/nerf.py:260-272) on CUDA, and against the numpy oracle; ragged tiles and every supported ksize."""
    from fgs_nerf_b200 import engine
    g = torch.Generator().manual_seed(41)
    vol = torch.randn(1, 1, *shape, generator=g)
    w = engine.gaussian_3dconv_weights(k, s, cuda)
    assert np.abs(w.numpy() - stage_ref.gaussian_3dconv_weights(k, s)).max() <= 1.2e-7 * float(w.max())
    m = torch.nn.Conv3d(1, 1, k, stride=1, padding=k // 2, padding_mode='replicate').to(cuda)
    m.weight.data = w[None, None].to(cuda)
    m.bias.data = torch.zeros(1, device=cuda)
    with torch.no_grad():
        want = m(vol.to(cuda))
    got = engine.smooth_conv(vol.to(cuda), k, s)
    tol = 2e-6 * float(want.abs().max())
    assert float((got - want).abs().max()) <= tol
    orc = stage_ref.gaussian_3dconv(vol[0, 0].numpy(), k, s)
    assert np.abs(got[0, 0].cpu().numpy() - orc).max() <= tol
    # folded divisions: m(x / r) / r, as init_sdf_from_sdf(smooth=True) computes it
    got = engine.smooth_conv(vol.to(cuda), k, s, pre_div=0.3, post_div=0.3)
    with torch.no_grad():
        want = m(vol.to(cuda) / 0.3) / 0.3
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())


def test_stage_handoff_init_sdf_from_sdf(cuda):
    """engine.init_sdf_from_sdf against the reference's steps made with torch's CUDA ops (model/nerf.py:280-296,
    called with smooth=False, reduce=sdf_reduce at nerf_training.py:128; smooth_scale defaults to True) and against
    the numpy oracle; the grid's storage must stay where it is (the C structs hold its pointer)."""
    from fgs_nerf_b200 import engine, scene
    sc = scene.make_scene(G=48, mask_G=40)
    eng = engine.FineStepEngine(sc, 64, cuda)
    eng.set_nonempty_mask()
    g = torch.Generator().manual_seed(43)
    sdf0 = torch.randn(1, 1, 30, 30, 30, generator=g)
    p0 = eng.sdf.data_ptr()
    eng.init_sdf_from_sdf(sdf0, smooth=False, reduce=0.3)
    assert eng.sdf.data_ptr() == p0 and tuple(eng.sdf.shape) == (1, 1, 48, 48, 48)
    x = F.interpolate(sdf0.to(cuda), size=(48, 48, 48), mode="trilinear", align_corners=True) / 0.3
    w = engine.gaussian_3dconv_weights(5, 1, cuda)[None, None].to(cuda)
    want = F.conv3d(F.pad(x, (2,) * 6, mode="replicate"), w)
    tol = 3e-6 * float(want.abs().max())
    assert float((eng.sdf - want).abs().max()) <= tol
    orc = stage_ref.init_sdf_from_sdf(sdf0[0, 0].numpy(), (48, 48, 48), smooth=False, reduce=0.3)
    assert np.abs(eng.sdf[0, 0].cpu().numpy() - orc).max() <= tol
    # the engine still steps on the new grid
    ro, rd, vd, tgt = [t.to(cuda) for t in scene.make_rays(N=64, H=100, W=100, n_views=8)]
    eng.sdf.copy_(sc.sdf.to(cuda))
    eng.init_sdf_from_sdf(sc.sdf, smooth=True, reduce=1.0, ksize=3, sigma=1.0, smooth_scale=False)
    eng.train_step(ro, rd, vd, tgt, 5000)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(eng.ws["rgb_marched"]).all()) and eng.counters()["M1"] > 0


def test_checkpoint_layout_and_round_trip(cuda, tmp_path):
    """The checkpoint dict (model/nerf_training.py:522-531) has the reference's keys, order, shapes and dtypes --
    tests/golden/ckpt_layout.json was dumped from the reference's own model / optimizer objects -- and restores an
    engine that continues exactly where the saved one was."""
    import json
    from fgs_nerf_b200 import checkpoint, engine, scene
    layout = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_layout.json")))
    N = 64
    sc = scene.make_scene(G=24, mask_G=20)                  # the sizes the layout was dumped at
    ro, rd, vd, tgt = [t.to(cuda) for t in scene.make_rays(N=N, H=100, W=100, n_views=8)]
    eng = engine.FineStepEngine(sc, N, cuda)
    eng.set_nonempty_mask()
    for gs in (5001, 5002):
        eng.train_step(ro, rd, vd, tgt, gs)
    eng.lr["sdf"] *= 0.1
    torch.cuda.synchronize()
    ck = checkpoint.checkpoint_dict(eng, 5002)
    const = ("grad_conv.", "tv_smooth_conv.")               # frozen stencils the reference's constructor re-creates
    want = [(k, s, d) for k, s, d in layout["model_state_dict"] if not k.startswith(const)]
    got = [(k, list(v.shape), str(v.dtype)) for k, v in ck["model_state_dict"].items() if k != "sdf_mask.grid"]
    assert got == want
    assert tuple(ck["model_state_dict"]["sdf_mask.grid"].shape) == (1, 1, 24, 24, 24)          # nerf.set_sdf_mask
    assert list(ck["model_kwargs"].keys()) == list(layout["model_kwargs"].keys())
    assert list(ck["MaskCache_kwargs"].keys()) == list(layout["MaskCache_kwargs"].keys())
    osd = ck["optimizer_state_dict"]
    assert len(osd["state"]) == layout["optimizer_state_count"]
    for g, w in zip(osd["param_groups"], layout["optimizer_param_groups"]):
        assert (g["name"], g["skip_zero_grad"], list(g["betas"]), g["eps"], len(g["params"])) == \
            (w["name"], w["skip_zero_grad"], list(w["betas"]), w["eps"], w["params"])
    assert set(osd["state"][0].keys()) == set(layout["optimizer_state_entry"].keys())
    assert list(osd["state"][0]["exp_avg"].shape) == layout["optimizer_state_entry"]["exp_avg"] and osd["state"][0]["step"] == 2
    assert tuple(osd["state"][2]["exp_avg"].shape) == (256, 106) and tuple(osd["state"][10]["exp_avg"].shape) == (256, 307)
    # round trip through torch.save
    path = str(tmp_path / "fine_last.tar")
    checkpoint.save_checkpoint(eng, 5002, path)
    eng2, start = checkpoint.engine_from_checkpoint(path, N, cuda)
    assert start == 5002 and eng2.adam_step == 2 and eng2.lr == eng.lr
    assert torch.equal(eng2.sdf, eng.sdf) and torch.equal(eng2.k0, eng.k0) and torch.equal(eng2.nonempty_mask, eng.nonempty_mask)
    for i in range(8):
        assert torch.equal(eng2.weights[i], eng.weights[i]) and torch.equal(eng2.Bt[i], eng.Bt[i])
        assert torch.equal(eng2.adam["W"][i][0], eng.adam["W"][i][0]) and torch.equal(eng2.adam["B"][i][1], eng.adam["B"][i][1])
    assert torch.equal(eng2.adam["k0"][1], eng.adam["k0"][1]) and torch.equal(eng2.adam["sdf"][0], eng.adam["sdf"][0])
    before = {"sdf": eng.sdf.clone(), "k0": eng.k0.clone(), "W4": eng.weights[4].clone()}
    eng.train_step(ro, rd, vd, tgt, 5003)
    eng2.train_step(ro, rd, vd, tgt, 5003)
    torch.cuda.synchronize()
    assert np.allclose(eng.ws["loss"].cpu().numpy(), eng2.ws["loss"].cpu().numpy(), rtol=1e-6, atol=1e-12)
    for name, a, b, lr in (("sdf", eng.sdf, eng2.sdf, 0.005), ("k0", eng.k0, eng2.k0, 0.1), ("W4", eng.weights[4], eng2.weights[4], 1e-3)):
        assert _frac_differing(a, b, before[name], lr) < 1e-3, name


def test_library_sequenced_step_equals_host_sequenced_step(cuda):
    """fgs_fine_train_step (one call per iteration) against the same iteration sequenced launch by launch from
    Python: three iterations with both TV terms on; parameters, Adam state and loss terms agree to the rounding of
    the float atomics.  The engines run in lockstep (see the graph test below): same state in, same state out."""
    from fgs_nerf_b200 import engine, scene
    N = 192
    sc = scene.make_scene(G=48, mask_G=40)
    cfg = dict(scene.FINE_TRAIN, tv_every=1)
    ro, rd, vd, tgt = [t.to(cuda) for t in scene.make_rays(N=N, H=100, W=100, n_views=8)]
    a, b = [engine.FineStepEngine(sc, N, cuda, train_cfg=cfg) for _ in range(2)]
    for e, native in ((a, True), (b, False)):
        e.set_nonempty_mask()
        e.native_step = native
        e.use_graph = False                                     # the graph launch has its own test
        e._ensure_adam()
    la, lb = [], []
    for gs in (5000, 5001, 5002):
        with torch.no_grad():
            for x, y in zip(_engine_state(a), _engine_state(b)):
                x.copy_(y)
        before = [t.clone() for t in (b.sdf, b.k0, b.weights[0], b.Bt[7])]
        for e, losses in ((a, la), (b, lb)):
            e.lr["sdf"] *= 0.9                                   # the per-group rates reach both drivers
            e.train_step(ro, rd, vd, tgt, gs)
            torch.cuda.synchronize()
            losses.append(e.ws["loss"].cpu().numpy().copy())
        assert np.allclose(la[-1], lb[-1], rtol=1e-6, atol=1e-12), (gs, la[-1], lb[-1])
        for name, x, y, x0, lr in (("sdf", a.sdf, b.sdf, before[0], b.lr["sdf"]), ("k0", a.k0, b.k0, before[1], 0.1),
                                   ("W0", a.weights[0], b.weights[0], before[2], 1e-3), ("B7", a.Bt[7], b.Bt[7], before[3], 1e-3)):
            assert _frac_differing(x, y, x0, lr) < 1e-2, (gs, name)
        assert float((a.adam["sdf"][1] - b.adam["sdf"][1]).abs().max()) <= 1e-3 * float(b.adam["sdf"][1].abs().max())
    assert a.adam_step == b.adam_step == 3 and np.stack(la)[:, 4].min() > 0           # the smooth-TV term was computed
    assert int(a.ws["k0_brick_flags"].sum()) == 0 and float(a.g_sdf.abs().max()) == 0.0   # cleared by the optimiser
    # argument errors come back as RuntimeError, like every other entry point
    a._S.sdf_m = None
    with pytest.raises(RuntimeError):
        a.train_step(ro, rd, vd, tgt, 5003)


def _engine_state(e):
    """Everything an iteration reads and writes: parameters and Adam moments (the gradients are clear between steps)."""
    e._ensure_adam()
    t = [e.sdf, e.k0] + list(e.Wt) + list(e.Bt) + list(e.adam["sdf"]) + list(e.adam["k0"])
    for pair in e.adam["W"] + e.adam["B"]:
        t += list(pair)
    return t


def test_graph_launched_step_equals_directly_enqueued_step(cuda):
    """fgs_fine_train_step_graph (the iteration stream-captured and launched as ONE CUDA graph, its executable updated in
    place every iteration) against the directly enqueued iteration: six iterations that alternate TV / non-TV steps
    (two graph shapes) with changing learning rates and ray batches, and the graph path was really taken.

    The two engines run in LOCKSTEP: before every iteration the graph engine takes over the other's parameters and Adam
    moments, so each comparison is `same state in -> same state out` and only that iteration's float-atomics order
    differs.  (Free-running both for six iterations was measured flaky on hardware, 5 failures in 12 runs: with ~1000
    MLP rows one ReLU / threshold decision taken differently after a last-bit difference changes every weight gradient
    by ~1e-3, which Adam then integrates past any per-element bound -- a property of the trajectory, not of the graph.)"""
    from fgs_nerf_b200 import engine, scene
    N = 192
    sc = scene.make_scene(G=48, mask_G=40)
    cfg = dict(scene.FINE_TRAIN, tv_every=2)
    batches = [[t.to(cuda) for t in scene.make_rays(N=N, H=100, W=100, n_views=8, offset=b * N)] for b in range(2)]
    a, b = [engine.FineStepEngine(sc, N, cuda, train_cfg=cfg) for _ in range(2)]
    for e, graph in ((a, True), (b, False)):
        e.set_nonempty_mask()
        e.use_graph = graph
        e._ensure_adam()
    la, lb = [], []
    for i, gs in enumerate(range(5000, 5006)):
        with torch.no_grad():
            for x, y in zip(_engine_state(a), _engine_state(b)):
                x.copy_(y)
        before = [t.clone() for t in (b.sdf, b.k0, b.weights[0], b.Bt[7])]
        for e, losses in ((a, la), (b, lb)):
            e.lr["sdf"] *= 0.9
            e.train_step(*batches[i % 2], gs)
            torch.cuda.synchronize()
            losses.append(e.ws["loss"].cpu().numpy().copy())
        assert np.allclose(la[-1], lb[-1], rtol=1e-6, atol=1e-12), (i, la[-1], lb[-1])
        for name, x, y, x0, lr in (("sdf", a.sdf, b.sdf, before[0], b.lr["sdf"]), ("k0", a.k0, b.k0, before[1], 0.1),
                                   ("W0", a.weights[0], b.weights[0], before[2], 1e-3), ("B7", a.Bt[7], b.Bt[7], before[3], 1e-3)):
            assert _frac_differing(x, y, x0, lr) < 1e-2, (i, name)
        assert float((a.adam["sdf"][1] - b.adam["sdf"][1]).abs().max()) <= 1e-3 * float(b.adam["sdf"][1].abs().max())
    a.check_overflow()
    b.check_overflow()
    la, lb = np.stack(la), np.stack(lb)
    assert a.graph_steps >= 4 and b.graph_steps == 0           # (the first iterations warm up outside the graph)
    assert a.adam_step == b.adam_step == 6
    assert (la[::2, 4] > 0).all() and (la[1::2, 4] == 0).all()  # smooth TV on the even steps only
