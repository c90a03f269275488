"""GPU tests of the data-parallel path (need >= 2 devices; skipped otherwise).

1. The ray-sharded 2-GPU step produces the GRADIENTS of the 1-GPU full-batch step after its exchange
   (union of touched bricks -> pack -> one all-reduce over [g_sdf | MLP bucket | packed bricks] -> unpack, then the
   TV add-grads): every tensor within the north star's 1e-4 (relative L2 and max norm; observed ~1e-6, the sums
   only differ in order), at a small size and at BASELINE config 4's shape (256^3, 8192 rays per rank).
2. Two full steps (Adam included) leave parameters whose first update lies inside the element-wise interval the
   1e-4 gradient bar allows (tests/parity_util.py) and whose second update agrees to Adam's noise floor.
3. engine.render with the chunks dealt to two ranks returns bit-identical images to the 1-GPU render.
A log of a passing 2-GPU run is kept under profiles/."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    # name: (grid, mask grid, global rays, image H, W, views)
    "small": dict(G=48, mask_G=40, N=512, H=100, W=100, n_views=8),
    "config4": dict(G=256, mask_G=120, N=16384, H=800, W=800, n_views=40),      # 8192 rays per rank at 256^3
}


def _worker(rank, world, port, q, case, voxel_exchange):
    """Never leave the parent waiting: whatever happens, one entry goes on the queue."""
    try:
        _worker_body(rank, world, port, q, case, voxel_exchange)
    except BaseException as e:                      # noqa: BLE001 -- reported to the parent, which fails the test
        import traceback
        q.put((rank, {"error": f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"}))


def _worker_body(rank, world, port, q, case, voxel_exchange):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from fgs_nerf_b200 import engine, scene
        from fgs_nerf_b200.dist import ray_shard
        import parity_util as U
        c = CASES[case]
        N, step = c["N"], 5001            # 5001 % 3 == 0: both TV terms are on
        sc = scene.make_scene(G=c["G"], mask_G=c["mask_G"])
        rays = [t.to(dev) for t in scene.make_rays(N=N, H=c["H"], W=c["W"], n_views=c["n_views"])]
        a, b = ray_shard(N, world, rank)
        shard = [t[a:b].contiguous() for t in rays]
        kw = dict(cap_m2_per_ray=24)
        out = {}
        # ---- 1. gradients after the exchange ----
        eng = engine.FineStepEngine(sc, b - a, dev, world_size=world, rank=rank, voxel_exchange=voxel_exchange, **kw)
        eng.set_nonempty_mask()
        eng.train_step(*shard, step, n_rays_global=N, has_last_ray=(rank == world - 1), _stop_before_optimizer=True)
        torch.cuda.synchronize()
        eng.check_overflow()
        if rank == 0:
            ref = engine.FineStepEngine(sc, N, dev, **kw)
            ref.set_nonempty_mask()
            ref.train_step(*rays, step, _stop_before_optimizer=True)
            torch.cuda.synchronize()
            ref.check_overflow()
            g = {"g_sdf": (eng.g_sdf, ref.g_sdf), "g_k0": (eng.g_k0, ref.g_k0)}
            for i in range(8):
                g[f"gW{i}"] = (eng.weight_grads[i], ref.weight_grads[i])
                g[f"gB{i}"] = (eng.gB[i], ref.gB[i])
            out["grad_err"] = {k: U.rel_errors(m, w) for k, (m, w) in g.items()}
            out["loss"] = (eng.ws["loss"].cpu().tolist(), ref.ws["loss"].cpu().tolist())
            out["M2"] = (eng.counters()["M2"], ref.counters()["M2"])
            g_ref = {"sdf": ref.g_sdf.clone(), "k0": ref.g_k0.clone(), "W0": ref.weight_grads[0].clone(),
                     "W7": ref.weight_grads[7].clone()}
            del ref
        del eng
        torch.cuda.empty_cache()
        # ---- 2. two full steps ----
        eng = engine.FineStepEngine(sc, b - a, dev, world_size=world, rank=rank, voxel_exchange=voxel_exchange, **kw)
        eng.set_nonempty_mask()
        eng.train_step(*shard, step, n_rays_global=N, has_last_ray=(rank == world - 1))
        torch.cuda.synchronize()
        if rank == 0:
            t = eng.train_cfg
            first = {}
            for name, mine, before, lr, masked in (
                    ("sdf", eng.sdf, sc.sdf.to(dev), t["lrate_sdf"], False),
                    ("k0", eng.k0, sc.k0.to(dev), t["lrate_k0"], True),
                    ("W0", eng.weights[0], sc.rgbnet[0][0].to(dev), 1e-3, False),
                    ("W7", eng.weights[7], sc.refnet[3][0].to(dev), 1e-3, False)):
                first[name] = (U.adam_first_step_violations(mine, before, g_ref[name], lr, masked=masked), mine.numel())
            out["first_step_violations"] = first
            p1 = {"sdf": eng.sdf.clone(), "k0": eng.k0.clone(), "W0": eng.weights[0].clone(), "W7": eng.weights[7].clone()}
        eng.train_step(*shard, step + 1, n_rays_global=N, has_last_ray=(rank == world - 1))   # flags / lists left clean?
        torch.cuda.synchronize()
        eng.check_overflow()
        if rank == 0:
            ref = engine.FineStepEngine(sc, N, dev, **kw)
            ref.set_nonempty_mask()
            for k in range(2):
                ref.train_step(*rays, step + k)
            torch.cuda.synchronize()
            second = {}
            for name, mine, want, lr in (("sdf", eng.sdf, ref.sdf, t["lrate_sdf"]), ("k0", eng.k0, ref.k0, t["lrate_k0"]),
                                         ("W0", eng.weights[0], ref.weights[0], 1e-3), ("W7", eng.weights[7], ref.weights[7], 1e-3)):
                d_m, d_w = (mine - p1[name]).float(), (want - p1[name]).float()
                second[name] = float(((d_m - d_w).abs() > 1e-3 * lr + 1e-4 * d_w.abs()).float().mean())
            out["second_step_frac"] = second
            del ref
        # ---- 3. rank-sharded rendering ----
        if case == "small":
            R_total = 5 * 256 + 77
            ro, rd, vd, _ = [t.to(dev) for t in scene.make_rays(N=R_total, H=c["H"], W=c["W"], n_views=c["n_views"])]
            reng = engine.FineStepEngine(sc, 256, dev)
            got = reng.render(ro, rd, vd, global_step=20000, world_size=world, rank=rank)
            torch.cuda.synchronize()
            if rank == 0:
                want = reng.render(ro, rd, vd, global_step=20000)
                out["render_equal"] = {k: bool(torch.equal(got[k], want[k])) for k in want}
                out["render_shape"] = tuple(got["rgb_marched"].shape)
        dist.barrier()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,voxel_exchange", [("small", False), ("small", True), ("config4", False)])
def test_two_gpu_step_equals_single_gpu_step(case, voxel_exchange):
    """voxel_exchange: the touched part of g_k0 travels as packed 4^3 bricks (default) or as packed voxels."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + 3 * list(CASES).index(case) + (1 if voxel_exchange else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, case, voxel_exchange)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    import queue as _queue
    import time as _time
    deadline = _time.time() + 600
    while len(res) < len(procs) and _time.time() < deadline:
        try:
            r, o = q.get(timeout=5)
            res[r] = o
            if isinstance(o, dict) and "error" in o:
                break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):      # a worker died without reporting
                break
    for p in procs:
        p.join(timeout=20)
        if p.is_alive():
            p.terminate()
    errs = {r: o["error"] for r, o in res.items() if isinstance(o, dict) and "error" in o}
    assert not errs, errs
    assert len(res) == len(procs), f"workers did not report: got {sorted(res)}; exit codes {[p.exitcode for p in procs]}"
    out = res[0]
    print(f"[{case}, voxel_exchange={voxel_exchange}] M2 (rank 0 shard, full batch): {out['M2']}")
    print("  gradient errors after the exchange (relative L2, max norm):")
    for k, (l2, mx) in out["grad_err"].items():
        print(f"    {k:6s} {l2:.3e} {mx:.3e}")
    print("  first Adam step, elements outside the interval the 1e-4 gradient bar allows:", out["first_step_violations"])
    print("  second step, fraction of updates beyond Adam's noise floor:", out["second_step_frac"])
    print("  loss terms (rank 0 shard vs full batch):", out["loss"])
    for k, (l2, mx) in out["grad_err"].items():
        assert l2 <= 1e-4 and mx <= 1e-4, (k, l2, mx)
    for k, (bad, n) in out["first_step_violations"].items():
        assert bad == 0, f"{k}: {bad} of {n} first-step updates outside the element-wise bound"
    for k, v in out["second_step_frac"].items():
        assert v < 2e-3, (k, v)
    if case == "small":
        assert all(out["render_equal"].values()), out["render_equal"]
        assert out["render_shape"] == (5 * 256 + 77, 3)
