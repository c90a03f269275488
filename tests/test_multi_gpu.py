"""GPU test (needs >= 2 devices; skipped otherwise): the ray-sharded 2-GPU step equals the 1-GPU
full-batch step -- same parameters after forward + loss + backward + TV + NCCL gradient sum +
masked Adam, up to summation order."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, voxel_exchange=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from fgs_nerf_b200 import engine, scene
        from fgs_nerf_b200.dist import ray_shard
        N, step = 512, 5001            # 5001 % 3 == 0: both TV terms are on
        sc = scene.make_scene(G=48, mask_G=40)
        rays = [t.to(dev) for t in scene.make_rays(N=N, H=100, W=100, n_views=8)]
        a, b = ray_shard(N, world, rank)
        eng = engine.FineStepEngine(sc, b - a, dev, world_size=world, rank=rank, voxel_exchange=voxel_exchange)
        eng.set_nonempty_mask()
        shard = [t[a:b].contiguous() for t in rays]
        for k in range(2):             # two steps: the second one also checks that the exchange left flags / lists clean
            eng.train_step(*shard, step + k, n_rays_global=N, has_last_ray=(rank == world - 1))
        torch.cuda.synchronize()
        out = None
        if rank == 0:
            ref = engine.FineStepEngine(sc, N, dev)
            ref.set_nonempty_mask()
            for k in range(2):
                ref.train_step(*rays, step + k)
            torch.cuda.synchronize()
            errs = {}
            t = eng.train_cfg
            for name, mine, want, before, lr in (
                    ("sdf", eng.sdf, ref.sdf, sc.sdf.to(dev), t["lrate_sdf"]),
                    ("k0", eng.k0, ref.k0, sc.k0.to(dev), t["lrate_k0"]),
                    ("W0", eng.weights[0], ref.weights[0], sc.rgbnet[0][0].to(dev), 1e-3),
                    ("W7", eng.weights[7], ref.weights[7], sc.refnet[3][0].to(dev), 1e-3)):
                d_m, d_w = (mine - before).float(), (want - before).float()
                bad = (d_m - d_w).abs() > 1e-3 * lr + 1e-4 * d_w.abs()
                errs[name] = float(bad.float().mean())
            loss = (eng.ws["loss"].cpu().tolist(), ref.ws["loss"].cpu().tolist())
            out = (errs, loss)
        dist.barrier()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("voxel_exchange", [False, True])
def test_two_gpu_step_equals_single_gpu_step(voxel_exchange):
    """voxel_exchange: the touched part of g_k0 travels as packed 4^3 bricks (default) or as packed voxels."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + (1 if voxel_exchange else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, voxel_exchange)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    errs, loss = res[0]
    print("fraction of updates that differ:", errs, " rank-0 loss terms vs full-batch:", loss)
    for k, v in errs.items():
        assert v < 2e-3, (k, v)
