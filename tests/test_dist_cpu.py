"""CPU tests (gloo, world_size 2) of the host-side multi-GPU logic: ray sharding and the gradient
exchange of fgs_nerf_b200.dist.  The kernels themselves need a GPU (tests/test_multi_gpu.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ray_shard_partitions_the_batch():
    from fgs_nerf_b200.dist import ray_shard
    for n, w in ((8192, 1), (8192, 2), (65536, 8), (10, 4), (3, 8), (0, 2)):
        spans = [ray_shard(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ray_shard(8, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fgs_nerf_b200.dist import all_reduce_sum_
        g = torch.Generator().manual_seed(100 + rank)
        big = torch.randn(1, 3, 8, 8, 8, generator=g)                                   # "grid", reduced in place
        cl = torch.randn(1, 12, 6, 6, 6, generator=g).contiguous(memory_format=torch.channels_last_3d)
        small = [torch.randn(5, 7, generator=g), torch.randn(7, generator=g), torch.randn(3, 4, generator=g)]
        mine = [t.clone() for t in [big, cl] + small]
        all_reduce_sum_([big, cl] + small, bucket_numel=500)
        # expected: sum over both ranks' tensors, regenerated locally
        want = None
        for r in range(world):
            gg = torch.Generator().manual_seed(100 + r)
            ts = [torch.randn(1, 3, 8, 8, 8, generator=gg),
                  torch.randn(1, 12, 6, 6, 6, generator=gg).contiguous(memory_format=torch.channels_last_3d),
                  torch.randn(5, 7, generator=gg), torch.randn(7, generator=gg), torch.randn(3, 4, generator=gg)]
            want = ts if want is None else [a + b for a, b in zip(want, ts)]
        ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip([big, cl] + small, want))
        ok = ok and cl.is_contiguous(memory_format=torch.channels_last_3d)                # layout untouched
        ok = ok and not torch.equal(mine[0], big)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
