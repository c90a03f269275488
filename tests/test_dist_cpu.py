"""CPU tests (gloo, world_size 2) of the host-side multi-GPU logic: ray sharding (training) and the chunk dealing /
gathering of multi-rank rendering.  The gradient exchange itself is kernels + NCCL (brick flags -> list -> pack ->
all-reduce -> unpack) and is tested on two GPUs in tests/test_multi_gpu.py."""
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
        from fgs_nerf_b200.dist import chunks_of_rank, gather_dealt_chunks, local_chunk_count
        ok = True
        for n_chunks in (1, 2, 5, 8):
            rows, c = 6, 3
            n_local = local_chunk_count(n_chunks, world)
            local = torch.full((n_local * rows, c), -1.0)
            for li, ci in enumerate(chunks_of_rank(n_chunks, world, rank)):
                local[li * rows:(li + 1) * rows] = float(ci) + torch.arange(rows * c).reshape(rows, c) / 100.0
            full = gather_dealt_chunks(local, rows, world)
            ok = ok and tuple(full.shape) == (n_local * world * rows, c)
            for ci in range(n_chunks):                     # global chunk order, every chunk from the rank that owns it
                want = float(ci) + torch.arange(rows * c).reshape(rows, c) / 100.0
                ok = ok and torch.equal(full[ci * rows:(ci + 1) * rows], want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_render_chunk_gather_gloo_world2():
    """The host side of multi-rank rendering (engine.render): chunks dealt round-robin, one all_gather per output,
    global chunk order restored -- including chunk counts that do not divide by the world size."""
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
