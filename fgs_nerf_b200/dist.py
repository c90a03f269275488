"""Ray-sharded data parallelism for the fused step: one process per GPU, each
marches its slice of the ray batch against replicated grids and MLPs; the only exchange is the
gradient sum before the masked Adam step.  The reference has no multi-GPU path at all
(single process, single device: run.py:11, model/nerf_training.py:66-69)."""
import torch


def ray_shard(n_global, world_size, rank):
    """Contiguous slice [start, stop) of a batch of n_global rays owned by `rank`.  Rays are
    ray-major in every flat sample array, so segments never straddle ranks; the last rank owns the
    batch's last ray (the reference's entropy-last-ray quirk, model/nerf_training.py:317)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(n_global), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def chunks_of_rank(n_chunks, world_size, rank):
    """Rendering (model/utils.py:296-299 cuts an image into 8192-ray chunks): chunk ci is rendered by rank
    ci % world_size as its local chunk ci // world_size."""
    return range(rank, n_chunks, world_size)


def local_chunk_count(n_chunks, world_size):
    return (n_chunks + world_size - 1) // world_size


def gather_dealt_chunks(local, chunk_rows, world_size, group=None):
    """local: [n_local * chunk_rows, c], this rank's chunks in its local order (the same n_local on every rank; a rank
    with one chunk fewer leaves its last slot as it is).  Returns [n_local * world_size * chunk_rows, c] in GLOBAL
    chunk order: one all_gather -- every rank sends only what it rendered -- and a strided view."""
    import torch.distributed as dist
    if world_size == 1:
        return local
    n_local = local.shape[0] // chunk_rows
    c = local.shape[1]
    full = torch.empty((world_size * local.shape[0], c), dtype=local.dtype, device=local.device)   # rank-major
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full.view(world_size, n_local, chunk_rows, c).permute(1, 0, 2, 3).reshape(n_local * world_size * chunk_rows, c)
