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


def all_reduce_sum_(tensors, group=None, bucket_numel=1 << 22):
    """In-place sum over ranks of a list of gradient tensors.  Tensors of at least `bucket_numel`
    elements (the grids) are reduced in place; the small ones (16 MLP tensors) travel in one
    packed bucket, so a step issues 3 collectives instead of 18."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    small = []
    for t in tensors:
        if t.numel() >= bucket_numel and t.is_contiguous():
            dist.all_reduce(t, group=group)
        elif t.numel() >= bucket_numel:
            # channels_last_3d etc.: reduce the underlying storage order, not a contiguous copy
            flat = torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())
            dist.all_reduce(flat, group=group)
        else:
            small.append(t)
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        dist.all_reduce(flat, group=group)
        o = 0
        for t in small:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
