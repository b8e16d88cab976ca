"""Data parallelism for the hot path: one process per GPU, the shape batch sharded on dim 0, ONE all-reduce of the
flat fp32 gradient buffer per optimizer step (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Replaces the reference's single-process torch.nn.DataParallel (train_parsenet.py:90-92 and 8 more call sites): no
parameter broadcast per forward, no gather of outputs; GroupNorm statistics are per shape, so sharding the batch
leaves every per-shape result unchanged."""
import torch
import torch.distributed as dist


def shard_batch(n_items, rank, world):
    """contiguous slice of the batch owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_grads(params, world=None):
    """average .grad of every parameter across ranks with a single flat all-reduce; returns the number of elements"""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= world
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return o
