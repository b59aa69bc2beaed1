"""Host-side data-parallel plumbing (one process per GPU, ``torch.distributed``): ray / tile sharding, the final gather of
an inference frame, and the sharded optimizer-update protocol of the main hash table.

Nothing here launches a kernel itself, so it is exercised on CPU with the ``gloo`` backend at world size 2
(``tests/test_dist_cpu.py``); on the GPUs the same functions run over NCCL.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rays(n_total, rank, world_size):
    """Contiguous [start, stop) slice of ``n_total`` rays / pixels owned by ``rank`` (tiles for inference,
    ray shards for training); sizes differ by at most one."""
    base, rem = divmod(n_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_frame(local, n_total, rank, world_size):
    """Inference: every rank renders its tile; one all_gather stitches [rays, C] tensors."""
    if world_size == 1:
        return local
    sizes = [shard_rays(n_total, r, world_size) for r in range(world_size)]
    longest = max(b - a for a, b in sizes)
    pad = torch.zeros(longest, *local.shape[1:], device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], dim=0)


def shard_bounds(a, b, world_size, rank):
    """Equal 16-byte-aligned shards of the flat range [a, b) (FusedAdam slots are multiples of 32 elements)."""
    shard = (b - a) // world_size
    if shard * world_size != b - a or shard % 4:
        raise ValueError(f"flat range [{a}, {b}) does not split into {world_size} aligned shards")
    return a + rank * shard, a + (rank + 1) * shard


def reduce_scatter_range(flat_grad, a, b, world_size=None, rank=None):
    """First half of the sharded update of ``[a, b)``: after it, ``flat_grad[lo:hi]`` (this rank's shard) holds the SUM
    over ranks.  Backends without reduce-scatter (gloo, CPU tests) all-reduce the whole range instead."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(a, b, world_size, rank)
    if dist.get_backend() == "nccl":
        dist.reduce_scatter_tensor(flat_grad[lo:hi], flat_grad[a:b], op=dist.ReduceOp.SUM)        # in place
    else:
        dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM)
    return lo, hi


def apply_and_gather_range(flat_param, flat_grad, a, b, apply_fn, world_size=None, rank=None):
    """Second half: ``apply_fn(lo, hi)`` on this rank's shard (it must consume ``flat_grad[lo:hi]`` and leave it zeroed),
    the rest of ``flat_grad[a:b]`` is zeroed, and the updated shards are all-gathered into ``flat_param[a:b]``."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(a, b, world_size, rank)
    apply_fn(lo, hi)
    if lo > a:
        flat_grad[a:lo].zero_()
    if hi < b:
        flat_grad[hi:b].zero_()
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(flat_param[a:b], flat_param[lo:hi])                            # in place
    else:
        shard = hi - lo
        dist.all_gather([flat_param[a + r * shard:a + (r + 1) * shard] for r in range(world_size)], flat_param[lo:hi].clone())
    return lo, hi


def sharded_update(flat_param, flat_grad, a, b, apply_fn, world_size=None, rank=None):
    """Update ``flat_param[a:b]`` from per-rank gradients ``flat_grad[a:b]``:

      reduce-scatter of the gradient  ->  ``apply_fn(lo, hi)`` on this rank's shard only (it must consume
      ``flat_grad[lo:hi]``, holding the SUM over ranks, and leave it zeroed)  ->  all-gather of the updated shard.

    Same bytes on the wire as an all-reduce, 1/world of the optimizer traffic, bit-identical parameters on every rank.
    The rest of ``flat_grad[a:b]`` is zeroed.  The two halves are separate functions so that a trainer can run the
    reduce-scatter of a finished slice of a table while the backward of the next slice is still running
    (``sanerf_b200/step.py: FusedSAMStep``)."""
    reduce_scatter_range(flat_grad, a, b, world_size, rank)
    return apply_and_gather_range(flat_param, flat_grad, a, b, apply_fn, world_size, rank)
