"""Multi-process host logic on CPU (gloo, world size 2): ray / tile sharding, the frame gather and the sharded
optimizer-update protocol of sanerf_b200/parallel.py (on the GPUs the same functions run over NCCL; the kernels
themselves are covered by the -m gpu tests and tools/check_ddp.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sanerf_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _spawn(fn, *args, world=2):
    mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=True)


@pytest.mark.parametrize("n,world", [(10, 2), (11, 2), (262144, 8), (7, 8), (5, 3)])
def test_shard_rays_partitions_exactly(n, world):
    spans = [parallel.shard_rays(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and 0 <= (b - a) - (d - c) <= 1           # contiguous, sizes differ by at most one, larger first


def _gather_case(rank, world, n_total):
    full = torch.arange(n_total * 3, dtype=torch.float32).view(n_total, 3)
    a, b = parallel.shard_rays(n_total, rank, world)
    out = parallel.gather_frame(full[a:b].clone(), n_total, rank, world)
    assert torch.equal(out, full)


@pytest.mark.parametrize("n_total", [10, 11])
def test_gather_frame_stitches_uneven_tiles(n_total):
    _spawn(_gather_case, n_total)


def _sharded_update_case(rank, world):
    """reduce(-scatter) + update of the local shard + all-gather == all-reduce + full update, on every rank."""
    torch.manual_seed(0)
    n, a, b = 64 * 5, 64, 64 * 4                       # a flat buffer with untouched slots before and after the range
    param0 = torch.randn(n)
    grads = [torch.randn(n, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    lr = 0.1

    def sgd(param, grad):
        def apply(lo, hi):
            param[lo:hi] -= lr * grad[lo:hi] / world   # consumes the SUM over ranks ...
            grad[lo:hi].zero_()                        # ... and leaves the gradient cleared (like sanerf_adam_step)
        return apply

    param, grad = param0.clone(), grads[rank].clone()
    lo, hi = parallel.sharded_update(param, grad, a, b, sgd(param, grad), world, rank)
    assert (lo, hi) == parallel.shard_bounds(a, b, world, rank) and (hi - lo) * world == b - a
    expect = param0.clone()
    expect[a:b] -= lr * sum(g[a:b] for g in grads) / world
    assert torch.equal(param, expect)                                     # exact at world = 2 (one summation order)
    assert float(grad[a:b].abs().max()) == 0.0                            # whole range cleared
    assert torch.equal(grad[:a], grads[rank][:a]) and torch.equal(grad[b:], grads[rank][b:])   # neighbours untouched
    ref = param.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, param)                                        # ranks bit-identical


def test_sharded_update_equals_allreduce_update():
    _spawn(_sharded_update_case)


def _sliced_update_case(rank, world):
    """The two halves used separately, slice by slice (a trainer reduces a finished slice of a table while the next
    slice's gradient is still being written): reduce every slice first, update + gather later == one all-reduce update."""
    torch.manual_seed(1)
    n = 64 * 6
    slices = [(0, 64), (64, 64 * 3), (64 * 3, 64 * 6)]
    param0 = torch.randn(n)
    grads = [torch.randn(n, generator=torch.Generator().manual_seed(20 + r)) for r in range(world)]
    lr = 0.05
    param, grad = param0.clone(), torch.zeros(n)

    def apply(lo, hi):
        param[lo:hi] -= lr * grad[lo:hi] / world
        grad[lo:hi].zero_()

    for a, b in slices:                                  # the slice's gradient "arrives", then its reduction starts
        grad[a:b] = grads[rank][a:b]
        lo, hi = parallel.reduce_scatter_range(grad, a, b, world, rank)
        assert (lo, hi) == parallel.shard_bounds(a, b, world, rank)
        assert torch.equal(grad[lo:hi], sum(g[lo:hi] for g in grads))      # the local shard holds the sum over ranks
    for a, b in slices:
        parallel.apply_and_gather_range(param, grad, a, b, apply, world, rank)
    expect = param0 - lr * sum(grads) / world
    assert torch.equal(param, expect) and float(grad.abs().max()) == 0.0
    ref = param.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, param)


def test_sliced_reduce_then_update_equals_allreduce_update():
    _spawn(_sliced_update_case)


def test_shard_bounds_reject_unaligned_ranges():
    assert parallel.shard_bounds(0, 12_599_936, 8, 3) == (3 * 1_574_992, 4 * 1_574_992)
    with pytest.raises(ValueError):
        parallel.shard_bounds(0, 12_599_920, 8, 0)      # the unpadded main-table size does not split into aligned shards


def test_flat_bucket_layout_on_cpu():
    """FusedAdam's flat buffers (no kernel involved): slots are multiples of 32 elements, parameters and gradients are
    views, ranges are recorded in declaration order."""
    from sanerf_b200.fused import FusedAdam
    ps = [torch.nn.Parameter(torch.randn(12_599_920 // 1000, 2)), torch.nn.Parameter(torch.randn(64, 32)),
          torch.nn.Parameter(torch.randn(3, 32))]
    before = [p.detach().clone() for p in ps]
    opt = FusedAdam(ps)
    off = 0
    for p, b in zip(ps, before):
        lo, hi = opt.ranges[id(p)]
        assert lo == off and (hi - lo) % 32 == 0 and hi - lo >= p.numel()
        assert torch.equal(p.detach(), b) and p.data_ptr() == opt.flat_param[lo:].data_ptr()
        assert p.grad.data_ptr() == opt.flat_grad[lo:].data_ptr() and p.grad.shape == p.shape
        off = hi
    assert opt.flat_param.numel() == off
    for world in (2, 4, 8):
        lo, hi = opt.ranges[id(ps[0])]
        parallel.shard_bounds(lo, hi, world, world - 1)
