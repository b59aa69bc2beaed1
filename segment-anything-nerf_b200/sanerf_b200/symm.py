"""Symmetric-memory plumbing of the fused data-parallel update (csrc/symm_adam.cu).

``torch.distributed._symmetric_memory`` is used for what it is — an allocator that maps the same buffer of every rank into
every rank's address space (unicast peer addresses over NVLink, plus one multicast / NVLS address when the fabric supports
it) and a rendezvous; the exchange itself is this repository's kernel.  The flat parameter and gradient buffers of
``FusedAdam`` are allocated here when ``world_size > 1``, so every gradient kernel of the step writes straight into memory
the peers (or the switch) can read.

``SANERF_SYMM=0`` keeps ordinary allocations and the NCCL exchange (reduce-scatter -> Adam shard -> all-gather /
all-reduce -> Adam), which is also the checker of the fused kernel (tools/check_ddp.py, bench ``grad_equiv``).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

MAX_BLOCKS, MAX_WORLD = 256, 8
CHANNEL_BLOCKS = {0: 160, 1: 32, 2: 32}


def enabled(world_size):
    return world_size in (2, 4, 8) and os.environ.get("SANERF_SYMM", "1") != "0" and dist.is_initialized() \
        and dist.get_backend() == "nccl"


class SymmetricState:
    """Symmetric flat parameter / gradient buffers of ``n`` floats + the barrier flags, and their peer / multicast addresses."""

    def __init__(self, n, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world not in (2, 4, 8):
            raise RuntimeError("the fused symmetric-memory update supports 2, 4 or 8 ranks of one node")
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:  # noqa: BLE001  (newer torch enables every group implicitly)
            pass
        self.param = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.grad = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(MAX_BLOCKS * MAX_WORLD, dtype=torch.int32, device=device)
        self.param.zero_(); self.grad.zero_(); self.flags.zero_()
        hp = symm_mem.rendezvous(self.param, group)
        hg = symm_mem.rendezvous(self.grad, group)
        hf = symm_mem.rendezvous(self.flags, group)
        self._handles = (hp, hg, hf)                           # keep the mappings alive
        use_mc = os.environ.get("SANERF_SYMM_MULTICAST", "1") != "0"
        self.param_mc = int(hp.multicast_ptr) if (use_mc and hp.has_multicast_support) else 0
        self.grad_mc = int(hg.multicast_ptr) if (use_mc and hg.has_multicast_support) else 0
        # gradient reduction: in-switch (multimem.ld_reduce, NVLS) by default; SANERF_SYMM_REDUCE=p2p reads every peer's
        # slice with plain loads instead.  8 GPUs, RGB step, ms/step: multimem 0.931, p2p 0.977, NCCL exchange 1.050
        # (profiles/r2_symm_sweep.md); at 2 GPUs the two are equal within noise.
        if os.environ.get("SANERF_SYMM_REDUCE", "multimem") != "multimem":
            self.grad_mc = 0
        arr = ctypes.c_uint64 * self.world
        self.param_peers = arr(*[int(p) for p in hp.buffer_ptrs])
        self.grad_peers = arr(*[int(p) for p in hg.buffer_ptrs])
        self.flag_peers = arr(*[int(p) for p in hf.buffer_ptrs])
        assert int(hp.buffer_ptrs[self.rank]) == self.param.data_ptr() and int(hg.buffer_ptrs[self.rank]) == self.grad.data_ptr()
        self.epoch = torch.zeros(MAX_BLOCKS, device=device, dtype=torch.int32)
        self.error = torch.zeros(1, device=device, dtype=torch.int32)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                    # every rank's buffers are zeroed before anyone signals

    def multicast(self):
        return bool(self.param_mc)

    def describe(self):
        return (f"gradient reduction: {'multimem.ld_reduce (NVLS)' if self.grad_mc else 'peer loads over NVLink'}; parameter "
                f"broadcast: {'multimem.st (NVLS)' if self.param_mc else 'peer stores'}")

    def launch(self, opt, start, stop, grad_scale, gated, blocks, threads=256, channel=0):
        """Fused reduce + Adam(+EMA) + broadcast of the flat range [start, stop) on the current stream."""
        lib = _lib.load()
        dev = opt.flat_param.device
        with torch.cuda.device(dev), _lib.stats.span("symm_adam_step", n=stop - start):
            rc = lib.sanerf_symm_adam_step(
                opt.flat_param.data_ptr(), opt.flat_grad.data_ptr(), opt.exp_avg.data_ptr(), opt.exp_avg_sq.data_ptr(),
                opt.ema.data_ptr() if opt.ema_every_step else None, self.param_mc or None, self.grad_mc or None,
                self.param_peers, self.grad_peers, self.flag_peers, self.epoch.data_ptr(), self.error.data_ptr(),
                int(start), int(stop), self.world, self.rank, opt.dyn.data_ptr(), opt.betas[0], opt.betas[1], opt.eps,
                float(grad_scale), opt.gate.data_ptr() if gated else None, int(blocks), int(threads), int(channel), _lib.current_stream(dev))
        _lib.check(rc, "symm_adam_step")

    def check(self):
        """Raise if a cross-rank barrier of the fused update ever timed out (synchronises; call at checkpoints)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("fused symmetric-memory update: a peer rank never reached the barrier")


def slice_bounds(start, stop, world, rank):
    """The slice of [start, stop) rank ``rank`` reduces, updates and broadcasts (same arithmetic as symm_adam_kernel)."""
    n4 = (stop - start) // 4
    per = (n4 + world - 1) // world
    end = start // 4 + n4
    lo = min(start // 4 + rank * per, end)              # ranks past the end of a short range own an empty slice
    hi = min(lo + per, end)
    return 4 * lo, 4 * hi
