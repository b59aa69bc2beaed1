"""Large-scene sweep of BASELINE.json configs[4]: hash table L16 F2 T = 2^22 with HALF-precision parameters
(42.6 M rows, 170 MB — outside the 126 MB L2), 2^20 samples per GPU, at 1 / 2 / 4 / 8 GPUs.

One step = what a data-parallel training step does to such a table (SURVEY §5 "Distributed", §8 e2):

  gather    sanerf_grid_encode_forward   fp16 table -> [B, 32] fp16 encoding                 (588 B / sample)
  composite sanerf_composite_forward / _backward of C = 32 fp32 channels over 8192 rays x 128 samples
  scatter   sanerf_grid_encode_backward  [B, 32] fp16 gradient -> fp16 gradient table         (588 B / sample)
  update    N > 1: NCCL reduce-scatter of the fp16 gradient table (170 MB), Adam on this rank's 1/N shard with fp32 master
            weights and moments (sanerf_adam_step_half rewrites the half table in the same pass), all-gather of the updated
            half shard; N = 1: the same kernel over the whole table

The operators run back to back on static buffers (the encoder's half output feeds nothing here: the reference reaches half
tables only by calling the operator directly, main.py:222 forces autocast off); what the sweep reports is each kernel's
algorithmic-bytes fraction of the HBM peak and the whole step's samples/s, per GPU count.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, fused


class LargeSceneStep:
    def __init__(self, dev, world_size=1, rank=0, n_rays=8192, T=128, log2_hashmap_size=22, channels=32, lr=1e-2):
        from gridencoder import GridEncoder
        self.dev, self.world, self.rank = dev, int(world_size), int(rank)
        self.N, self.T, self.B, self.C = n_rays, T, n_rays * T, channels
        torch.manual_seed(0)                                   # identical replicas on every rank
        enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=log2_hashmap_size,
                          desired_resolution=4096).to(dev)
        self.offsets = enc.offsets
        self.L, self.S, self.H = 16, float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
        n = enc.embeddings.numel()
        self.n = n
        self.n_pad = (n + 511) // 512 * 512                    # 8 ranks x 8-element Adam vectors x 16-byte alignment
        f32 = dict(device=dev, dtype=torch.float32)
        self.master = torch.zeros(self.n_pad, **f32)
        self.master[:n].copy_(enc.embeddings.detach().reshape(-1))
        del enc
        self.table16 = self.master.half()
        self.grad16 = torch.zeros(self.n_pad, device=dev, dtype=torch.float16)
        self.exp_avg, self.exp_avg_sq = torch.zeros(self.n_pad, **f32), torch.zeros(self.n_pad, **f32)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.dyn = torch.tensor([lr, 1.0, 1.0, 0.0], **f32)
        self.lr = lr
        # ray-ordered sample positions (the coherence a real batch has): 8192 rays x 128 uniform samples
        g = torch.Generator().manual_seed(1 + rank)
        o = (torch.rand(n_rays, 3, generator=g) - 0.5).to(dev)
        d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).to(dev)
        aabb = torch.tensor([-128.0] * 3 + [128.0] * 3, **f32)
        noise = torch.rand(n_rays, T + 1, generator=g).to(dev)
        bins, t_mid, deltas, x01 = fused.sample_uniform(o, d, aabb, 0.2, T, noise)
        self.x01 = x01.reshape(-1, 3).contiguous()
        self.deltas, self.ts = deltas.contiguous(), t_mid.contiguous()
        self.enc16 = torch.empty(self.B, 32, device=dev, dtype=torch.float16)
        self.g_enc16 = (torch.randn(self.B, 32, generator=g) * 1e-3).to(dev).half()
        self.sigma = torch.randn(n_rays, T, generator=g).exp().to(dev)
        self.feats = torch.randn(n_rays, T, channels, generator=g).to(dev)
        self.weights, self.g_sigma = torch.empty(n_rays, T, **f32), torch.empty(n_rays, T, **f32)
        self.ws, self.depth = torch.empty(n_rays, **f32), torch.empty(n_rays, **f32)
        self.out, self.g_out = torch.empty(n_rays, channels, **f32), torch.randn(n_rays, channels, generator=g).to(dev)
        self.g_ws, self.g_depth = torch.randn(n_rays, generator=g).to(dev), torch.randn(n_rays, generator=g).to(dev)
        self.g_feats = torch.empty(n_rays, T, channels, **f32)
        self.alive = torch.empty(n_rays, device=dev, dtype=torch.int32)
        shard = self.n_pad // self.world
        self.lo, self.hi = self.rank * shard, (self.rank + 1) * shard

    # ------------------------------------------------------------------------------------------------
    def gather(self):
        lib, st = _lib.load(), _lib.current_stream(self.dev)
        with _lib.stats.span("grid_encode_forward", B=self.B, half=True):
            rc = lib.sanerf_grid_encode_forward(self.x01.data_ptr(), self.table16.data_ptr(), self.offsets.data_ptr(),
                                                self.enc16.data_ptr(), self.B, 3, 2, self.L, self.L, self.S, self.H, None, 0, 0, 0,
                                                _lib.SANERF_F16, _lib.LAYOUT_BLC, 0, st)
        _lib.check(rc, "grid_encode_forward")

    def composite(self):
        lib, st = _lib.load(), _lib.current_stream(self.dev)
        N, T, C = self.N, self.T, self.C
        with _lib.stats.span("composite_forward", N=N, T=T, C=C):
            rc = lib.sanerf_composite_forward(self.sigma.data_ptr(), self.deltas.data_ptr(), self.ts.data_ptr(), self.feats.data_ptr(),
                                              0, None, N, T, C, 1, 0.0, self.weights.data_ptr(), self.ws.data_ptr(),
                                              self.depth.data_ptr(), self.out.data_ptr(), self.alive.data_ptr(), st)
        _lib.check(rc, "composite_forward")
        with _lib.stats.span("composite_backward", N=N, T=T, C=C):
            rc = lib.sanerf_composite_backward(self.sigma.data_ptr(), self.deltas.data_ptr(), self.ts.data_ptr(), self.feats.data_ptr(),
                                               0, None, N, T, C, 1, 0.0, self.weights.data_ptr(), None, self.g_ws.data_ptr(),
                                               self.g_depth.data_ptr(), self.g_out.data_ptr(), self.g_sigma.data_ptr(),
                                               self.g_feats.data_ptr(), 0, st)
        _lib.check(rc, "composite_backward")

    def scatter(self):
        lib, st = _lib.load(), _lib.current_stream(self.dev)
        with _lib.stats.span("grid_encode_backward", B=self.B, half=True):
            rc = lib.sanerf_grid_encode_backward(self.g_enc16.data_ptr(), self.x01.data_ptr(), self.table16.data_ptr(),
                                                 self.offsets.data_ptr(), self.grad16.data_ptr(), self.B, 3, 2, self.L, self.L, self.S,
                                                 self.H, None, None, 0, 0, 0, _lib.SANERF_F16, _lib.LAYOUT_BLC, st)
        _lib.check(rc, "grid_encode_backward")

    def update(self):
        lib, st = _lib.load(), _lib.current_stream(self.dev)
        with _lib.stats.span("adam_schedule"):
            _lib.check(lib.sanerf_adam_schedule(self.step_count.data_ptr(), self.dyn.data_ptr(), self.lr, 0.9, 0.999, 20000.0,
                                                None, 0.0, st), "adam_schedule")
        lo, hi = (0, self.n_pad) if self.world == 1 else (self.lo, self.hi)
        if self.world > 1:
            dist.reduce_scatter_tensor(self.grad16[lo:hi], self.grad16, op=dist.ReduceOp.SUM)      # in place, fp16 on the wire
        with _lib.stats.span("adam_step_half", n=hi - lo):
            rc = lib.sanerf_adam_step_half(self.master.data_ptr() + 4 * lo, self.table16.data_ptr() + 2 * lo,
                                           self.grad16.data_ptr() + 2 * lo, self.exp_avg.data_ptr() + 4 * lo,
                                           self.exp_avg_sq.data_ptr() + 4 * lo, hi - lo, self.dyn.data_ptr(), 0.9, 0.999, 1e-15,
                                           1.0 / self.world, 1, st)
        _lib.check(rc, "adam_step_half")
        if self.world > 1:
            if lo > 0:
                self.grad16[:lo].zero_()
            if hi < self.n_pad:
                self.grad16[hi:].zero_()
            dist.all_gather_into_tensor(self.table16, self.table16[lo:hi])

    def __call__(self, i=0):
        with torch.cuda.device(self.dev):
            self.gather()
            self.composite()
            self.scatter()
            self.update()

    # ------------------------------------------------------------------------------------------------
    def algorithmic_bytes(self):
        N, T, C, B = self.N, self.T, self.C, self.B
        fwd = N * (T * (12 + 4 * C) + 4 * (C + 2))
        shard = (self.hi - self.lo) if self.world > 1 else self.n_pad
        return {"grid_encode_forward": B * (12 + 16 * 8 * 2 * 2 + 16 * 2 * 2),        # 588 B / sample
                "grid_encode_backward": B * (12 + 16 * 8 * 2 * 2 + 16 * 2 * 2),
                "composite_forward": fwd, "composite_backward": fwd + N * 4 * (C + 2) + N * T * (4 + 4 * C),
                "adam_step_half": shard * 30}                                          # p, m, v rd+wr fp32; g rd + zero, p16 wr

    def kernel_times(self, ctx, n=5):
        """Eager replays with CUDA events around every C-ABI launch (L2 flushed between steps): mean ms and the
        algorithmic-bytes fraction of the HBM peak per kernel."""
        ctx.timed(self, n, "*", None)
        acc = {}
        for ms, info in _lib.stats.durations_ms():
            c, s = acc.get(info["name"], (0, 0.0))
            acc[info["name"]] = (c + 1, s + ms)
        alg = self.algorithmic_bytes()
        out = {}
        for name, (c, s) in acc.items():
            if name in alg:
                ms = s / c
                out[name] = {"avg_ms": ms, "algorithmic_bytes": alg[name], "frac": alg[name] / (ms * 1e-3) / 1e9 / ctx.peak}
        return out
