"""ORACLE / BASELINE (never on the product path): the reference's GPU path on this GPU.

The UNMODIFIED reference CUDA extension rebuilt for sm_100 (``oracle/_ref``, recipe ``oracle/build_ref.py``) driven exactly as
``gridencoder/grid.py:24-95`` drives it — ``[L,B,C]`` output + permute copy; permuted gradient copy + zero-filled gradient
table + scatter kernel — and, on top of it, the torch restatement of ``nerf/renderer.py`` + ``nerf/network.py``
(``oracle/render_torch.py``: torch glue kernels, cuBLAS ``nn.Linear`` MLPs, autograd, ``torch.optim.Adam``).  This is "the
real before" of SURVEY §8 d5: ``tests/test_gpu_speed_vs_reference.py`` asserts values and speed against it, and
``bench.py`` times it in the same run as a REPORTED baseline (``gpu_reference``), never as the thing measured.
"""
from __future__ import annotations

import numpy as np
import torch

from . import build_ref
from . import render_torch as R


def ref_forward(ext, x, table, offsets, S, H):
    """grid.py:27-69: [L,B,C] output, kernel, permute(1,0,2).reshape copy."""
    B, D = x.shape
    L, C = offsets.shape[0] - 1, table.shape[1]
    out = torch.empty(L, B, C, device=x.device, dtype=table.dtype)
    ext.grid_encode_forward(x, table, offsets, out, B, D, C, L, L, S, H, None, 0, False, 0)
    return out.permute(1, 0, 2).reshape(B, L * C)


def ref_backward(ext, grad, x, table, offsets, S, H):
    """grid.py:74-95: permuted contiguous gradient copy, zero-filled gradient table, scatter kernel."""
    B, D = x.shape
    L, C = offsets.shape[0] - 1, table.shape[1]
    g = grad.view(B, L, C).permute(1, 0, 2).contiguous()
    gt = torch.zeros_like(table)
    ext.grid_encode_backward(g, x, table, offsets, gt, B, D, C, L, L, S, H, None, None, 0, False, 0)
    return gt


class _RefGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, offsets, S, H, ext):
        ctx.save_for_backward(x, table, offsets)
        ctx.cfg = (S, H, ext)
        return ref_forward(ext, x, table, offsets, S, H)

    @staticmethod
    def backward(ctx, grad):
        x, table, offsets = ctx.saved_tensors
        S, H, ext = ctx.cfg
        return None, ref_backward(ext, grad.contiguous(), x, table, offsets, S, H), None, None, None, None


def ref_grid_cls(ext):
    class RefExtGridEncoder(R.GridEncoderRef):
        """GridEncoder of the reference (grid.py:102-168) on the rebuilt reference kernels."""

        def forward(self, inputs, bound=1, max_level=None):
            x = ((inputs + bound) / (2 * bound)).reshape(-1, self.input_dim).contiguous()
            out = _RefGridFn.apply(x, self.embeddings, self.offsets, float(np.log2(self.per_level_scale)),
                                   int(self.base_resolution), ext)
            return out.view(list(inputs.shape[:-1]) + [self.output_dim])

    return RefExtGridEncoder


def reference_rgb_step(device, lr=1e-2, ema_decay=None):
    """The stage-1 training step the reference runs on a GPU (nerf/utils.py:897-930, 1811-1836; main.py:296; the EMA of
    main.py:316 is updated once per epoch, nerf/utils.py:1862, so ``ema_decay`` is None unless a per-step variant is wanted):
    returns ``step(rays_o, rays_d, gt) -> None`` on a fresh random-init model, or raises FileNotFoundError when
    ``oracle/_ref`` is not built."""
    ext = build_ref.load("gridencoder")
    model = R.NeRFNetworkRef(grid_cls=ref_grid_cls(ext)).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-15)
    params = [p for p in model.parameters()]
    shadow = [p.detach().clone() for p in params] if ema_decay is not None else None
    state = {"t": 0}

    def step(o, d, rgb):
        opt.zero_grad(set_to_none=True)
        loss, _ = model.rgb_loss(o, d, rgb, update_proposal=True, perturb=True)
        loss.backward()
        opt.step()
        if shadow is not None:                       # torch_ema.ExponentialMovingAverage.update
            state["t"] += 1
            decay = min(ema_decay, (1 + state["t"]) / (10 + state["t"]))
            with torch.no_grad():
                torch._foreach_sub_(shadow, torch._foreach_mul(torch._foreach_sub(shadow, params), 1.0 - decay))

    return step
