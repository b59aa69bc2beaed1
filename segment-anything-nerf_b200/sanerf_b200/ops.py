"""Renderer-level operators on top of the C ABI: per-ray compositing and index dumps.

``composite`` is additive to the reference surface: it replaces the torch statements
of ``nerf/renderer.py:309-338, :377`` (sigma -> alpha -> exclusive-cumsum transmittance -> weights,
``weights_sum``, ``depth`` and the weighted channel sums) with one kernel per direction.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function

from . import _lib


def _f32c(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Composite(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, deltas, ts, feats, ray_offsets, max_count, last_sample_opaque, t_thresh):
        sigmas, deltas, ts = _f32c(sigmas, "sigmas"), _f32c(deltas, "deltas"), _f32c(ts, "ts")
        feats = _f32c(feats, "feats")
        dev = sigmas.device
        if ray_offsets is None:
            N, T = sigmas.shape
        else:
            if ray_offsets.dtype != torch.int32 or not ray_offsets.is_cuda:
                raise RuntimeError("ray_offsets must be a CUDA int32 tensor [N+1]")
            ray_offsets = ray_offsets.contiguous()
            N, T = ray_offsets.numel() - 1, int(max_count)
        C = 0 if feats is None else feats.shape[-1]
        weights = torch.empty_like(sigmas)
        weights_sum = torch.empty(N, device=dev, dtype=torch.float32)
        depth = torch.empty(N, device=dev, dtype=torch.float32)
        out = torch.empty(N, C, device=dev, dtype=torch.float32)
        n_alive = torch.empty(N, device=dev, dtype=torch.int32)
        lib = _lib.load()
        with torch.cuda.device(dev), _lib.stats.span("composite_forward", N=N, T=T, C=C):
            rc = lib.sanerf_composite_forward(
                sigmas.data_ptr(), deltas.data_ptr(), ts.data_ptr(), _lib.ptr(feats), 0, _lib.ptr(ray_offsets),
                N, T, C, int(bool(last_sample_opaque)), float(t_thresh), weights.data_ptr(),
                weights_sum.data_ptr(), depth.data_ptr(), out.data_ptr() if C else None, n_alive.data_ptr(),
                _lib.current_stream(dev))
        _lib.check(rc, "composite_forward")
        ctx.save_for_backward(sigmas, deltas, ts, feats, ray_offsets, weights)
        ctx.meta = (N, T, C, bool(last_sample_opaque), float(t_thresh))
        ctx.mark_non_differentiable(n_alive)
        return weights, weights_sum, depth, out, n_alive

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_weights, g_weights_sum, g_depth, g_out, _g_alive):
        sigmas, deltas, ts, feats, ray_offsets, weights = ctx.saved_tensors
        N, T, C, opaque, t_thresh = ctx.meta
        g_weights, g_weights_sum = _f32c(g_weights, "g_weights"), _f32c(g_weights_sum, "g_weights_sum")
        g_depth = _f32c(g_depth, "g_depth")
        g_out = _f32c(g_out, "g_out") if C else None
        grad_sigmas = torch.empty_like(sigmas)
        need_feats = C > 0 and ctx.needs_input_grad[3]
        grad_feats = torch.empty_like(feats) if need_feats else None
        lib = _lib.load()
        with torch.cuda.device(sigmas.device), _lib.stats.span("composite_backward", N=N, T=T, C=C):
            rc = lib.sanerf_composite_backward(
                sigmas.data_ptr(), deltas.data_ptr(), ts.data_ptr(), _lib.ptr(feats), 0, _lib.ptr(ray_offsets),
                N, T, C, int(opaque), t_thresh, weights.data_ptr(), _lib.ptr(g_weights), _lib.ptr(g_weights_sum),
                _lib.ptr(g_depth), _lib.ptr(g_out), grad_sigmas.data_ptr(), _lib.ptr(grad_feats), 0,
                _lib.current_stream(sigmas.device))
        _lib.check(rc, "composite_backward")
        return grad_sigmas, None, None, grad_feats, None, None, None, None


def composite(sigmas, deltas, ts, feats=None, ray_offsets=None, max_count=None, last_sample_opaque=True,
              t_thresh=0.0):
    """Front-to-back compositing of one per-sample channel block.

    Dense: ``sigmas, deltas, ts`` are ``[N, T]`` and ``feats`` ``[N, T, C]`` (or None).
    Packed: they are flat ``[M]`` / ``[M, C]`` with ``ray_offsets`` int32 ``[N+1]`` and ``max_count`` an upper
    bound on samples per ray.  Returns ``weights`` (same shape as sigmas), ``weights_sum [N]``,
    ``depth [N]``, ``out [N, C]`` and ``n_alive [N]`` (samples with T_i >= t_thresh).
    """
    if ray_offsets is not None and max_count is None:
        raise ValueError("packed composite needs max_count (upper bound of samples per ray)")
    return _Composite.apply(sigmas, deltas, ts, feats, ray_offsets, max_count, last_sample_opaque, t_thresh)


def grid_dump_indices(inputs, offsets, per_level_scale, base_resolution, gridtype=0, align_corners=False):
    """Debug / parity helper: table rows ``[B, L, 2^D]`` and device-side level geometry ``[L, 4]`` =
    (resolution, rows, hashed, covered_dims) exactly as the forward kernel computes them."""
    inputs = _f32c(inputs, "inputs")
    B, D = inputs.shape
    L = offsets.numel() - 1
    rows = torch.empty(B, L, 1 << D, device=inputs.device, dtype=torch.int32)
    geom = torch.empty(L, 4, device=inputs.device, dtype=torch.int32)
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_grid_dump_indices(inputs.data_ptr(), offsets.data_ptr(), rows.data_ptr(), geom.data_ptr(),
                                          B, D, L, float(np.log2(per_level_scale)), int(base_resolution),
                                          int(gridtype), int(bool(align_corners)), _lib.current_stream(inputs.device))
    _lib.check(rc, "grid_dump_indices")
    return rows, geom
