"""Fused operators of the training / rendering step (additive to the reference surface).

Each one collapses a chain of torch kernels of ``nerf/renderer.py`` / ``nerf/network.py`` / the Trainer into one
or two sm_100a kernels behind the C ABI:

* ``sample_uniform`` / ``sample_pdf``   — renderer.py:122-139, 250-286, 60-69, 84-119 (one kernel per level)
* ``prop_density``                      — network.py:248-252: grid encode -> MLP(2L,16,1) -> trunc_exp
* ``head_composite``                    — network.py:226-227 + renderer.py:309-338: trunc_exp on column 0 of the
                                          16-wide MLP output and compositing of columns 1..15, read in place
* ``ray_features``                      — renderer.py:302-303, 377: s_grid encode + weighted sum over the ray, and its
                                          scatter, without the per-sample [N*T, 128] feature / gradient matrices
* ``proposal_loss`` / ``distort_loss``  — renderer.py:17-57 (loss and d loss / d weights together)
* ``FusedAdam``                         — main.py:296,312-313: Adam(eps=1e-15) + LambdaLR on flat buffers
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function

from . import _lib


def _stream(t):
    return _lib.current_stream(t.device)


# ----------------------------------------------------------------------------------------- sampling
def _sample_outputs(N, T, dev):
    return (torch.empty(N, T + 1, device=dev), torch.empty(N, T, device=dev), torch.empty(N, T, device=dev),
            torch.empty(N, T, 3, device=dev))


def _cnf(cam_near_far):
    if cam_near_far is None:
        return None, 0
    c = cam_near_far.contiguous().float()
    return c, (0 if c.shape[0] == 1 else 2)


@torch.no_grad()
def sample_uniform(rays_o, rays_d, aabb, min_near, T, noise=None, cam_near_far=None, contract=True, bound=2.0):
    """Level-0 sampling: returns bins [N,T+1], t_mid [N,T], deltas [N,T], x01 [N,T,3] (already in the grid's unit cube)."""
    N, dev = rays_o.shape[0], rays_o.device
    bins, t_mid, deltas, x01 = _sample_outputs(N, T, dev)
    cnf, stride = _cnf(cam_near_far)
    lib = _lib.load()
    with torch.cuda.device(dev), _lib.stats.span("sample_uniform", N=N, T=T):
        rc = lib.sanerf_sample_uniform(rays_o.data_ptr(), rays_d.data_ptr(), aabb.data_ptr(), float(min_near),
                                       _lib.ptr(cnf), stride, _lib.ptr(noise), N, T, int(bool(contract)), float(bound),
                                       bins.data_ptr(), t_mid.data_ptr(), deltas.data_ptr(), x01.data_ptr(),
                                       _stream(rays_o))
    _lib.check(rc, "sample_uniform")
    return bins, t_mid, deltas, x01


@torch.no_grad()
def sample_pdf(rays_o, rays_d, aabb, min_near, prev_bins, prev_weights, T, noise=None, cam_near_far=None,
               contract=True, bound=2.0):
    """Inverse-CDF resampling of T+1 edges from the previous level, plus the same outputs as ``sample_uniform``."""
    N, dev = rays_o.shape[0], rays_o.device
    T0 = prev_weights.shape[1]
    bins, t_mid, deltas, x01 = _sample_outputs(N, T, dev)
    cnf, stride = _cnf(cam_near_far)
    prev_bins, prev_weights = prev_bins.contiguous(), prev_weights.detach().contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev), _lib.stats.span("sample_pdf", N=N, T=T):
        rc = lib.sanerf_sample_pdf(rays_o.data_ptr(), rays_d.data_ptr(), aabb.data_ptr(), float(min_near),
                                   _lib.ptr(cnf), stride, prev_bins.data_ptr(), prev_weights.data_ptr(), T0,
                                   _lib.ptr(noise), N, T, int(bool(contract)), float(bound), bins.data_ptr(),
                                   t_mid.data_ptr(), deltas.data_ptr(), x01.data_ptr(), None, None, 0, None, _stream(rays_o))
    _lib.check(rc, "sample_pdf")
    return bins, t_mid, deltas, x01


# ----------------------------------------------------------------------------------------- proposal density
class _PropDensity(Function):
    @staticmethod
    def forward(ctx, x01, table, offsets, w1, w2, S, H):
        x01 = x01.contiguous()
        B = x01.numel() // 3
        L = offsets.numel() - 1
        sigma = torch.empty(x01.shape[:-1], device=x01.device, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad[1:5])
        enc = torch.empty(B, 2 * L, device=x01.device, dtype=torch.float32) if need_grad else None
        lib = _lib.load()
        with torch.cuda.device(x01.device), _lib.stats.span("prop_density_forward", B=B, L=L):
            rc = lib.sanerf_prop_density_forward(x01.data_ptr(), table.data_ptr(), offsets.data_ptr(), w1.data_ptr(),
                                                 w2.data_ptr(), B, L, float(S), int(H), sigma.data_ptr(), _lib.ptr(enc),
                                                 _stream(x01))
        _lib.check(rc, "prop_density_forward")
        ctx.save_for_backward(x01, table, offsets, w1, w2, enc)
        ctx.meta = (B, L, float(S), int(H))
        return sigma

    @staticmethod
    def backward(ctx, g_sigma):
        x01, table, offsets, w1, w2, enc = ctx.saved_tensors
        B, L, S, H = ctx.meta
        g_sigma = g_sigma.contiguous()
        g_table, g_w1, g_w2 = torch.zeros_like(table), torch.zeros_like(w1), torch.zeros_like(w2)
        lib = _lib.load()
        with torch.cuda.device(x01.device), _lib.stats.span("prop_density_backward", B=B, L=L):
            rc = lib.sanerf_prop_density_backward(x01.data_ptr(), table.data_ptr(), offsets.data_ptr(), w1.data_ptr(),
                                                  w2.data_ptr(), B, L, S, H, _lib.ptr(enc), g_sigma.data_ptr(),
                                                  g_table.data_ptr(), g_w1.data_ptr(), g_w2.data_ptr(), _stream(x01))
        _lib.check(rc, "prop_density_backward")
        return None, g_table, None, g_w1, g_w2, None, None


def prop_density_supported(encoder, mlp):
    """The fused kernel covers the proposal networks the reference builds (network.py:211-219)."""
    net = getattr(mlp, "net", None)
    return (encoder.input_dim == 3 and encoder.level_dim == 2 and encoder.num_levels <= 8
            and encoder.gridtype_id == 0 and encoder.interp_id == 0 and not encoder.align_corners
            and encoder.embeddings.dtype == torch.float32 and net is not None and len(net) == 2
            and net[0].bias is None and net[1].bias is None and net[0].out_features == 16
            and net[0].in_features == 2 * encoder.num_levels and net[1].out_features == 1
            and not torch.is_autocast_enabled())


def prop_density(x01, encoder, mlp):
    """sigma = trunc_exp(mlp(encoder(x))) for positions already mapped to [0,1]^3."""
    return _PropDensity.apply(x01, encoder.embeddings, encoder.offsets, mlp.net[0].weight, mlp.net[1].weight,
                              float(np.log2(encoder.per_level_scale)), int(encoder.base_resolution))


# ----------------------------------------------------------------------------------------- field head (tensor cores)
PRECISION_IDS = {"fp32": 0, "tf32": 1}


# The activations the field-head forward keeps for its backward (encoding, H1, H2) are "tile-chunk-major":
# [tile of 128 samples][4-element chunk][sample][4], so that both kernels touch 512 contiguous bytes per warp
# instruction.  They are private to the two kernels; these helpers exist for tests and diagnostics.
def tcm_rows(B):
    return (B + 127) // 128 * 128


def tcm_to_rows(t, B, W):
    tiles = t.numel() // (128 * W)
    return t.view(tiles, W // 4, 128, 4).permute(0, 2, 1, 3).reshape(tiles * 128, W)[:B]


def rows_to_tcm(x):
    B, W = x.shape
    Bp = tcm_rows(B)
    pad = x.new_zeros(Bp, W)
    pad[:B] = x
    return pad.view(Bp // 128, 128, W // 4, 4).permute(0, 2, 1, 3).contiguous().view(Bp, W)


class _FieldHead(Function):
    """out [.., 16] = grid_mlp(grid(x)) in ONE tcgen05 kernel; backward = one tcgen05 kernel (data + weight
    gradients of the MLP) + the hash-grid scatter."""

    @staticmethod
    def forward(ctx, x01, table, offsets, w1, w2, w3, S, H, precision):
        x01 = x01.contiguous()
        B = x01.numel() // 3
        dev = x01.device
        need_grad = any(ctx.needs_input_grad[1:6])
        out = torch.empty(*x01.shape[:-1], 16, device=dev, dtype=torch.float32)
        Bp = tcm_rows(B)                       # saved activations are tile-chunk-major: whole 128-sample tiles
        enc = torch.empty(Bp, 32, device=dev, dtype=torch.float32) if need_grad else None
        h1 = torch.empty(Bp, 64, device=dev, dtype=torch.float32) if need_grad else None
        h2 = torch.empty(Bp, 64, device=dev, dtype=torch.float32) if need_grad else None
        w1, w2, w3 = w1.contiguous(), w2.contiguous(), w3.contiguous()
        lib = _lib.load()
        with torch.cuda.device(dev), _lib.stats.span("field_head_forward", B=B):
            rc = lib.sanerf_field_head_forward(x01.data_ptr(), table.data_ptr(), offsets.data_ptr(), float(S), int(H),
                                               None, w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, _lib.ptr(enc),
                                               _lib.ptr(h1), _lib.ptr(h2), out.data_ptr(), int(precision), _stream(x01))
        _lib.check(rc, "field_head_forward")
        if need_grad:
            ctx.save_for_backward(x01, table, offsets, w1, w2, w3, enc, h1, h2)
            ctx.meta = (B, float(S), int(H), int(precision))
        return out

    @staticmethod
    def backward(ctx, g_out):
        x01, table, offsets, w1, w2, w3, enc, h1, h2 = ctx.saved_tensors
        B, S, H, precision = ctx.meta
        dev = x01.device
        g_out = g_out.contiguous()
        g_w1, g_w2, g_w3 = torch.zeros_like(w1), torch.zeros_like(w2), torch.zeros_like(w3)
        lib = _lib.load()
        st = _stream(x01)
        g_enc = torch.empty(B, 32, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            with _lib.stats.span("field_head_backward", B=B):
                rc = lib.sanerf_field_head_backward(enc.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(),
                                                    w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, g_enc.data_ptr(), None,
                                                    None, 0.0, 0, None, g_w1.data_ptr(), g_w2.data_ptr(), g_w3.data_ptr(),
                                                    precision, st)
            _lib.check(rc, "field_head_backward")
            g_table = None
            if ctx.needs_input_grad[1]:
                g_table = torch.zeros_like(table)
                with _lib.stats.span("grid_encode_backward", B=B, L=16, C=2, D=3, half=False):
                    rc = lib.sanerf_grid_encode_backward(g_enc.data_ptr(), x01.data_ptr(), table.data_ptr(),
                                                         offsets.data_ptr(), g_table.data_ptr(), B, 3, 2, 16, 16, S, H,
                                                         None, None, 0, 0, 0, _lib.SANERF_F32, _lib.LAYOUT_BLC, st)
                _lib.check(rc, "grid_encode_backward")
        return None, g_table, None, g_w1, g_w2, g_w3, None, None, None


def field_head_supported(encoder, mlp):
    """The fused kernel covers the main grid + grid_mlp the reference builds (network.py:102-103)."""
    net = getattr(mlp, "net", None)
    return (encoder.input_dim == 3 and encoder.level_dim == 2 and encoder.num_levels == 16
            and encoder.gridtype_id == 0 and encoder.interp_id == 0 and not encoder.align_corners
            and encoder.embeddings.dtype == torch.float32 and net is not None and len(net) == 3
            and all(l.bias is None for l in net) and tuple(net[0].weight.shape) == (64, 32)
            and tuple(net[1].weight.shape) == (64, 64) and tuple(net[2].weight.shape) == (16, 64)
            and not torch.is_autocast_enabled())


def field_head(x01, encoder, mlp, precision="fp32"):
    """[.., 3] positions in [0,1]^3 -> [.., 16] = mlp(encoder(x)) (column 0: density logit, 1..15: geometry feature)."""
    return _FieldHead.apply(x01, encoder.embeddings, encoder.offsets, mlp.net[0].weight, mlp.net[1].weight,
                            mlp.net[2].weight, float(np.log2(encoder.per_level_scale)), int(encoder.base_resolution),
                            PRECISION_IDS[precision])


# ----------------------------------------------------------------------------------------- head + composite
class _HeadComposite(Function):
    """f [N,T,W] = MLP output: column 0 -> sigma = exp(.), columns 1..W-1 composited in place (row stride W)."""

    @staticmethod
    def forward(ctx, f, deltas, ts, last_sample_opaque, t_thresh):
        f = f.contiguous()
        N, T, W = f.shape
        C = W - 1
        dev = f.device
        deltas, ts = deltas.contiguous(), ts.contiguous()
        sigma = torch.empty(N, T, device=dev)
        weights = torch.empty(N, T, device=dev)
        weights_sum, depth = torch.empty(N, device=dev), torch.empty(N, device=dev)
        out = torch.empty(N, C, device=dev)
        n_alive = torch.empty(N, device=dev, dtype=torch.int32)
        lib = _lib.load()
        st = _stream(f)
        ctx.save_for_backward(f, sigma, deltas, ts, weights)
        ctx.meta = (N, T, W, bool(last_sample_opaque), float(t_thresh))
        ctx.mark_non_differentiable(n_alive)
        if W == 16:        # the reference's head (1 logit + 15 features): one kernel, a lane owns a sample's whole row
            with torch.cuda.device(dev), _lib.stats.span("head_composite_forward", N=N, T=T):
                rc = lib.sanerf_head_composite_forward(f.data_ptr(), deltas.data_ptr(), ts.data_ptr(), N, T,
                                                       int(bool(last_sample_opaque)), float(t_thresh), sigma.data_ptr(),
                                                       weights.data_ptr(), weights_sum.data_ptr(), depth.data_ptr(),
                                                       out.data_ptr(), n_alive.data_ptr(), st)
            _lib.check(rc, "head_composite_forward")
            return sigma, weights, weights_sum, depth, out, n_alive
        with torch.cuda.device(dev):
            with _lib.stats.span("trunc_exp_forward", n=N * T):
                rc = lib.sanerf_trunc_exp_forward(f.data_ptr(), sigma.data_ptr(), N * T, W, 0, st)
            _lib.check(rc, "trunc_exp_forward")
            with _lib.stats.span("composite_forward", N=N, T=T, C=C):
                rc = lib.sanerf_composite_forward(sigma.data_ptr(), deltas.data_ptr(), ts.data_ptr(),
                                                  f.data_ptr() + 4, W, None, N, T, C, int(bool(last_sample_opaque)),
                                                  float(t_thresh), weights.data_ptr(), weights_sum.data_ptr(),
                                                  depth.data_ptr(), out.data_ptr(), n_alive.data_ptr(), st)
            _lib.check(rc, "composite_forward")
        return sigma, weights, weights_sum, depth, out, n_alive

    @staticmethod
    def backward(ctx, g_sigma_direct, g_weights, g_weights_sum, g_depth, g_out, _g_alive):
        f, sigma, deltas, ts, weights = ctx.saved_tensors
        N, T, W, opaque, t_thresh = ctx.meta
        C = W - 1
        dev = f.device
        cont = lambda t: None if t is None else t.contiguous()  # noqa: E731
        g_weights, g_weights_sum, g_depth, g_out = cont(g_weights), cont(g_weights_sum), cont(g_depth), cont(g_out)
        grad_f = torch.empty_like(f)
        lib = _lib.load()
        st = _stream(f)
        if W == 16:
            with torch.cuda.device(dev), _lib.stats.span("head_composite_backward", N=N, T=T):
                rc = lib.sanerf_head_composite_backward(f.data_ptr(), deltas.data_ptr(), ts.data_ptr(), N, T, int(opaque),
                                                        t_thresh, _lib.ptr(g_weights), _lib.ptr(g_weights_sum),
                                                        _lib.ptr(g_depth), _lib.ptr(g_out), _lib.ptr(cont(g_sigma_direct)),
                                                        grad_f.data_ptr(), st)
            _lib.check(rc, "head_composite_backward")
            return grad_f, None, None, None, None
        grad_sigma = torch.empty(N, T, device=dev)
        with torch.cuda.device(dev):
            with _lib.stats.span("composite_backward", N=N, T=T, C=C):
                rc = lib.sanerf_composite_backward(sigma.data_ptr(), deltas.data_ptr(), ts.data_ptr(), f.data_ptr() + 4,
                                                   W, None, N, T, C, int(opaque), t_thresh, weights.data_ptr(),
                                                   _lib.ptr(g_weights), _lib.ptr(g_weights_sum), _lib.ptr(g_depth),
                                                   _lib.ptr(g_out), grad_sigma.data_ptr(), grad_f.data_ptr() + 4, W, st)
            _lib.check(rc, "composite_backward")
            if g_sigma_direct is not None:
                grad_sigma = grad_sigma + g_sigma_direct
            with _lib.stats.span("trunc_exp_backward", n=N * T):
                rc = lib.sanerf_trunc_exp_backward(grad_sigma.data_ptr(), f.data_ptr(), grad_f.data_ptr(), N * T, W, 0, st)
            _lib.check(rc, "trunc_exp_backward")
        return grad_f, None, None, None, None


def head_composite(f, deltas, ts, last_sample_opaque=True, t_thresh=0.0):
    """Returns sigma [N,T], weights [N,T], weights_sum [N], depth [N], out [N,W-1], n_alive [N]."""
    return _HeadComposite.apply(f, deltas, ts, last_sample_opaque, t_thresh)


# ----------------------------------------------------------------------------------------- losses
class _RayFeatures(Function):
    """f_sam[r] = sum_i w[r,i] * s_grid(x[r,i])  (renderer.py:302-303, 377) without the per-sample feature matrix; the
    weights are constants (the density field is frozen in stage 2), the table receives the scatter."""

    @staticmethod
    def forward(ctx, x01, weights, table, offsets, S, H):
        N, T = weights.shape
        L, C = offsets.numel() - 1, table.shape[1]
        x01 = x01.reshape(N * T, 3).contiguous().float()
        weights = weights.detach().contiguous().float()
        out = torch.empty(N, L * C, device=table.device, dtype=torch.float32)
        with _lib.stats.span("ray_features_forward", N=N, T=T, C=C):
            rc = _lib.load().sanerf_ray_features_forward(x01.data_ptr(), weights.data_ptr(), table.data_ptr(),
                                                         offsets.data_ptr(), N, T, C, L, S, H, out.data_ptr(), 0, _stream(table))
        _lib.check(rc, "ray_features_forward")
        ctx.save_for_backward(x01, weights, offsets)
        ctx.meta = (N, T, C, L, S, H, tuple(table.shape))
        return out

    @staticmethod
    def backward(ctx, g_out):
        x01, weights, offsets = ctx.saved_tensors
        N, T, C, L, S, H, shape = ctx.meta
        g_table = torch.zeros(shape, device=g_out.device, dtype=torch.float32)
        g_out = g_out.contiguous().float()
        with _lib.stats.span("ray_features_backward", N=N, T=T, C=C):
            rc = _lib.load().sanerf_ray_features_backward(x01.data_ptr(), weights.data_ptr(), g_out.data_ptr(),
                                                          offsets.data_ptr(), N, T, C, L, S, H, g_table.data_ptr(), 0, L, 0,
                                                          _stream(g_out))
        _lib.check(rc, "ray_features_backward")
        return None, None, g_table, None, None, None


def ray_features_supported(encoder):
    return (encoder.input_dim == 3 and encoder.level_dim in (2, 4, 8) and encoder.gridtype_id == 0
            and encoder.interp_id == 0 and not encoder.align_corners and encoder.embeddings.dtype == torch.float32)


def ray_features(x01, weights, encoder):
    """x01 [N,T,3] in [0,1]^3, weights [N,T] -> [N, L*C] ray-composited features of ``encoder`` (a GridEncoder)."""
    if not x01.is_cuda:
        raise RuntimeError("ray_features needs CUDA tensors (there is no CPU fallback)")
    return _RayFeatures.apply(x01, weights, encoder.embeddings, encoder.offsets, float(np.log2(encoder.per_level_scale)),
                              int(encoder.base_resolution))


# ------------------------------------------------------------------------- wide MLP head on the tensor cores
def gemm_tc(A, B, C, M, N, K, a_trans=False, b_trans=False, k_splits=1, epilogue=0, bias=None, act=False, slope=0.01,
            mask=None, mask_cols=0, colsum=None, precision=0):
    """C[M,N] (op)= A . B^T through ``sanerf_gemm_tc`` (csrc/gemm_tc.cu); row strides are taken from the tensors."""
    for t in (A, B, C, mask):
        if t is not None and (t.stride(-1) != 1 or t.dtype != torch.float32 or not t.is_cuda):
            raise RuntimeError("gemm_tc needs fp32 CUDA matrices with unit column stride")
    with _lib.stats.span("gemm_tc", M=M, N=N, K=K, epilogue=epilogue):
        rc = _lib.load().sanerf_gemm_tc(A.data_ptr(), A.stride(0), int(a_trans), B.data_ptr(), B.stride(0), int(b_trans),
                                        C.data_ptr(), C.stride(0), M, N, K, k_splits, epilogue, _lib.ptr(bias), int(act),
                                        float(slope), _lib.ptr(mask), 0 if mask is None else mask.stride(0), mask_cols,
                                        _lib.ptr(colsum), precision, _stream(C))
    _lib.check(rc, "gemm_tc")


def skip_mlp_forward(x, weights, biases, skip_layers, precision=0, slope=0.01):
    """SkipConnMLP.forward (network.py:57-75): every layer = one tensor-core GEMM with bias + leaky ReLU in its epilogue.
    Returns (output [M, out], the list of layer inputs saved for the backward)."""
    M, n = x.shape[0], len(weights)
    inputs, h = [], x
    for i, (W, b) in enumerate(zip(weights, biases)):
        if i in skip_layers:
            h = torch.cat([h, x], dim=-1)
        inputs.append(h)
        out = torch.empty(M, W.shape[0], device=x.device, dtype=torch.float32)
        gemm_tc(h, W, out, M, W.shape[0], W.shape[1], bias=b, act=(i + 1 < n), slope=slope, precision=precision)
        h = out
    return h, inputs


def skip_mlp_backward(g_out, inputs, weights, skip_layers, g_weights, g_biases, precision=0, slope=0.01,
                      need_input_grad=True, k_splits=16, input_cols=None, side_stream=None, join=True):
    """Backward of ``skip_mlp_forward``: ACCUMULATES weight / bias gradients into ``g_weights`` / ``g_biases`` (pre-zeroed
    or holding earlier contributions) and returns the gradient of the MLP input.  Per layer: one weight-gradient GEMM
    split over the batch rows + a column sum (bias), and one data-gradient GEMM whose epilogue applies the derivative of
    the previous layer's leaky ReLU (its saved output is the mask).  ``input_cols``: return only that
    many leading columns of the input gradient.  ``side_stream``: the weight-gradient GEMMs (nothing downstream reads
    them) run there, off the dependent chain of data-gradient GEMMs; joined before returning, or, with ``join=False``,
    by the caller (``current.wait_stream(side_stream)``), who then receives ``(grad, keep)`` and must hold ``keep`` —
    the tensors the side stream still reads — until that join is enqueued."""
    M = g_out.shape[0]
    lib = _lib.load()
    dev = g_out.device
    main = torch.cuda.current_stream(dev)
    g = g_out.contiguous()
    g_skip, gx = None, None
    keep = []                                              # tensors the side stream reads: alive until the join
    for i in range(len(weights) - 1, -1, -1):
        W, h_in = weights[i], inputs[i]
        n_out, n_in = W.shape
        keep.append(g)
        if side_stream is not None:
            side_stream.wait_stream(main)                  # g (and everything before it) is complete
        with torch.cuda.stream(side_stream if side_stream is not None else main):
            gemm_tc(g, h_in, g_weights[i], n_out, n_in, M, a_trans=True, b_trans=True, k_splits=k_splits, epilogue=2,
                    precision=precision)
            if g_biases[i] is not None:                    # bias gradient: column sums of g, beside the weight-gradient GEMM
                with _lib.stats.span("colsum_add", M=M, N=n_out):
                    rc = lib.sanerf_colsum_add(g.data_ptr(), g.stride(0), M, n_out, g_biases[i].data_ptr(), _stream(g))
                _lib.check(rc, "colsum_add")
        if i == 0:
            if need_input_grad:
                cols = n_in if input_cols is None else min(int(input_cols), n_in)    # only the leading columns are wanted
                gx = torch.empty(M, cols, device=dev, dtype=torch.float32)
                gemm_tc(g, W, gx, M, cols, n_out, b_trans=True, precision=precision)
                if g_skip is not None:
                    gx = gx + g_skip[:, :cols]
            break
        hid = weights[i - 1].shape[0]
        d_in = torch.empty(M, n_in, device=dev, dtype=torch.float32)
        gemm_tc(g, W, d_in, M, n_in, n_out, b_trans=True, epilogue=1, mask=h_in, mask_cols=hid, slope=slope,
                precision=precision)       # (colsum= in this epilogue costs 11 us of same-address reductions on the chain)
        if i in skip_layers:
            g_skip = d_in[:, hid:] if g_skip is None else g_skip + d_in[:, hid:]
            g = d_in[:, :hid].contiguous()
        else:
            g = d_in
    if side_stream is not None and not join:
        return gx, keep
    if side_stream is not None:
        main.wait_stream(side_stream)
    del keep
    return gx


def layernorm_mse(x, ln, target_map, loss, y_out=None):
    """LayerNorm + MSE against ``target_map`` [1, C, h, w] (row r of x <-> pixel r), forward and backward in one kernel:
    accumulates the loss into ``loss`` (1 float) and the affine gradients into ``ln.weight.grad`` / ``ln.bias.grad``,
    returns d loss / d x."""
    M, N = x.shape
    hw = target_map.shape[-2] * target_map.shape[-1]
    if hw != M or target_map.shape[1] != N or not target_map.is_contiguous():
        raise RuntimeError("layernorm_mse needs a contiguous [1, C, h, w] target with h*w rows")
    g_x = torch.empty_like(x)
    with _lib.stats.span("layernorm_mse", M=M):
        rc = _lib.load().sanerf_layernorm_mse(x.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps),
                                              target_map.data_ptr(), 1, hw, M, N, _lib.ptr(y_out), loss.data_ptr(),
                                              g_x.data_ptr(), ln.weight.grad.data_ptr(), ln.bias.grad.data_ptr(), _stream(x))
    _lib.check(rc, "layernorm_mse")
    return g_x


class SamvitHead:
    """The samvit head of the stage-2 step (network.py:120-123: SkipConnMLP 163 -> 256 x 4 -> 256 with the input re-joined at
    layer 2, then LayerNorm) as an autograd-free plan on STATIC buffers: no concatenation, no contiguous() copies, no
    additions and no allocations inside the captured step.

    Layout trick: one [M, 420] buffer ``skip`` holds layer 1's output in columns 0..255 and the head's INPUT in columns
    256..418 (column 419 pads the row to a 16-byte multiple), so the skip layer's 419-wide input exists without a copy:
    ``ray_features_forward`` writes the 128 feature columns and ``sanerf_sam_pack`` the other 35 straight into it.  The
    backward mirrors it with ``dskip``: the skip layer's data gradient lands there, its first 256 columns feed layer 1's
    backward as a strided operand, and layer 0's data gradient is ACCUMULATED onto columns 256..383 by the GEMM's reducing
    epilogue — which leaves d loss / d f_sam exactly where ``ray_features_backward`` reads it (row stride 420)."""

    WIDTH, IN, LD = 256, 163, 420

    def __init__(self, mlp, ln, M, device):
        net = mlp.net
        ok = (len(net) == 5 and list(mlp.skip_layers) == [2] and net[0].in_features == self.IN
              and all(l.out_features == self.WIDTH for l in net) and net[2].in_features == self.WIDTH + self.IN
              and all(l.bias is not None for l in net) and isinstance(ln, torch.nn.LayerNorm)
              and tuple(ln.normalized_shape) == (self.WIDTH,) and ln.elementwise_affine)
        if not ok:
            raise ValueError("SamvitHead covers the reference's samvit_mlp (163 -> 256 x 5, skip at layer 2, LayerNorm(256))")
        self.mlp, self.ln, self.M = mlp, ln, int(M)
        f32 = dict(device=device, dtype=torch.float32)
        W = self.WIDTH
        self.skip = torch.zeros(M, self.LD, **f32)
        self.dskip = torch.zeros(M, self.LD, **f32)
        self.f = self.skip[:, W:W + self.IN]                        # the head's input row (strided view)
        self.h0, self.h2, self.h3, self.out, self.samvit = (torch.empty(M, W, **f32) for _ in range(5))
        self.g_out, self.g3, self.g2, self.g0 = (torch.empty(M, W, **f32) for _ in range(4))
        self.precision = PRECISION_IDS[mlp.precision]
        # copies of the two weights whose rows are not 16-byte multiples ([256,163], [256,419]) with padded leading
        # dimensions: TMA-addressable, so that every GEMM of the dependent chain takes the TMA-fed kernel (gemm_tma.cu)
        self.w0p = torch.zeros(W, self.IN + 1, **f32)
        self.w2p = torch.zeros(W, self.LD, **f32)

    def refresh_weights(self):
        """Re-copy the padded weight copies from the parameters (two copy nodes; call once per step before ``forward``)."""
        lib, net = _lib.load(), self.mlp.net
        st = _stream(self.w0p)
        for dst, lin in ((self.w0p, net[0]), (self.w2p, net[2])):
            rc = lib.sanerf_copy_rows(dst.data_ptr(), dst.stride(0), lin.weight.data_ptr(), lin.weight.stride(0), lin.out_features,
                                      lin.in_features, st)
            _lib.check(rc, "copy_rows")

    def _weights(self):
        net = self.mlp.net
        w = [l.weight.detach() for l in net]
        w[0], w[2] = self.w0p[:, :self.IN], self.w2p[:, :self.WIDTH + self.IN]
        return w

    def forward(self):
        """skip[:, 256:419] (filled by the caller) -> out [M, 256] (pre-LayerNorm)."""
        net, W, M, pr = self.mlp.net, self.WIDTH, self.M, self.precision
        w = self._weights()
        b = [l.bias.detach() for l in net]
        gemm_tc(self.f, w[0], self.h0, M, W, self.IN, bias=b[0], act=True, precision=pr)
        gemm_tc(self.h0, w[1], self.skip, M, W, W, bias=b[1], act=True, precision=pr)             # -> skip[:, :256]
        gemm_tc(self.skip, w[2], self.h2, M, W, W + self.IN, bias=b[2], act=True, precision=pr)   # 419-wide input, no concat
        gemm_tc(self.h2, w[3], self.h3, M, W, W, bias=b[3], act=True, precision=pr)
        gemm_tc(self.h3, w[4], self.out, M, W, W, bias=b[4], act=False, precision=pr)
        return self.out

    def loss_backward(self, target_map, loss):
        """LayerNorm + MSE forward / backward (one kernel): loss accumulated, ln gradients accumulated, g_out filled."""
        M, N = self.out.shape
        hw = target_map.shape[-2] * target_map.shape[-1]
        if hw != M or target_map.shape[1] != N or not target_map.is_contiguous():
            raise RuntimeError("SamvitHead needs a contiguous [1, 256, h, w] target with h*w rows")
        ln = self.ln
        with _lib.stats.span("layernorm_mse", M=M):
            rc = _lib.load().sanerf_layernorm_mse(self.out.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps),
                                                  target_map.data_ptr(), 1, hw, M, N, self.samvit.data_ptr(), loss.data_ptr(),
                                                  self.g_out.data_ptr(), ln.weight.grad.data_ptr(), ln.bias.grad.data_ptr(),
                                                  _stream(self.out))
        _lib.check(rc, "layernorm_mse")

    def backward(self, side_stream, n_feat, slope=0.01, k_splits=16):
        """g_out -> weight / bias gradients ACCUMULATED into the parameters' .grad (flat views), and d loss / d f[:, :n_feat]
        returned as a view of ``dskip`` (row stride 420).  The weight-gradient GEMMs and bias column sums run on
        ``side_stream`` beside the dependent chain of data-gradient GEMMs; the CALLER joins the side stream."""
        net, W, M, pr = self.mlp.net, self.WIDTH, self.M, self.precision
        w = self._weights()
        gw = [l.weight.grad for l in net]
        gb = [l.bias.grad for l in net]
        lib = _lib.load()
        main = torch.cuda.current_stream(self.out.device)
        inputs = [self.f, self.h0, self.skip, self.h2, self.h3]
        in_cols = [self.IN, W, W + self.IN, W, W]
        g1 = self.dskip[:, :W]                                      # gradient w.r.t. layer 1's output (strided)

        def weight_grad(i, g):
            side_stream.wait_stream(main)                           # g is complete
            with torch.cuda.stream(side_stream):
                gemm_tc(g, inputs[i], gw[i], W, in_cols[i], M, a_trans=True, b_trans=True, k_splits=k_splits, epilogue=2,
                        precision=pr)
                with _lib.stats.span("colsum_add", M=M, N=W):
                    rc = lib.sanerf_colsum_add(g.data_ptr(), g.stride(0), M, W, gb[i].data_ptr(), _stream(g))
                _lib.check(rc, "colsum_add")

        weight_grad(4, self.g_out)
        gemm_tc(self.g_out, w[4], self.g3, M, W, W, b_trans=True, epilogue=1, mask=self.h3, mask_cols=W, slope=slope, precision=pr)
        weight_grad(3, self.g3)
        gemm_tc(self.g3, w[3], self.g2, M, W, W, b_trans=True, epilogue=1, mask=self.h2, mask_cols=W, slope=slope, precision=pr)
        weight_grad(2, self.g2)
        # skip layer: [M, 419] data gradient; the first 256 columns pass through layer 1's leaky-ReLU derivative
        gemm_tc(self.g2, w[2], self.dskip, M, W + self.IN, W, b_trans=True, epilogue=1, mask=self.skip, mask_cols=W, slope=slope,
                precision=pr)
        weight_grad(1, g1)
        gemm_tc(g1, w[1], self.g0, M, W, W, b_trans=True, epilogue=1, mask=self.h0, mask_cols=W, slope=slope, precision=pr)
        weight_grad(0, self.g0)
        g_feat = self.dskip[:, W:W + n_feat]                        # holds the skip path's share already
        gemm_tc(self.g0, w[0], g_feat, M, n_feat, W, b_trans=True, epilogue=2, precision=pr)      # += layer 0's share
        return g_feat


class _SkipMLP(Function):
    @staticmethod
    def forward(ctx, x, skip_layers, precision, has_bias, *params):
        n = len(params) // 2 if has_bias else len(params)
        weights = [p.detach() for p in (params[0::2] if has_bias else params)]
        biases = [p.detach() for p in params[1::2]] if has_bias else [None] * n
        out, inputs = skip_mlp_forward(x.detach().contiguous().float(), weights, biases, skip_layers, precision)
        ctx.save_for_backward(*inputs, *weights)
        ctx.meta = (n, tuple(skip_layers), precision, has_bias)
        return out

    @staticmethod
    def backward(ctx, g_out):
        n, skip_layers, precision, has_bias = ctx.meta
        saved = ctx.saved_tensors
        inputs, weights = list(saved[:n]), list(saved[n:])
        gW = [torch.zeros_like(w) for w in weights]
        gB = [torch.zeros(w.shape[0], device=w.device) if has_bias else None for w in weights]
        gx = skip_mlp_backward(g_out.float(), inputs, weights, skip_layers, gW, gB, precision,
                               need_input_grad=ctx.needs_input_grad[0])
        grads = [v for pair in zip(gW, gB) for v in pair] if has_bias else gW
        return (gx, None, None, None, *grads)


def skip_mlp_supported(mlp, x):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and not torch.is_autocast_enabled()
            and all(l.weight.dtype == torch.float32 for l in mlp.net)
            and len({l.bias is None for l in mlp.net}) == 1)


def skip_mlp(x, mlp, precision="fp32"):
    """SkipConnMLP on the tensor cores (autograd-aware); ``mlp`` = nerf.network.SkipConnMLP."""
    has_bias = mlp.net[0].bias is not None
    params = [t for l in mlp.net for t in ((l.weight, l.bias) if has_bias else (l.weight,))]
    return _SkipMLP.apply(x, tuple(mlp.skip_layers), PRECISION_IDS[precision], has_bias, *params)


class _ProposalLoss(Function):
    """sum over proposal levels of the inter-level loss; weights of the final level are constants."""

    @staticmethod
    def forward(ctx, t_ref, w_ref, *levels):
        t_ref, w_ref = t_ref.contiguous(), w_ref.detach().contiguous()
        N, Tr = w_ref.shape
        loss = torch.zeros(1, device=w_ref.device)
        grads = []
        lib = _lib.load()
        with torch.cuda.device(w_ref.device):
            for t_p, w_p in zip(levels[0::2], levels[1::2]):
                t_p, w_p = t_p.contiguous(), w_p.contiguous()
                g = torch.empty_like(w_p)
                with _lib.stats.span("proposal_loss", N=N, Tp=w_p.shape[1]):
                    rc = lib.sanerf_proposal_loss(t_ref.data_ptr(), w_ref.data_ptr(), Tr, t_p.data_ptr(), w_p.data_ptr(),
                                                  w_p.shape[1], N, 1.0, loss.data_ptr(), g.data_ptr(), _stream(w_ref))
                _lib.check(rc, "proposal_loss")
                grads.append(g)
        ctx.save_for_backward(*grads)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        out = [None, None]
        for gw in ctx.saved_tensors:
            out += [None, gw * g]
        return tuple(out)


def proposal_loss(all_bins, all_weights):
    """renderer.py:30-57 — last entries are the (detached) final level."""
    args = []
    for b, w in zip(all_bins[:-1], all_weights[:-1]):
        args += [b, w]
    return _ProposalLoss.apply(all_bins[-1], all_weights[-1], *args)


class _DistortLoss(Function):
    @staticmethod
    def forward(ctx, bins, weights):
        bins, weights = bins.contiguous(), weights.contiguous()
        N, T = weights.shape
        loss = torch.zeros(1, device=weights.device)
        g = torch.empty_like(weights)
        lib = _lib.load()
        with torch.cuda.device(weights.device), _lib.stats.span("distortion_loss", N=N, T=T):
            rc = lib.sanerf_distortion_loss(bins.data_ptr(), weights.data_ptr(), T, N, 1.0, loss.data_ptr(), g.data_ptr(),
                                            _stream(weights))
        _lib.check(rc, "distortion_loss")
        ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return None, ctx.saved_tensors[0] * g


def distort_loss(bins, weights):
    """renderer.py:17-27."""
    return _DistortLoss.apply(bins, weights)


# ----------------------------------------------------------------------------------------- optimizer
class FusedAdam:
    """Adam(eps=1e-15) + LambdaLR(0.1**min(it/iters,1)) (+ optional EMA, ``ema_decay``) over parameters flattened into ONE
    fp32 buffer.

    Parameters and gradients become views of ``flat_param`` / ``flat_grad`` (the gradient bucket doubles as the
    NCCL all-reduce buffer); one kernel updates everything and clears the gradient in the same pass.  The step
    counter and the schedule live on the device, so a captured CUDA graph replays with a moving learning rate.

    ``ema_decay`` (the reference trains with 0.95: main.py:316, nerf/utils.py:616 ``torch_ema.ExponentialMovingAverage``):
    a shadow copy of the flat parameters, updated by ``ema_update()`` — which the reference calls once per EPOCH
    (nerf/utils.py:1862, after the loop of ``train_one_epoch``; :1627 once per 16 GUI steps), with torch_ema's warm-up
    ``min(decay, (1 + k) / (10 + k))`` over the number of updates.  ``ema_every_step=True`` instead folds the update into
    every Adam pass (same formula over the step count; one extra read + write of the shadow in the same kernel).

    ``gate`` (1 int32 on the device) makes a DEFERRED range update idempotent: ``schedule()`` sets it at the start of
    a step, ``apply(..., gated=True)`` does nothing while it is 0, and ``clear_gate()`` (called by a trainer's
    ``flush()`` after it applied the pending update early) resets it — so the deferred pass baked into the next
    step's CUDA graph finds nothing to do instead of moving the parameters by momentum on a zero gradient.
    """

    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-15, decay_iters=20000, ema_decay=None, world_size=1,
                 ema_every_step=False):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        # every slot is a multiple of 32 elements: views stay 16-byte aligned and any range of whole slots splits into
        # 2 / 4 / 8 equal, 16-byte-aligned shards (reduce-scatter + sharded update + all-gather, sanerf_b200/step.py)
        sizes = [(p.numel() + 31) // 32 * 32 for p in self.params]
        total = sum(sizes)
        # multi-GPU: parameters and gradients live in symmetric memory, so that the update kernel of every rank can read
        # the peers' gradients and write the peers' parameters over NVLink (sanerf_b200/symm.py, csrc/symm_adam.cu)
        from . import symm as _symm
        self.symm = _symm.SymmetricState(total, dev) if _symm.enabled(world_size) else None
        self.sharded = {}                                   # (start, stop) -> (lo, hi): ranges whose m / v / EMA are rank-sharded
        if self.symm is not None:
            self.flat_param, self.flat_grad = self.symm.param, self.symm.grad
        else:
            self.flat_param = torch.zeros(total, device=dev)
            self.flat_grad = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        off = 0
        self.ranges = {}                                    # id(param) -> (start, stop) in the flat buffers
        for p, n in zip(self.params, sizes):
            self.ranges[id(p)] = (off, off + n)
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += n
        self.lr, self.betas, self.eps, self.decay_iters = float(lr), betas, float(eps), float(decay_iters)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.dyn = torch.tensor([self.lr, 1.0, 1.0, 0.0], device=dev)     # {lr_t, 1-b1^t, 1-b2^t, 1-ema_t}: neutral before step 1
        self.gate = torch.zeros(1, device=dev, dtype=torch.int32)
        self.ema_decay = None if ema_decay is None else float(ema_decay)
        self.ema = self.flat_param.clone() if ema_decay is not None else None
        self.ema_every_step = bool(ema_every_step) and self.ema is not None
        self.ema_updates = 0
        self._ema_backup = None

    def zero_grad(self):
        self.flat_grad.zero_()

    def range_of(self, params):
        """Smallest flat range [start, stop) covering ``params``; raises unless they fill it without foreign slots."""
        spans = sorted(self.ranges[id(p)] for p in params)
        if not spans:
            return 0, 0
        for (_, b0), (a1, _) in zip(spans, spans[1:]):
            if b0 != a1:
                raise ValueError("parameters are not contiguous in the flat buffer")
        return spans[0][0], spans[-1][1]

    def schedule(self):
        """Advance the device-side step counter and learning-rate / bias-correction / EMA terms (one tiny kernel);
        marks the deferred ranges as pending (``gate`` = 1)."""
        lib = _lib.load()
        dev = self.flat_param.device
        with torch.cuda.device(dev), _lib.stats.span("adam_schedule"):
            rc = lib.sanerf_adam_schedule(self.step_count.data_ptr(), self.dyn.data_ptr(), self.lr, self.betas[0],
                                          self.betas[1], self.decay_iters, self.gate.data_ptr(),
                                          self.ema_decay if self.ema_every_step else 0.0, _lib.current_stream(dev))
        _lib.check(rc, "adam_schedule")

    def apply(self, start=0, stop=None, grad_scale=1.0, zero_grad=True, gated=False):
        """Adam update of the flat range [start, stop) with the terms of the LAST ``schedule()`` (multiples of 4).
        ``gated``: a deferred range — skipped on the device while ``gate`` is 0."""
        stop = self.flat_param.numel() if stop is None else stop
        n = stop - start
        if n <= 0:
            return
        lib = _lib.load()
        dev = self.flat_param.device
        off = 4 * start
        with torch.cuda.device(dev), _lib.stats.span("adam_step", n=n):
            rc = lib.sanerf_adam_step(self.flat_param.data_ptr() + off, self.flat_grad.data_ptr() + off,
                                      self.exp_avg.data_ptr() + off, self.exp_avg_sq.data_ptr() + off, n, self.dyn.data_ptr(),
                                      self.betas[0], self.betas[1], self.eps, float(grad_scale), int(zero_grad),
                                      self.gate.data_ptr() if gated else None,
                                      self.ema.data_ptr() + off if self.ema_every_step else None, _lib.current_stream(dev))
        _lib.check(rc, "adam_step")

    def apply_symm(self, start, stop, gated=False, blocks=None, channel=0):
        """Multi-GPU form of ``apply``: ONE kernel sums the gradient of this rank's slice of [start, stop) over all ranks
        through NVLink / NVSwitch, updates the slice (1/world scaling, EMA) and writes the new parameters into every rank's
        buffer, clearing the gradient everywhere (csrc/symm_adam.cu).  Collective: every rank must call it.  Calls that
        may overlap in time (different streams) must use different ``channel``s (0..3)."""
        from . import symm as _symm
        n = stop - start
        import os
        threads = int(os.environ.get("SANERF_SYMM_THREADS", 1024))
        if blocks is None:                                  # wide and short: a slice of n / world at ~4 float4 per thread
            cap = min(_symm.CHANNEL_BLOCKS[channel], int(os.environ.get("SANERF_SYMM_BLOCKS", 64)))
            blocks = max(1, min(cap, (n // (4 * self.symm.world) + threads * 4 - 1) // (threads * 4)))
        self.symm.launch(self, start, stop, 1.0 / self.symm.world, gated, blocks, threads, channel)
        self.sharded[(start, stop)] = _symm.slice_bounds(start, stop, self.symm.world, self.symm.rank)

    def gather_sharded_state(self):
        """Ranges updated rank-sharded (``apply_symm`` / the NCCL reduce-scatter form) keep exp_avg / exp_avg_sq / EMA
        only for the rank's own slice: make them whole on every rank (checkpoints, EMA evaluation).  Collective."""
        import torch.distributed as dist
        bufs = [self.exp_avg, self.exp_avg_sq] + ([self.ema] if self.ema_every_step else [])
        for (a, b), (lo, hi) in self.sharded.items():
            for buf in bufs:
                tmp = torch.zeros(b - a, device=buf.device)
                tmp[lo - a:hi - a] = buf[lo:hi]
                dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
                buf[a:b] = tmp

    def clear_gate(self):
        self.gate.zero_()

    def step(self, grad_scale=1.0, zero_grad=True):
        self.schedule()
        self.apply(0, None, grad_scale, zero_grad)

    # ---- EMA (torch_ema surface the reference's Trainer uses: update once per epoch, nerf/utils.py:1862; store / copy_to /
    # restore bracket evaluation and best-checkpoint saving, :1684-1695, 1900-1902, 2035-2036, 2083-2095)
    def ema_update(self):
        """``ExponentialMovingAverage.update()``: call after the epoch's last step (and after the trainer's ``flush()``, so
        that a deferred table update has landed).  A no-op in ``ema_every_step`` mode."""
        if self.ema is None:
            raise RuntimeError("FusedAdam was built without ema_decay")
        if self.ema_every_step:
            return
        self.ema_updates += 1
        decay = min(self.ema_decay, (1 + self.ema_updates) / (10 + self.ema_updates))
        dev = self.flat_param.device
        with torch.cuda.device(dev), _lib.stats.span("ema_update", n=self.flat_param.numel()):
            rc = _lib.load().sanerf_ema_update(self.ema.data_ptr(), self.flat_param.data_ptr(), self.flat_param.numel(),
                                               1.0 - decay, _lib.current_stream(dev))
        _lib.check(rc, "ema_update")

    def ema_store(self):
        self._ema_backup = self.flat_param.clone()

    def ema_copy_to(self):
        if self.ema is None:
            raise RuntimeError("FusedAdam was built without ema_decay")
        if self.sharded and self.ema_every_step:
            self.gather_sharded_state()
        self.flat_param.copy_(self.ema)

    def ema_restore(self):
        if self._ema_backup is None:
            raise RuntimeError("ema_restore() without ema_store()")
        self.flat_param.copy_(self._ema_backup)
        self._ema_backup = None

    def ema_state_dict(self):
        """Same layout as ``torch_ema.ExponentialMovingAverage.state_dict()`` over the trainable parameters."""
        if self.ema is None:
            raise RuntimeError("FusedAdam was built without ema_decay")
        if self.sharded and self.ema_every_step:
            self.gather_sharded_state()
        shadow = [self.ema[a:a + p.numel()].view_as(p).clone() for p, (a, _) in ((p, self.ranges[id(p)]) for p in self.params)]
        n_upd = int(self.step_count.item()) if self.ema_every_step else self.ema_updates
        return {"decay": self.ema_decay, "num_updates": n_upd, "shadow_params": shadow,
                "collected_params": None}

    def load_ema_state_dict(self, state):
        if self.ema is None:
            raise RuntimeError("FusedAdam was built without ema_decay")
        shadow = state["shadow_params"]
        if len(shadow) != len(self.params):
            raise ValueError("EMA state holds a different number of parameters")
        for p, s in zip(self.params, shadow):
            a, _ = self.ranges[id(p)]
            self.ema[a:a + p.numel()].copy_(s.reshape(-1))
        self.ema_updates = int(state.get("num_updates") or 0)
