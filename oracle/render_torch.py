"""ORACLE (test infrastructure, never imported by the product path): pure-torch restatement of
the reference's render hot path, runnable on CPU.

It serves three purposes: (i) the checker of the CUDA path in ``tests/`` and ``smoke()``;
(ii) the "pure-torch reference path on the host cores" that ``bench.py`` times as
``cpu_baseline`` / ``--impl reference`` (BASELINE.json north_star); (iii) fp64 gradchecks.

Each function cites the reference lines it follows (paths relative to the reference checkout).
The grid / SH encoders, which are CUDA kernels in the reference, are restated with differentiable
torch ops (gather + index_add through autograd); everything from ``nerf/renderer.py`` is torch in
the reference already and is restated statement by statement.

Third-party arithmetic not vendored by the reference: ``torch_efficient_distloss.eff_distloss``
(requirements.txt:22, un-pinned; call site renderer.py:14,25).  ``eff_distloss`` below restates
its published O(N) algorithm (Sun et al., "Improved Direct Voxel Grid Optimization", 2022, eq. 4-6):
  L = 1/3 sum_i d_i w_i^2 + 2 sum_i w_i (m_i W_{<i} - WM_{<i}),  averaged over rays.
Parity for that term is unpinned by the reference (no tests, package absent here).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoders_np, grid_np

PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]


# ----------------------------------------------------------------------------- encoders
def _level_corners(x, level, offs, S, H, D, gridtype, align_corners, interp):
    """Rows (absolute) [2^D, B] and weights [2^D, B] of every corner of one level (gridencoder.cu:132-186)."""
    dev = x.device
    res, rows, mult, hashed, _ = grid_np.level_geometry(offs, level, S, H, D, gridtype)
    if align_corners:  # :144-146
        pos = x * float(res - 1)
        pg = torch.clamp(torch.floor(pos), max=res - 2)
    else:  # :148-149
        pos = torch.clamp(x * float(res) - 0.5, min=0.0, max=float(res - 1))
        pg = torch.floor(pos)
    frac = pos - pg
    if interp == 1:
        frac = frac * frac * (3.0 - 2.0 * frac)
    pg = pg.to(torch.int64)
    # all 2^D corners at once: bit d of corner k selects the upper cell border in dimension d (:171-186)
    bits = torch.tensor([[(k >> d) & 1 for d in range(D)] for k in range(1 << D)], device=dev)   # [2^D, D]
    up = bits.bool().unsqueeze(1)                                                                 # [2^D, 1, D]
    coord = torch.where(up, torch.clamp(pg + 1, max=res - 1).unsqueeze(0), pg.unsqueeze(0))      # [2^D, B, D]
    wd = torch.where(up, frac.unsqueeze(0), (1 - frac).unsqueeze(0))                              # [2^D, B, D]
    w = wd[..., 0]
    for d in range(1, D):
        w = w * wd[..., d]
    terms = (coord * torch.tensor([int(m) for m in mult], device=dev)) & 0xFFFFFFFF
    idx = terms[..., 0]
    for d in range(1, D):
        idx = (idx ^ terms[..., d]) if hashed else ((idx + terms[..., d]) & 0xFFFFFFFF)
    return idx % rows + offs[level], w  # :78, :101


class _GridEncodeTorch(torch.autograd.Function):
    """kernel_grid / kernel_grid_backward (gridencoder.cu:82-202, 252-349) with torch gathers and
    ``index_add_`` scatters; the integer index math runs once, outside autograd."""

    @staticmethod
    def forward(ctx, x01, table, offs, S, H, gridtype, align_corners, interp, max_level):
        B, D = x01.shape
        C = table.shape[1]
        L = len(offs) - 1
        x = x01.to(torch.float32)
        keep = ~((x < 0) | (x > 1)).any(dim=1)  # :106-112
        out = torch.zeros(B, L, C, dtype=table.dtype, device=x.device)
        saved = []
        for level in range(max_level):
            idx, w = _level_corners(x, level, offs, S, H, D, gridtype, align_corners, interp)
            w = w * keep.to(w.dtype)  # :114-118 (forward zero) and :279-284 (gradient dropped)
            out[:, level] = (w.unsqueeze(-1).to(table.dtype) * table[idx]).sum(0)  # :188-192
            saved.append((idx, w))
        ctx.saved = saved
        ctx.table_shape, ctx.table_dtype = table.shape, table.dtype
        return out.reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        B = grad.shape[0]
        C = ctx.table_shape[1]
        g = grad.reshape(B, -1, C)
        gt = torch.zeros(ctx.table_shape, dtype=ctx.table_dtype, device=grad.device)
        for level, (idx, w) in enumerate(ctx.saved):  # :313-347
            upd = w.unsqueeze(-1).to(g.dtype) * g[:, level].unsqueeze(0)
            gt.index_add_(0, idx.reshape(-1), upd.reshape(-1, C))
        return None, gt, None, None, None, None, None, None, None


def grid_encode(x01, table, offsets, S, H, gridtype=0, align_corners=False, interp=0, max_level=None):
    """Differentiable (w.r.t. ``table``) torch restatement of the grid encoder and the permute of grid.py:63.
    x01 [B,D] in [0,1]; returns [B, L*C] in table dtype."""
    offs = [int(v) for v in offsets]
    L = len(offs) - 1
    max_level = L if max_level is None else min(max_level, L)
    return _GridEncodeTorch.apply(x01, table, offs, S, H, gridtype, align_corners, interp, max_level)


class GridEncoderRef(nn.Module):
    """gridencoder/grid.py:102-168 on top of ``grid_encode`` above (same parameter names/shapes)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = grid_np.per_level_scale(desired_resolution, base_resolution, num_levels)
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype_id = {"hash": 0, "tiled": 1}[gridtype]
        self.interp_id = {"linear": 0, "smoothstep": 1}[interpolation]
        self.align_corners = align_corners
        offsets = grid_np.level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim).uniform_(-1e-4, 1e-4))

    def forward(self, inputs, bound=1, max_level=None):
        x = (inputs + bound) / (2 * bound)  # grid.py:156
        lead = list(x.shape[:-1])
        out = grid_encode(x.reshape(-1, self.input_dim), self.embeddings, self.offsets.tolist(),
                          float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id,
                          self.align_corners, self.interp_id, max_level)
        return out.view(lead + [self.output_dim])


def sh_encode(dirs, degree):
    """SHEncoder.forward (sphere_harmonics.py:75-90) + kernel_sh values (shencoder.cu:43-121), closed form
    N_l^m Q_l^m(z) {Re,Im}(x+iy)^m evaluated in torch fp32 (no gradient w.r.t. dirs, like the repo's usage)."""
    d = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    lead = list(d.shape[:-1])
    d = d.reshape(-1, 3).to(torch.float32)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    A, Bm = [torch.ones_like(x)], [torch.zeros_like(x)]
    for m in range(1, degree):
        A.append(x * A[m - 1] - y * Bm[m - 1])
        Bm.append(x * Bm[m - 1] + y * A[m - 1])
    cols = [None] * (degree * degree)
    for l in range(degree):
        centre = l * l + l
        for m in range(l + 1):
            coef = encoders_np._q_poly(l, m)
            q = torch.zeros_like(z)
            for c in coef[::-1]:  # Horner
                q = q * z + float(c)
            N = encoders_np.sh_norm(l, m)
            if m == 0:
                cols[centre] = N * q
            else:
                cols[centre + m] = N * q * A[m]
                cols[centre - m] = N * q * Bm[m]
    return torch.stack(cols, dim=1).reshape(lead + [degree * degree])


def freq_encode(x, degree):
    """FreqEncoder_torch (encoding.py:30-44) == kernel_freq column order (freqencoder.cu:48-57)."""
    parts = [x]
    for f in range(degree):
        parts += [torch.sin(x * 2.0 ** f), torch.cos(x * 2.0 ** f)]
    return torch.cat(parts, dim=-1)


class _TruncExp(torch.autograd.Function):
    """activation.py:5-18."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


# ----------------------------------------------------------------------------- renderer pieces
def contract(x):
    """renderer.py:60-69."""
    shape, C = x.shape[:-1], x.shape[-1]
    x = x.reshape(-1, C)
    mag, idx = x.abs().max(1, keepdim=True)
    scale = 1 / mag.repeat(1, C)
    scale.scatter_(1, idx, (2 - 1 / mag) / mag)
    z = torch.where(mag < 1, x, x * scale)
    return z.view(*shape, C)


def sample_pdf(bins, weights, T, perturb=False, u_noise=None):
    """renderer.py:84-119.  ``u_noise`` (uniform [0,1) of shape [N,T]) replaces torch.rand_like so that
    the CUDA path and the oracle can share the random draw."""
    N, T0 = weights.shape
    weights = weights + 0.01
    weights_sum = torch.sum(weights, -1, keepdim=True)
    pdf = weights / weights_sum
    cdf = torch.cumsum(pdf, -1).clamp(max=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T).to(weights.device)
    u = u.expand(N, T)
    if perturb:
        noise = torch.rand_like(u) if u_noise is None else u_noise
        u = u + (noise - 0.5) / T
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, 0, T0)
    above = torch.clamp(inds, 0, T0)
    cdf_g0 = torch.gather(cdf, -1, below)
    cdf_g1 = torch.gather(cdf, -1, above)
    bins_g0 = torch.gather(bins, -1, below)
    bins_g1 = torch.gather(bins, -1, above)
    bins_t = torch.clamp(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0)), 0, 1)
    return bins_g0 + bins_t * (bins_g1 - bins_g0)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.05):
    """renderer.py:122-139."""
    tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
    tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
    far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
    mask = far < near
    near = torch.where(mask, torch.full_like(near, 1e9), near)
    far = torch.where(mask, torch.full_like(far, 1e9), far)
    near = torch.clamp(near, min=min_near)
    return near, far


def spacing_fn(x):
    """renderer.py:250."""
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def spacing_fn_inv(x):
    """renderer.py:252-253."""
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


def sigma_to_weights(sigmas, deltas, last_sample_opaque=True):
    """renderer.py:309-326: delta*sigma -> alpha, exclusive-cumsum transmittance, weights, nan_to_num."""
    deltas_sigmas = deltas * sigmas
    if last_sample_opaque:
        deltas_sigmas = torch.cat([deltas_sigmas[..., :-1],
                                   torch.full_like(deltas_sigmas[..., -1:], torch.inf)], dim=-1)
    alphas = 1 - torch.exp(-deltas_sigmas)
    transmittance = torch.cumsum(deltas_sigmas[..., :-1], dim=-1)
    transmittance = torch.cat([torch.zeros_like(transmittance[..., :1]), transmittance], dim=-1)
    transmittance = torch.exp(-transmittance)
    weights = alphas * transmittance
    weights = torch.nan_to_num(weights, 0)
    return weights, transmittance


def composite(sigmas, deltas, ts, feats=None, last_sample_opaque=True, t_thresh=0.0):
    """renderer.py:309-338 (+ :377) with the derived early-termination rule of SURVEY §8 c5:
    samples whose incoming transmittance T_k < t_thresh get weight 0.  Returns
    (weights, weights_sum, depth, out, n_alive)."""
    weights, trans = sigma_to_weights(sigmas, deltas, last_sample_opaque)
    alive = ~(trans < t_thresh)
    weights = torch.where(alive, weights, torch.zeros_like(weights))
    weights_sum = weights.sum(-1)
    depth = (weights * ts).sum(-1)
    out = None if feats is None else (weights.unsqueeze(-1) * feats).sum(-2)
    return weights, weights_sum, depth, out, alive.sum(-1).to(torch.int32)


def n_alive_sequential(sigmas, deltas, t_thresh, last_sample_opaque=True):
    """termination counts from a SEQUENTIAL fp32 accumulation in sample order (numpy), with
    the ulp-scale ties reported separately: returns (counts [N], tie_mask [N])."""
    x = (deltas.detach().cpu().numpy().astype(np.float32) * sigmas.detach().cpu().numpy().astype(np.float32))
    N, T = x.shape
    S = np.zeros(N, dtype=np.float32)
    counts = np.zeros(N, dtype=np.int32)
    ties = np.zeros(N, dtype=bool)
    thr = np.float32(t_thresh)
    for k in range(T):
        Tk = np.exp(-S).astype(np.float32)
        counts += (~(Tk < thr)).astype(np.int32)
        ties |= np.abs(Tk - thr) <= 4 * np.spacing(np.maximum(Tk, thr))
        S = (S + x[:, k]).astype(np.float32)
    return counts, ties


def eff_distloss(w, m, interval):
    """torch_efficient_distloss.eff_distloss restated (see module docstring); w, m, interval: [N, T]."""
    n_rays = int(np.prod(w.shape[:-1]))
    wm = w * m
    w_cumsum = w.cumsum(dim=-1)
    wm_cumsum = wm.cumsum(dim=-1)
    w_prefix = torch.cat([torch.zeros_like(w_cumsum[..., :1]), w_cumsum[..., :-1]], dim=-1)
    wm_prefix = torch.cat([torch.zeros_like(wm_cumsum[..., :1]), wm_cumsum[..., :-1]], dim=-1)
    loss_uni = (1 / 3) * interval * w.pow(2)
    loss_bi = 2 * w * (m * w_prefix - wm_prefix)
    return (loss_bi.sum() + loss_uni.sum()) / n_rays


def distort_loss(bins, weights):
    """renderer.py:17-27."""
    intervals = bins[..., 1:] - bins[..., :-1]
    mid_points = bins[..., :-1] + intervals / 2
    return eff_distloss(weights, mid_points, intervals)


def proposal_loss(all_bins, all_weights):
    """renderer.py:30-57."""

    def loss_interlevel(t0, w0, t1, w1):
        cw1 = torch.cat([torch.zeros_like(w1[..., :1]), torch.cumsum(w1, dim=-1)], dim=-1)
        inds_lo = (torch.searchsorted(t1[..., :-1].contiguous(), t0[..., :-1].contiguous(), right=True) - 1
                   ).clamp(0, w1.shape[-1] - 1)
        inds_hi = torch.searchsorted(t1[..., 1:].contiguous(), t0[..., 1:].contiguous(), right=True
                                     ).clamp(0, w1.shape[-1] - 1)
        cw1_lo = torch.take_along_dim(cw1[..., :-1], inds_lo, dim=-1)
        cw1_hi = torch.take_along_dim(cw1[..., 1:], inds_hi, dim=-1)
        w = cw1_hi - cw1_lo
        return (w0 - w).clamp(min=0) ** 2 / (w0 + 1e-8)

    bins_ref = all_bins[-1].detach()
    weights_ref = all_weights[-1].detach()
    loss = 0
    for bins, weights in zip(all_bins[:-1], all_weights[:-1]):
        loss = loss + loss_interlevel(bins_ref, weights_ref, bins, weights).mean()
    return loss


# ----------------------------------------------------------------------------- field + renderer
class MLP(nn.Module):
    """network.py:9-34 (bias-free ReLU MLP)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.num_layers = num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=bias)
            for l in range(num_layers)])

    def forward(self, x):
        for l in range(self.num_layers):
            x = self.net[l](x)
            if l != self.num_layers - 1:
                x = F.relu(x)
        return x


class SkipConnMLP(nn.Module):
    """network.py:36-75 (leaky-ReLU, skip concat of the input at ``skip_layers``)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, skip_layers=(), bias=True):
        super().__init__()
        self.num_layers, self.skip_layers = num_layers, list(skip_layers)
        net = []
        for l in range(num_layers):
            fin = dim_in if l == 0 else (dim_hidden + dim_in if l in self.skip_layers else dim_hidden)
            fout = dim_out if l == num_layers - 1 else dim_hidden
            net.append(nn.Linear(fin, fout, bias=bias))
        self.net = nn.ModuleList(net)

    def forward(self, x):
        x_in = x
        for l in range(self.num_layers):
            if l in self.skip_layers:
                x = torch.cat([x, x_in], dim=-1)
            x = self.net[l](x)
            if l != self.num_layers - 1:
                x = F.leaky_relu(x)
        return x


class RendererRef(nn.Module):
    """nerf/renderer.py:142-464 (RGB + SAM-feature stages).  Subclasses provide ``density``, ``field``,
    ``view_mlp`` (and ``s_grid`` / ``samvit_mlp`` when ``with_sam``)."""

    def __init__(self, bound=128, contract_space=True, min_near=0.2, num_steps=(128, 64, 32),
                 background="last_sample", with_sam=False, lambda_proposal=1.0, lambda_distort=0.02):
        super().__init__()
        self.real_bound = bound
        self.bound = 2 if contract_space else bound  # renderer.py:152-155
        self.contract_space = contract_space
        self.min_near = min_near
        self.num_steps = list(num_steps)
        self.background = background
        self.with_sam = with_sam
        self.lambda_proposal, self.lambda_distort = lambda_proposal, lambda_distort
        self.register_buffer("aabb_train", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))

    def run(self, rays_o, rays_d, bg_color=1, perturb=False, update_proposal=True, return_feats=0, H=None, W=None,
            noise=None, training=True):
        """renderer.py:221-390.  ``noise`` is an optional list of three uniform [N, T_i+1] tensors
        replacing the torch.rand_like draws (prop_iter 0: :269; sample_pdf: :101)."""
        N = rays_o.shape[0]
        device = rays_o.device
        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb_train, self.min_near)
        results = {}
        all_bins, all_weights = [], []
        s_nears, s_fars = spacing_fn(nears), spacing_fn(fars)
        bins = weights = None
        n_levels = len(self.num_steps)
        for it in range(n_levels):
            T = self.num_steps[it]
            if it == 0:
                bins = torch.linspace(0, 1, T + 1, device=device).unsqueeze(0).expand(N, -1)
                if perturb:
                    u = torch.rand_like(bins) if noise is None else noise[it]
                    bins = (bins + (u - 0.5) / T).clamp(0, 1)
            else:
                bins = sample_pdf(bins, weights, T + 1, perturb, None if noise is None else noise[it]).detach()
            real_bins = spacing_fn_inv(s_nears * (1 - bins) + s_fars * bins)
            rays_t = (real_bins[..., 1:] + real_bins[..., :-1]) / 2
            xyzs = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * rays_t.unsqueeze(2)
            if self.contract_space:
                xyzs = contract(xyzs)
            if it != n_levels - 1:
                with torch.set_grad_enabled(update_proposal and torch.is_grad_enabled()):
                    sigmas = self.density(xyzs, it)
            else:
                dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
                dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
                sigmas, geo_feat, colors = self.field(xyzs, dirs)
                if self.with_sam:
                    features = self.s_grid(xyzs, bound=self.bound)
            deltas = real_bins[..., 1:] - real_bins[..., :-1]
            weights, _ = sigma_to_weights(sigmas, deltas, self.background == "last_sample")
            if training:
                all_bins.append(bins)
                all_weights.append(weights)
        results["_internals"] = dict(sigmas=sigmas, deltas=deltas, rays_t=rays_t, bins=bins, colors=colors,
                                     all_bins=all_bins, all_weights=all_weights)
        weights_sum = weights.sum(-1)
        depth = (weights * rays_t).sum(-1)
        f_image = (weights.unsqueeze(-1) * colors).sum(-2)
        image = torch.sigmoid(self.view_mlp(f_image))
        if training and not self.with_sam:
            results["num_points"] = xyzs.shape[0] * xyzs.shape[1]
            results["weights"] = weights
            if self.lambda_proposal > 0 and update_proposal:
                results["proposal_loss"] = proposal_loss(all_bins, all_weights)
            if self.lambda_distort > 0:
                results["distort_loss"] = distort_loss(bins, weights)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results.update(weights_sum=weights_sum, depth=depth, image=image)
        if self.with_sam:
            f_sam = (weights.unsqueeze(-1) * features).sum(-2)
            f = torch.cat([f_sam, f_image, image, depth.unsqueeze(-1)], dim=-1)  # renderer.py:380
            samvit = self.samvit_mlp(f)
            if return_feats > 0:
                results["samvit"] = samvit.view(H, W, -1)
        return results

    def rgb_loss(self, rays_o, rays_d, gt_rgb, update_proposal=True, noise=None, perturb=True):
        """Trainer.train_step RGB branch (nerf/utils.py:897-930): MSE + lambda_prop*L_prop + lambda_dist*L_dist."""
        out = self.run(rays_o, rays_d, bg_color=1, perturb=perturb, update_proposal=update_proposal, noise=noise)
        loss = F.mse_loss(out["image"], gt_rgb, reduction="none").mean()
        if "proposal_loss" in out and self.lambda_proposal > 0:
            loss = loss + self.lambda_proposal * out["proposal_loss"]
        if "distort_loss" in out and self.lambda_distort > 0:
            loss = loss + self.lambda_distort * out["distort_loss"]
        return loss, out


class NeRFNetworkRef(RendererRef):
    """nerf/network.py:94-259 for the RGB and SAM-feature stages (mask heads are out of scope).
    Parameter names match the reference's state_dict keys."""

    def __init__(self, bound=128, contract_space=True, min_near=0.2, num_steps=(128, 64, 32),
                 background="last_sample", with_sam=False, lambda_proposal=1.0, lambda_distort=0.02,
                 grid_cls=GridEncoderRef):
        super().__init__(bound, contract_space, min_near, num_steps, background, with_sam, lambda_proposal,
                         lambda_distort)
        self.grid = grid_cls(input_dim=3, level_dim=2, num_levels=16, log2_hashmap_size=19,
                             desired_resolution=2048 * self.bound)  # network.py:102
        self.grid_mlp = MLP(32, 16, 64, 3, bias=False)  # :103
        self.view_mlp = MLP(15 + 16, 3, 32, 3, bias=False)  # :107
        if with_sam:
            self.s_grid = grid_cls(input_dim=3, num_levels=16, level_dim=8, base_resolution=16,
                                   log2_hashmap_size=19, desired_resolution=512)  # :111
            self.samvit_mlp = nn.Sequential(SkipConnMLP(128 + 15 + 16 + 4, 256, 256, 5, skip_layers=[2], bias=True),
                                            nn.LayerNorm(256))  # :120-123
        self.prop_encoders = nn.ModuleList([
            grid_cls(input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17, desired_resolution=128),
            grid_cls(input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17, desired_resolution=256)])
        self.prop_mlp = nn.ModuleList([MLP(10, 1, 16, 2, bias=False), MLP(10, 1, 16, 2, bias=False)])  # :211-219

    def density(self, x, proposal):
        """network.py:248-259."""
        h = self.prop_encoders[proposal](x, bound=self.bound)
        return trunc_exp(self.prop_mlp[proposal](h).squeeze(-1))

    def field(self, x, d):
        """network.py:221-246."""
        grid_output = self.grid(x, bound=self.bound)
        f = self.grid_mlp(grid_output)
        sigma = trunc_exp(f[..., 0])
        geo = f[..., 1:]
        color = torch.cat([geo, sh_encode(d, 4)], dim=-1)
        return sigma, geo, color
