"""``NeRFRenderer`` — proposal sampling, sigma->weights, per-ray compositing (sm_100a path).

Mirrors the behaviour and the public surface of the reference's ``nerf/renderer.py``
(``NeRFRenderer.render / run``, module-level ``contract``, ``sample_pdf``, ``near_far_from_aabb``,
``proposal_loss``, ``distort_loss``) for the RGB and SAM-feature stages.  The numerics of every
statement are those of the reference (cited inline); what changes is the execution:

* sigma -> alpha -> transmittance -> weights -> weighted sums is ONE kernel per direction
  (``sanerf_b200.ops.composite``) instead of ~10 torch kernels and ``[N,T,C]`` temporaries
  (renderer.py:309-338, :377);
* the view direction is SH-encoded once per ray and broadcast, not once per sample
  (network.py:237 evaluates the same direction T times);
* with ``opt.fused`` (default on) each sampling level is ONE kernel (near/far, spacing, jitter or inverse-CDF
  resampling, mid points, intervals, positions, contraction, unit-cube mapping), the proposal density is ONE
  kernel (grid encode -> MLP -> trunc_exp), ``trunc_exp`` + compositing read the MLP head in place, and the
  proposal / distortion losses are one kernel each (``sanerf_b200/fused.py``);
* mask heads (stage 3) are not part of this path.
"""
import math

import torch
import torch.nn as nn

from sanerf_b200 import fused
from sanerf_b200.ops import composite


def contract(x):
    """Scene contraction of the inf-norm ball (renderer.py:60-69): identity inside the unit cube,
    ``(2 - 1/|x|_inf) / |x|_inf`` on the dominant axis and ``1/|x|_inf`` on the others outside."""
    lead, C = x.shape[:-1], x.shape[-1]
    flat = x.reshape(-1, C)
    mag, dominant = flat.abs().max(1, keepdim=True)
    scale = (1 / mag).repeat(1, C)
    scale.scatter_(1, dominant, (2 - 1 / mag) / mag)
    return torch.where(mag < 1, flat, flat * scale).view(*lead, C)


def uncontract(z):
    """Inverse of ``contract`` (renderer.py:72-81)."""
    lead, C = z.shape[:-1], z.shape[-1]
    flat = z.reshape(-1, C)
    mag, dominant = flat.abs().max(1, keepdim=True)
    scale = 1 / (2 - mag.repeat(1, C)).clamp(min=1e-8)
    scale.scatter_(1, dominant, 1 / (2 * mag - mag * mag).clamp(min=1e-8))
    return torch.where(mag < 1, flat, flat * scale).view(*lead, C)


def sample_pdf(bins, weights, T, perturb=False):
    """Inverse-CDF resampling of ``T`` bin edges from a piecewise-constant PDF (renderer.py:84-119)."""
    N, T0 = weights.shape
    w = weights + 0.01
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1).clamp(max=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T, device=weights.device).expand(N, T)
    if perturb:
        u = u + (torch.rand_like(u) - 0.5) / T
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp(0, T0)
    hi = hi.clamp(0, T0)
    c0, c1 = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b0, b1 = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    t = torch.nan_to_num((u - c0) / (c1 - c0)).clamp(0, 1)
    return b0 + t * (b1 - b0)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.05):
    """Slab test against ``aabb`` = (xmin, ymin, zmin, xmax, ymax, zmax) (renderer.py:122-139);
    rays that miss get near = far = 1e9."""
    inv = rays_d + 1e-15
    t0, t1 = (aabb[:3] - rays_o) / inv, (aabb[3:] - rays_o) / inv
    near = torch.minimum(t0, t1).amax(dim=-1, keepdim=True)
    far = torch.maximum(t0, t1).amin(dim=-1, keepdim=True)
    miss = far < near
    near = torch.where(miss, torch.full_like(near, 1e9), near).clamp(min=min_near)
    far = torch.where(miss, torch.full_like(far, 1e9), far)
    return near, far


def proposal_loss(all_bins, all_weights):
    """Inter-level upper-bound loss of Mip-NeRF 360 (renderer.py:30-57)."""
    t_ref, w_ref = all_bins[-1].detach(), all_weights[-1].detach()
    total = 0
    for t_p, w_p in zip(all_bins[:-1], all_weights[:-1]):
        n = w_p.shape[-1]
        cum = torch.cat([torch.zeros_like(w_p[..., :1]), torch.cumsum(w_p, dim=-1)], dim=-1)
        lo = (torch.searchsorted(t_p[..., :-1].contiguous(), t_ref[..., :-1].contiguous(), right=True) - 1).clamp(0, n - 1)
        hi = torch.searchsorted(t_p[..., 1:].contiguous(), t_ref[..., 1:].contiguous(), right=True).clamp(0, n - 1)
        bound = torch.take_along_dim(cum[..., 1:], hi, dim=-1) - torch.take_along_dim(cum[..., :-1], lo, dim=-1)
        total = total + ((w_ref - bound).clamp(min=0) ** 2 / (w_ref + 1e-8)).mean()
    return total


def distort_loss(bins, weights):
    """Distortion regulariser (renderer.py:17-27).  The reference calls the third-party
    ``torch_efficient_distloss.eff_distloss``; this is its O(T) prefix-sum form
    1/3 sum d_i w_i^2 + 2 sum_i w_i (m_i W_<i - WM_<i), averaged over rays."""
    d = bins[..., 1:] - bins[..., :-1]
    m = bins[..., :-1] + d / 2
    wm = weights * m
    W = torch.cumsum(weights, -1) - weights
    WM = torch.cumsum(wm, -1) - wm
    per_sample = d * weights.pow(2) / 3 + 2 * weights * (m * W - WM)
    return per_sample.sum() / max(1, weights[..., 0].numel())


def _spacing(x):
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))  # renderer.py:250


def _spacing_inv(s):
    return torch.where(s < 0.5, 2 * s, 1 / (2 - 2 * s))  # renderer.py:252-253


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.real_bound = opt.bound
        self.bound = 2 if opt.contract else opt.bound          # grid query bound (renderer.py:152-155)
        self.cascade = 1 + math.ceil(math.log2(self.bound))
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        box = torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32)
        self.register_buffer("aabb_train", box)
        self.register_buffer("aabb_infer", box.clone())
        # early ray termination is opt-in: the reference declares --T_thresh but never reads it
        self.t_thresh = float(getattr(opt, "t_thresh_composite", 0.0))
        self.fused = bool(getattr(opt, "fused", True))

    def forward(self, x, d, **kwargs):
        raise NotImplementedError()

    def density(self, x, **kwargs):
        raise NotImplementedError()

    def update_aabb(self, aabb):
        if not torch.is_tensor(aabb):
            aabb = torch.as_tensor(aabb).float()
        # in place: captured CUDA graphs (sanerf_b200/step.py) hold the addresses of these two buffers
        self.aabb_train.copy_(aabb.clamp(-self.real_bound, self.real_bound).to(self.aabb_train.device))
        self.aabb_infer.copy_(self.aabb_train)

    def render(self, rays_o, rays_d, staged=False, cam_near_far=None, **kwargs):
        """Same contract as renderer.py:185-219: ``staged`` splits the rays into ``opt.max_ray_batch`` chunks."""
        if not staged:
            return self.run(rays_o, rays_d, cam_near_far=cam_near_far, **kwargs)
        N, step = rays_o.shape[0], self.opt.max_ray_batch
        merged = {}
        for head in range(0, N, step):
            tail = min(head + step, N)
            cnf = cam_near_far
            if cnf is not None and cnf.shape[0] != 1:
                cnf = cnf[head:tail]
            part = self.run(rays_o[head:tail], rays_d[head:tail], cam_near_far=cnf, **kwargs)
            for k, v in part.items():
                if v is None:
                    continue
                if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == tail - head:
                    if k not in merged:
                        merged[k] = torch.empty(N, *v.shape[1:], device=v.device, dtype=v.dtype)
                    merged[k][head:tail] = v
                else:
                    merged[k] = v
        return merged

    def _sample_positions(self, rays_o, rays_d, s_near, s_far, bins):
        """bins in [0,1] -> metric bin edges, mid points, interval lengths and (contracted) positions
        (renderer.py:278-286)."""
        edges = _spacing_inv(s_near * (1 - bins) + s_far * bins)
        t_mid = (edges[..., 1:] + edges[..., :-1]) / 2
        deltas = edges[..., 1:] - edges[..., :-1]
        xyz = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * t_mid.unsqueeze(2)
        if self.opt.contract:
            xyz = contract(xyz)
        return t_mid, deltas, xyz

    def run(self, rays_o, rays_d, bg_color=None, perturb=False, cam_near_far=None, update_proposal=True,
            return_feats=0, return_mask=0, H=None, W=None, **kwargs):
        """rays [N,3] -> dict(image [N,3], depth [N], weights_sum [N], ...) (renderer.py:221-390)."""
        if return_mask:
            raise NotImplementedError("mask heads (stage 3) are outside the B200 render path")
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        N, device = rays_o.shape[0], rays_o.device
        opaque_last = self.opt.background == "last_sample"
        steps = list(self.opt.num_steps)

        if bg_color is None:
            bg_color = 1
        use_fused = (self.fused and not self.opt.sum_after_mlp and rays_o.is_cuda and rays_o.dtype == torch.float32
                     and not torch.is_autocast_enabled() and hasattr(self, "head_unit"))
        if use_fused:
            return self._run_fused(rays_o, rays_d, bg_color, perturb, cam_near_far, update_proposal, return_feats,
                                   H, W, opaque_last, steps)

        near, far = near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer,
                                       self.min_near)
        if cam_near_far is not None:
            near = torch.maximum(near, cam_near_far[:, [0]])
            far = torch.minimum(far, cam_near_far[:, [1]])
        s_near, s_far = _spacing(near), _spacing(far)

        all_bins, all_weights = [], []
        bins = weights = None
        for level, T in enumerate(steps):
            if level == 0:  # uniform (optionally jittered) edges, renderer.py:263-271
                bins = torch.linspace(0, 1, T + 1, device=device).unsqueeze(0).expand(N, -1)
                if perturb:
                    bins = (bins + (torch.rand_like(bins) - 0.5) / T).clamp(0, 1)
            else:           # resample from the previous level's weights, renderer.py:272-275
                bins = sample_pdf(bins, weights, T + 1, perturb).detach()
            t_mid, deltas, xyz = self._sample_positions(rays_o, rays_d, s_near, s_far, bins)

            if level != len(steps) - 1:
                with torch.set_grad_enabled(update_proposal and torch.is_grad_enabled()):
                    sigmas = self.density(xyz, proposal=level)["sigma"]
                weights = composite(sigmas, deltas, t_mid, None, last_sample_opaque=opaque_last)[0]
            else:
                field = self.field(xyz, rays_d)               # sigma [N,T], color [N,T,31], geo_feat
                sigmas, colors = field["sigma"], field["color"]
                weights, weights_sum, depth, f_image, n_alive = composite(
                    sigmas, deltas, t_mid, colors, last_sample_opaque=opaque_last, t_thresh=self.t_thresh)
                if self.opt.with_sam:
                    features = self.s_grid(xyz, bound=self.bound)       # [N,T,128]
            if self.training:
                all_bins.append(bins)
                all_weights.append(weights)

        results = {}
        if self.opt.sum_after_mlp:  # renderer.py:339-342: shade every sample, then composite
            shaded = self.view_mlp(colors)
            image = torch.sigmoid(composite(sigmas, deltas, t_mid, shaded, last_sample_opaque=opaque_last)[3])
        else:                       # deferred shading, renderer.py:345
            image = torch.sigmoid(self.view_mlp(f_image))

        if self.training and not self.opt.with_mask and not self.opt.with_sam:
            results["num_points"] = N * steps[-1]
            results["weights"] = weights
            if self.opt.lambda_proposal > 0 and update_proposal:
                results["proposal_loss"] = proposal_loss(all_bins, all_weights)
            if self.opt.lambda_distort > 0:
                results["distort_loss"] = distort_loss(bins, weights)

        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color     # renderer.py:358
        results["weights_sum"] = weights_sum
        results["depth"] = depth
        results["image"] = image
        results["n_alive"] = n_alive

        if self.opt.with_sam:
            if self.opt.sum_after_mlp:
                raise NotImplementedError("--with_sam --sum_after_mlp is broken in the reference (renderer.py:371-372)")
            # second channel block composited with the same weights (renderer.py:377)
            f_sam = composite(sigmas, deltas, t_mid, features, last_sample_opaque=opaque_last,
                              t_thresh=self.t_thresh)[3]
            if self.opt.sam_use_view_direction:
                f = torch.cat([f_sam, f_image, image, depth.unsqueeze(-1)], dim=-1)        # renderer.py:380
            else:
                geo_sum = composite(sigmas, deltas, t_mid, field["geo_feat"].contiguous(),
                                    last_sample_opaque=opaque_last)[3]
                f = torch.cat([f_sam, geo_sum, image, depth.unsqueeze(-1)], dim=-1)        # renderer.py:383
            samvit = self.samvit_mlp(f)
            if return_feats > 0:
                results["samvit"] = samvit.view(H, W, -1)
        return results

    def _run_fused(self, rays_o, rays_d, bg_color, perturb, cam_near_far, update_proposal, return_feats, H, W,
                   opaque_last, steps):
        """Same computation as ``run`` with one kernel per sampling level, fused proposal density, in-place head
        compositing and fused losses.  Random draws are taken in the reference's order ([N,T+1] uniforms per level)."""
        N, device = rays_o.shape[0], rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        if bg_color is None:
            bg_color = 1
        common = dict(cam_near_far=cam_near_far, contract=self.opt.contract, bound=float(self.bound))
        # the reference evaluates the SAM branch (s_grid, 128-channel compositing, samvit_mlp) for every ray of every call
        # and drops the result unless return_feats > 0 (renderer.py:302-303, 367-389); skipping it then changes nothing
        want_sam = self.opt.with_sam and (return_feats > 0 or self.training)
        all_bins, all_weights = [], []
        bins = weights = None
        last = len(steps) - 1
        for level, T in enumerate(steps):
            noise = torch.rand(N, T + 1, device=device) if perturb else None
            if level == 0:
                bins, t_mid, deltas, x01 = fused.sample_uniform(rays_o, rays_d, aabb, self.min_near, T, noise, **common)
            else:
                bins, t_mid, deltas, x01 = fused.sample_pdf(rays_o, rays_d, aabb, self.min_near, bins, weights, T,
                                                            noise, **common)
            if level != last:
                with torch.set_grad_enabled(update_proposal and torch.is_grad_enabled()):
                    sigmas = self.density_unit(x01, level)
                weights = composite(sigmas, deltas, t_mid, None, last_sample_opaque=opaque_last)[0]
            else:
                head = self.head_unit(x01)                                        # [N,T,16]: sigma logit + 15 features
                sigmas, weights, weights_sum, depth, geo_sum, n_alive = fused.head_composite(
                    head, deltas, t_mid, opaque_last, self.t_thresh)
                sh = self.view_encoder(rays_d)                                    # once per ray
                f_image = torch.cat([geo_sum, weights_sum.unsqueeze(-1) * sh], dim=-1)   # = sum_i w_i [geo_i, sh]
                if want_sam:
                    features = self.features_unit(x01)                            # [N,T,128]
            if self.training:
                all_bins.append(bins)
                all_weights.append(weights)

        results = {}
        image = torch.sigmoid(self.view_mlp(f_image))
        if self.training and not self.opt.with_mask and not self.opt.with_sam:
            results["num_points"] = N * steps[-1]
            results["weights"] = weights
            if self.opt.lambda_proposal > 0 and update_proposal:
                results["proposal_loss"] = fused.proposal_loss(all_bins, all_weights)
            if self.opt.lambda_distort > 0:
                results["distort_loss"] = fused.distort_loss(bins, weights)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results.update(weights_sum=weights_sum, depth=depth, image=image, n_alive=n_alive)
        if want_sam:
            f_sam = composite(sigmas, deltas, t_mid, features, last_sample_opaque=opaque_last,
                              t_thresh=self.t_thresh)[3]
            if self.opt.sam_use_view_direction:
                f = torch.cat([f_sam, f_image, image, depth.unsqueeze(-1)], dim=-1)
            else:
                f = torch.cat([f_sam, geo_sum, image, depth.unsqueeze(-1)], dim=-1)
            samvit = self.samvit_mlp(f)
            if return_feats > 0:
                results["samvit"] = samvit.view(H, W, -1)
        return results
