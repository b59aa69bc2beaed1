"""Training / inference steps of the render hot path, single- and multi-GPU.

Stage-1 RGB step = what the reference's ``Trainer.train_step`` + ``train_one_epoch`` do per iteration
(nerf/utils.py:897-930, 1811-1836): render with jittered proposal sampling, MSE + proposal + distortion
losses, backward, Adam(eps=1e-15).  Stage-2 SAM step = the ``with_sam`` branch (nerf/utils.py:1095-1106) on
frozen stage-1 parameters with a synthetic [64,64,256] regression target.

Multi-GPU: one process per GPU, rays sharded by rank, full parameter replica, ONE NCCL
all-reduce (sum) over a flat fp32 gradient bucket per step, identical Adam updates on every rank.
"""
from __future__ import annotations

import types

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .fused import FusedAdam


def default_opt(**overrides):
    """The flags the shipped system actually runs with (main.py:222-226 hard overrides + defaults)."""
    opt = types.SimpleNamespace(
        bound=128, contract=True, min_near=0.2, density_thresh=10, num_steps=[128, 64, 32],
        background="last_sample", with_sam=False, with_mask=False, sum_after_mlp=False,
        sam_use_view_direction=True, mask_mlp_type="default", lambda_proposal=1.0, lambda_distort=0.02,
        max_ray_batch=16384, num_rays=4096, num_points=2 ** 18, lr=1e-2, iters=20000)
    opt.__dict__.update(overrides)
    return opt


class RGBTrainer:
    """Stage-1 step: ``step(rays_o, rays_d, gt_rgb) -> loss`` (device tensors), with optional DDP-style sharding."""

    def __init__(self, model, lr=1e-2, iters=20000, world_size=1, fused_step=True, use_graph=True, ema_decay=None):
        self.model = model.train()
        self.world_size = world_size
        self.global_step = 0
        self.fused_step, self.use_graph = bool(fused_step), bool(use_graph)
        self._plans = {}                                  # ray count -> FusedRGBStep (static buffers + CUDA graphs)
        self._checked_counts = set()
        params = [p for p in model.parameters() if p.requires_grad]
        # Adam(eps=1e-15) + LambdaLR 0.1**min(iter/iters,1)  (main.py:296,312-313) on flat buffers; the flat
        # gradient doubles as the all-reduce bucket and is cleared by the optimizer kernel; ema_decay=0.95 is what the
        # reference's Trainer is built with (main.py:316)
        self.optimizer = FusedAdam(params, lr=lr, eps=1e-15, decay_iters=iters, ema_decay=ema_decay, world_size=world_size)
        self._prop_range = self.optimizer.range_of([*model.prop_encoders.parameters(), *model.prop_mlp.parameters()])

    def loss(self, rays_o, rays_d, gt_rgb, update_proposal=True, perturb=True):
        out = self.model.render(rays_o, rays_d, staged=False, bg_color=1, perturb=perturb,
                                update_proposal=update_proposal)
        loss = F.mse_loss(out["image"], gt_rgb, reduction="none").mean()
        if "proposal_loss" in out:
            loss = loss + self.model.opt.lambda_proposal * out["proposal_loss"]
        if "distort_loss" in out:
            loss = loss + self.model.opt.lambda_distort * out["distort_loss"]
        return loss, out

    def plan(self, n_rays):
        """The hand-scheduled, CUDA-graph-replayed step for ``n_rays`` rays (sanerf_b200/step.py), or None when the
        model is not the reference's stage-1 configuration (then the autograd path below runs)."""
        if not self.fused_step or not getattr(self.model, "tc_head", False) or not getattr(self.model, "fused", False):
            return None
        if n_rays not in self._plans:
            from .step import FusedRGBStep, UnsupportedConfig
            try:
                self._plans[n_rays] = FusedRGBStep(self.model, self.optimizer, n_rays, world_size=self.world_size,
                                                   use_graph=self.use_graph)
            except UnsupportedConfig:
                self._plans[n_rays] = None
        return self._plans[n_rays]

    def flush(self):
        """Apply the main-table update the hand-scheduled step defers to the start of the next step (call before
        reading parameters: checkpoint, evaluation, switching ray counts)."""
        for plan in self._plans.values():
            if plan is not None:
                plan.flush()

    def end_epoch(self):
        """What ``train_one_epoch`` does after its loop (nerf/utils.py:1862): one EMA update (when built with ``ema_decay``)."""
        self.flush()
        if self.optimizer.ema is not None:
            self.optimizer.ema_update()

    def step(self, rays_o, rays_d, gt_rgb):
        _check_equal_shards(self, rays_o.shape[0])
        plan = self.plan(rays_o.shape[0])
        for other in self._plans.values():                # a pending update of another ray count's plan comes first
            if other is not None and other is not plan:
                other.flush()
        if plan is not None:
            plan.global_step = self.global_step
            self.global_step += 1
            return plan(rays_o, rays_d, gt_rgb)
        self.flush()
        self.global_step += 1
        update_proposal = self.global_step <= 3000 or self.global_step % 5 == 0   # nerf/utils.py:910-911
        dev = self.optimizer.flat_param.device
        rays_o, rays_d, gt_rgb = (t.to(dev, non_blocking=True) for t in (rays_o, rays_d, gt_rgb))
        loss, _ = self.loss(rays_o, rays_d, gt_rgb, update_proposal)
        loss.backward()                                   # accumulates into the (pre-zeroed) flat gradient
        # steps that do not train the proposal networks leave their range out of the exchange and the update (the
        # reference's Adam skips parameters whose .grad is None), see FusedRGBStep._update_rest
        lo, hi = self._prop_range
        stop = self.optimizer.flat_param.numel() if (update_proposal or hi != self.optimizer.flat_param.numel()) else lo
        if self.world_size > 1:
            dist.all_reduce(self.optimizer.flat_grad[:stop], op=dist.ReduceOp.SUM)
        self.optimizer.schedule()
        self.optimizer.apply(0, stop, grad_scale=1.0 / self.world_size, zero_grad=True)
        return loss.detach()


def _check_equal_shards(trainer, n_rays):
    """The data-parallel gradient is (sum of per-rank means) / world: the global-batch mean only when every rank
    renders the same number of rays.  Checked once per ray count."""
    if trainer.world_size == 1 or n_rays in trainer._checked_counts:
        return
    dev = trainer.optimizer.flat_param.device
    t = torch.tensor([n_rays, -n_rays], device=dev, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(t[0]) != -int(t[1]):
        raise RuntimeError(f"ray shards differ across ranks ({-int(t[1])}..{int(t[0])} rays): pad or drop rays so that "
                           "every rank renders the same count (the gradient is averaged with equal rank weights)")
    trainer._checked_counts.add(n_rays)


class SAMTrainer:
    """Stage-2 step on frozen stage-1 parameters: render the [h,w] low-resolution rays' 256-d feature map and
    regress a target feature map (nerf/utils.py:1095-1106; the ViT-H target is replaced by a given tensor)."""

    def __init__(self, model, lr=1e-2, iters=5000, world_size=1, use_graph=True, fused_step=True, ema_decay=None):
        assert model.opt.with_sam
        self.model = model.train()
        self.world_size = world_size
        self._checked_counts = set()
        self.use_graph = bool(use_graph)
        self.fused_step = bool(fused_step)
        self._graphs = {}                                 # (n_rays, h, w, target shape) -> captured step
        self._plans = {}                                  # same key -> FusedSAMStep (hand-scheduled step) or None
        if all(p.requires_grad for p in model.parameters()):
            trainable = set()                             # no warm start applied: freeze what stage 1 trains, as
            for m in (model.s_grid, model.samvit_mlp):    # checkpoint.warm_start (main.py:255-262) would from its keys
                trainable.update(id(p) for p in m.parameters())
            for p in model.parameters():
                p.requires_grad_(id(p) in trainable)
        self.optimizer = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=lr, eps=1e-15,
                                   decay_iters=iters, ema_decay=ema_decay, world_size=world_size)

    def _forward_backward(self, rays_o, rays_d, target, h, w):
        out = self.model.render(rays_o, rays_d, staged=False, bg_color=1, perturb=False, update_proposal=False,
                                return_feats=1, H=h, W=w)
        pred = out["samvit"].permute(2, 0, 1).unsqueeze(0)
        if pred.shape[-2:] != target.shape[-2:]:
            pred = F.interpolate(pred, target.shape[-2:], mode="bilinear")
        loss = F.mse_loss(pred, target)
        loss.backward()
        return loss.detach()

    def step(self, rays_o, rays_d, target, h, w):
        """target: [1, 256, H_t, W_t]; prediction is bilinearly resized to it (nerf/utils.py:1100-1106).

        The step is autograd-driven (the 163 -> 256 x 5 SkipConnMLP + LayerNorm run on cuBLAS); after one eager call
        per input shape it is captured into a CUDA graph and replayed (static input buffers), which removes the
        ~150 host-side launches that otherwise dominate this 4096-ray step."""
        dev = self.optimizer.flat_param.device
        _check_equal_shards(self, rays_o.shape[0])
        key = (tuple(rays_o.shape), h, w, tuple(target.shape))
        plan = self.plan(rays_o.shape[0], h, w, tuple(target.shape))
        for other in self._plans.values():                # a pending table update of another shape's plan comes first
            if other is not None and other is not plan:
                other.flush()
        if plan is not None:
            return plan(rays_o, rays_d, target)
        entry = self._graphs.get(key)
        if not self.use_graph or entry is None:
            rays_o, rays_d, target = (t.to(dev, non_blocking=True) for t in (rays_o, rays_d, target))
            loss = self._forward_backward(rays_o, rays_d, target, h, w)
            self._update()
            if self.use_graph:
                self._graphs[key] = "warm"                # next call with this shape captures
            return loss
        if entry == "warm":
            static = [torch.empty_like(t, device=dev) for t in (rays_o, rays_d, target)]
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):                 # forward + backward; the NCCL exchange stays outside the graph
                loss = self._forward_backward(static[0], static[1], static[2], h, w)
                if self.world_size == 1:
                    self._update()
            entry = self._graphs[key] = (graph, static, loss)
        graph, static, loss = entry
        for buf, src in zip(static, (rays_o, rays_d, target)):
            buf.copy_(src, non_blocking=True)
        graph.replay()
        if self.world_size > 1:
            self._update()
        return loss

    def _update(self):
        if self.world_size > 1:
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM)
        self.optimizer.step(grad_scale=1.0 / self.world_size, zero_grad=True)

    def plan(self, n_rays, h, w, target_shape):
        """The hand-scheduled stage-2 step (sanerf_b200/step.py: FusedSAMStep) for this input shape, or None when the
        model is not the configuration it is written for (then the autograd path of ``step`` runs)."""
        if not self.fused_step or not getattr(self.model, "tc_head", False) or not getattr(self.model, "fused", False):
            return None
        key = (n_rays, h, w, tuple(target_shape))
        if key not in self._plans:
            from .step import FusedSAMStep, UnsupportedConfig
            try:
                self._plans[key] = FusedSAMStep(self.model, self.optimizer, n_rays, h, w, target_shape,
                                                world_size=self.world_size, use_graph=self.use_graph)
            except UnsupportedConfig:
                self._plans[key] = None
        return self._plans[key]

    def end_epoch(self):
        """What ``train_one_epoch`` does after its loop (nerf/utils.py:1862): one EMA update (when built with ``ema_decay``)."""
        self.flush()
        if self.optimizer.ema is not None:
            self.optimizer.ema_update()

    def flush(self):
        """Apply the s_grid update the hand-scheduled step defers to the start of the next step (call before reading
        parameters: checkpoint, evaluation)."""
        for plan in self._plans.values():
            if plan is not None:
                plan.flush()


_MAX_FRAME_PLANS = 4          # ray counts kept per model (static buffers + one CUDA graph each)


@torch.no_grad()
def render_frame(model, rays_o, rays_d, feat_rays_o=None, feat_rays_d=None, h=64, w=64, fused=True):
    """Interactive frame (nerf/utils.py:1647-1712): full-resolution RGB + depth, plus the low-resolution 256-d SAM
    feature map when feature rays are given.  The RGB pass is ONE CUDA-graph replay of the hand-scheduled forward
    (``FusedRGBFrame``) when the model has the reference's shapes, else the staged renderer (renderer.py:185-219).
    Runs in eval mode (aabb_infer, no jitter) and restores the caller's mode; the plans live on the model object
    (least-recently-used bound), so they die with it."""
    was_training = model.training
    model.eval()
    try:
        plan = None
        if fused and rays_o.is_cuda and getattr(model, "tc_head", False):
            from .step import FusedRGBFrame, UnsupportedConfig
            plans = model.__dict__.setdefault("_frame_plans", {})
            key = int(rays_o.shape[0])
            if key in plans:
                plans[key] = plans.pop(key)                   # most recently used last
            else:
                try:
                    plans[key] = FusedRGBFrame(model, key)
                except UnsupportedConfig:
                    plans[key] = None
                while len(plans) > _MAX_FRAME_PLANS:
                    plans.pop(next(iter(plans)))
            plan = plans[key]
        if plan is not None:
            out = plan(rays_o, rays_d)
        else:
            out = model.render(rays_o, rays_d, staged=True, bg_color=1, perturb=False)
        res = {"image": out["image"], "depth": out["depth"]}
        if "n_alive" in out:
            res["n_alive"] = out["n_alive"]
        if feat_rays_o is not None:
            f = model.render(feat_rays_o, feat_rays_d, staged=False, bg_color=1, perturb=False, return_feats=1, H=h, W=w)
            res["samvit"] = f["samvit"]
        return res
    finally:
        model.train(was_training)


from .parallel import gather_frame, shard_rays  # noqa: E402,F401  (re-exported: sharding helpers live in parallel.py)
