"""Hand-scheduled stage-1 (RGB) training step: ~25 kernel launches on static buffers, replayed as ONE CUDA graph.

Same computation as ``RGBTrainer.loss(...).backward(); optimizer.step()`` (the reference's ``Trainer.train_step`` +
``train_one_epoch`` iteration, nerf/utils.py:897-930, 1811-1836, with ``NeRFRenderer.run`` renderer.py:221-390
underneath), but without autograd, without torch glue kernels and without per-launch CPU work:

  level 0/1:  sample -> proposal density (encode + MLP + trunc_exp, one kernel) -> weights (composite, C = 0)
  level 2:    sample -> field head (gather + 3-layer MLP on tcgen05) -> trunc_exp + composite of the 15 geometry channels
              -> view head (SH, view MLP, sigmoid, background, MSE; forward AND backward in one kernel)
  losses:     proposal loss (2 levels) and distortion loss, each returning its gradient, pre-multiplied by lambda
  backward:   composite + trunc_exp -> field head (tcgen05) -> hash-grid scatter
              || (forked stream / parallel graph branch) proposal losses -> composite -> proposal density (x2)
  update:     [NCCL all-reduce when world_size > 1] -> fused Adam (clears the gradient): MLPs and proposal tables at the
              end of the step; the main hash table (89 % of the parameters) at the START of the next step, on a second
              stream, hidden behind jitter + the two proposal levels (which do not read it).  ``flush()`` applies it.

Every gradient kernel accumulates straight into the views of ``FusedAdam.flat_grad`` (pre-zeroed by the optimizer
kernel), so there is no ``zeros_like`` + add per parameter.  ``tests/test_gpu_step.py`` checks loss and every gradient
against the autograd path.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, fused
from .parallel import shard_bounds
from .fused import PRECISION_IDS, field_head_supported, prop_density_supported


class UnsupportedConfig(RuntimeError):
    """The model is not the reference's stage-1 configuration the hand-scheduled step is written for."""


def _update_priority(world_size):
    """Priority of the stream that carries the deferred table update.  One GPU: default (0), the Adam pass only fills
    what the high-priority chain leaves free.  Several GPUs: the stream also carries the NCCL reduce-scatter /
    all-gather, the longest pole before the field head, whose few CTAs must not queue behind the chain's kernels:
    high (-1) as well (measured at 2 GPUs, ms/step: chain -1 / update 0: 1.015; 0 / 0: 0.959; -1 / -1: 0.949)."""
    return int(os.environ.get("SANERF_UPD_PRIO", -1 if world_size > 1 else 0))


def _critical(plan, fn):
    """Run ``fn`` (the step's dependent chain) on the plan's HIGH-priority stream, forked from / joined to the current
    stream (works eagerly and under CUDA-graph capture, where kernel nodes inherit the priority).  The side branches
    (deferred table update, proposal backward) stay on default-priority streams: their CTAs fill the slots the chain
    leaves free instead of queueing ahead of it — without this the 1.2 GB s_grid Adam pass and the frozen forward it is
    meant to hide behind simply serialise (207 + 229 us -> 408 us measured)."""
    cur = torch.cuda.current_stream(plan.dev)
    hi = plan.critical_stream
    hi.wait_stream(cur)
    with torch.cuda.stream(hi):
        fn()
    cur.wait_stream(hi)


class FusedRGBStep:
    def __init__(self, model, optimizer, n_rays, world_size=1, use_graph=True, perturb=True, bg_color=1.0):
        opt = model.opt
        if opt.with_sam or opt.with_mask or opt.sum_after_mlp:
            raise UnsupportedConfig("FusedRGBStep covers the stage-1 RGB step (no SAM / mask heads, deferred shading)")
        if not field_head_supported(model.grid, model.grid_mlp):
            raise UnsupportedConfig("FusedRGBStep needs the reference's main grid + grid_mlp shapes")
        for enc, mlp in zip(model.prop_encoders, model.prop_mlp):
            if not prop_density_supported(enc, mlp):
                raise UnsupportedConfig("FusedRGBStep needs the reference's proposal network shapes")
        vw = [l.weight for l in model.view_mlp.net]
        if [tuple(w.shape) for w in vw] != [(32, 31), (32, 32), (3, 32)] or any(l.bias is not None for l in model.view_mlp.net):
            raise UnsupportedConfig("FusedRGBStep needs the reference's view_mlp (31 -> 32 -> 32 -> 3, no bias)")
        self.model, self.optimizer, self.world_size = model, optimizer, world_size
        self.N, self.steps = int(n_rays), [int(t) for t in opt.num_steps]
        if len(self.steps) != 3:
            raise UnsupportedConfig("FusedRGBStep expects two proposal levels and one final level")
        self.perturb, self.bg = bool(perturb), float(bg_color)
        self.precision = PRECISION_IDS[model.mlp_precision]
        self.opaque = int(opt.background == "last_sample")
        self.use_graph = bool(use_graph)
        dev = self.dev = next(model.parameters()).device
        N = self.N
        f32 = dict(device=dev, dtype=torch.float32)
        self.rays_o, self.rays_d, self.gt = (torch.zeros(N, 3, **f32) for _ in range(3))
        sizes = [N * (t + 1) for t in self.steps]
        self.noise_flat = torch.rand(sum(sizes), **f32)       # first step's jitter; later steps: sanerf_uniform_fill (Philox)
        self.rng_state = torch.zeros(4, device=dev, dtype=torch.int32)      # {call number, arrivals} of the jitter and of clear_loss
        rank = dist.get_rank() if (world_size > 1 and dist.is_initialized()) else 0
        self.seed = (int(torch.initial_seed()) ^ (rank * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF    # ranks draw different jitter
        self.noise, off = [], 0
        for t, n in zip(self.steps, sizes):
            self.noise.append(self.noise_flat[off:off + n].view(N, t + 1))
            off += n
        self.lv = []
        for li, t in enumerate(self.steps):
            self.lv.append(dict(T=t, bins=torch.empty(N, t + 1, **f32), t_mid=torch.empty(N, t, **f32),
                                deltas=torch.empty(N, t, **f32), x01=torch.empty(N, t, 3, **f32),
                                sigma=torch.empty(N, t, **f32), weights=torch.empty(N, t, **f32),
                                g_weights=torch.empty(N, t, **f32), g_sigma=torch.empty(N, t, **f32),
                                enc=(torch.empty(N * t, 2 * model.prop_encoders[li].num_levels, **f32) if li < 2 else None),
                                ws=torch.empty(N, **f32), depth=torch.empty(N, **f32)))
        B = N * self.steps[2]
        self.head = torch.empty(B, 16, **f32)
        self.g_head = torch.empty(B, 16, **f32)
        Bp = (B + 127) // 128 * 128             # saved activations are tile-chunk-major (fused.tcm_rows)
        self.enc, self.g_enc = torch.empty(Bp, 32, **f32), torch.empty(B, 32, **f32)
        self.h1, self.h2 = torch.empty(Bp, 64, **f32), torch.empty(Bp, 64, **f32)
        self.geo_sum, self.g_geo_sum = torch.empty(N, 15, **f32), torch.empty(N, 15, **f32)
        self.g_ws = torch.empty(N, **f32)
        self.n_alive = torch.empty(N, device=dev, dtype=torch.int32)
        self.image = torch.empty(N, 3, **f32)
        self.loss = torch.zeros(1, **f32)
        self.side_stream = torch.cuda.Stream(dev, priority=int(os.environ.get("SANERF_SIDE_PRIO", 0)))
        self.update_stream = torch.cuda.Stream(dev, priority=_update_priority(world_size))
        self.critical_stream = torch.cuda.Stream(dev, priority=int(os.environ.get("SANERF_CRIT_PRIO", -1)))
        self.distort_done = torch.cuda.Event()
        self.view_done = torch.cuda.Event()
        self.tail_stream = torch.cuda.Stream(dev, priority=int(os.environ.get("SANERF_SIDE_PRIO", 0)))
        self.fuse_tail = os.environ.get("SANERF_FUSE_TAIL", "1") == "1"      # multi-GPU: tail exchanges beside the backward
        # multi-GPU: with the fused symmetric-memory update there is no NCCL call in the step, and the whole step is ONE
        # graph per rank (no host launches between its phases); with the NCCL exchange (SANERF_SYMM=0) the forward /
        # backward halves are two graphs around the eager collectives unless SANERF_ONE_GRAPH=1 captures them as well
        self.one_graph = optimizer.symm is not None or os.environ.get("SANERF_ONE_GRAPH", "0") == "1"
        self.pending_main = False
        self.sharded_update = True
        self.graphs = {}
        self.eager_runs = {}
        self.global_step = 0
        for p in [model.grid.embeddings, *model.grid_mlp.parameters(), *model.view_mlp.parameters(),
                  *model.prop_encoders.parameters(), *model.prop_mlp.parameters()]:
            if p.grad is None or not p.grad.is_contiguous():
                raise RuntimeError("FusedRGBStep needs parameters registered with FusedAdam (flat gradient views)")
        self.prop_range = optimizer.range_of([*model.prop_encoders.parameters(), *model.prop_mlp.parameters()])
        if self.prop_range[1] != optimizer.flat_param.numel():
            raise UnsupportedConfig("FusedRGBStep expects the proposal networks at the tail of the flat parameter buffer")

    def _clear_loss(self):
        """loss <- 0 with the library's own kernel (no torch fill inside the captured step)."""
        with _lib.stats.span("clear_loss"):
            rc = _lib.load().sanerf_uniform_fill(self.loss.data_ptr(), 0, 0, self.rng_state.data_ptr() + 8, self.loss.data_ptr(), 1,
                                                 _lib.current_stream(self.dev))
        _lib.check(rc, "clear_loss")

    # ------------------------------------------------------------------------------------------------------
    def _launch(self, update_proposal):
        """Enqueue forward + backward on the current stream (eagerly, or under CUDA-graph capture)."""
        self._launch_front(update_proposal)
        self._launch_back(update_proposal)

    def _launch_front(self, update_proposal):
        """Jitter + the two proposal levels + the final level's sample positions: nothing here touches the main table."""
        m, lib, N = self.model, _lib.load(), self.N
        st = _lib.current_stream(self.dev)
        span, check, ptr = _lib.stats.span, _lib.check, _lib.ptr
        aabb = m.aabb_train if m.training else m.aabb_infer
        bound, contract, min_near = float(m.bound), int(bool(m.opt.contract)), float(m.min_near)
        self._clear_loss()         # (the jitter of THIS step was drawn during the previous step's backward, off the critical path)
        o, d = self.rays_o.data_ptr(), self.rays_d.data_ptr()

        # ---------------- forward: proposal levels
        for li in (0, 1, 2):
            L = self.lv[li]
            T = L["T"]
            noise = self.noise[li].data_ptr() if self.perturb else None
            if li == 0:
                with span("sample_uniform", N=N, T=T):
                    rc = lib.sanerf_sample_uniform(o, d, aabb.data_ptr(), min_near, None, 0, noise, N, T, contract, bound,
                                                   L["bins"].data_ptr(), L["t_mid"].data_ptr(), L["deltas"].data_ptr(),
                                                   L["x01"].data_ptr(), st)
                check(rc, "sample_uniform")
            else:                  # the previous level's compositing (sigma -> weights) is fused into the resampling kernel
                P = self.lv[li - 1]
                with span("sample_pdf", N=N, T=T):
                    rc = lib.sanerf_sample_pdf(o, d, aabb.data_ptr(), min_near, None, 0, P["bins"].data_ptr(), None, P["T"],
                                               noise, N, T, contract, bound, L["bins"].data_ptr(), L["t_mid"].data_ptr(),
                                               L["deltas"].data_ptr(), L["x01"].data_ptr(), P["sigma"].data_ptr(),
                                               P["deltas"].data_ptr(), self.opaque, P["weights"].data_ptr(), st)
                check(rc, "sample_pdf")
            if li < 2:
                enc, mlp = m.prop_encoders[li], m.prop_mlp[li]
                with span("prop_density_forward", B=N * T, L=enc.num_levels):
                    rc = lib.sanerf_prop_density_forward(L["x01"].data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(),
                                                         mlp.net[0].weight.data_ptr(), mlp.net[1].weight.data_ptr(), N * T,
                                                         enc.num_levels, float(np.log2(enc.per_level_scale)),
                                                         int(enc.base_resolution), L["sigma"].data_ptr(),
                                                         L["enc"].data_ptr() if update_proposal else None, st)
                check(rc, "prop_density_forward")

    def _launch_back(self, update_proposal, fuse_updates=False):
        """Final level forward, losses, backward of everything.  ``fuse_updates`` (multi-GPU with the symmetric-memory
        exchange): the small parameter ranges are exchanged + updated as soon as their gradients are complete, on side
        branches BESIDE the rest of the backward — [view_mlp, proposal networks] after the view head and the proposal
        branch, grid_mlp after the field-head backward, both while the hash-grid scatter (the last ~100 us of the step)
        still runs — so that no exchange, barrier wait or optimizer launch is left at the end of the step."""
        m, lib, N = self.model, _lib.load(), self.N
        st = _lib.current_stream(self.dev)
        span, check = _lib.stats.span, _lib.check
        lam_p = float(m.opt.lambda_proposal) if update_proposal else 0.0
        lam_d = float(m.opt.lambda_distort)
        d = self.rays_d.data_ptr()
        # ---------------- forward: final level
        L = self.lv[2]
        T, B = L["T"], N * L["T"]
        g, gm = m.grid, m.grid_mlp
        S, H = float(np.log2(g.per_level_scale)), int(g.base_resolution)
        w1, w2, w3 = (l.weight for l in gm.net)
        with span("field_head_forward", B=B):
            rc = lib.sanerf_field_head_forward(L["x01"].data_ptr(), g.embeddings.data_ptr(), g.offsets.data_ptr(), S, H, None,
                                               w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, self.enc.data_ptr(),
                                               self.h1.data_ptr(), self.h2.data_ptr(), self.head.data_ptr(), self.precision, st)
        check(rc, "field_head_forward")
        with span("head_composite_forward", N=N, T=T):
            rc = lib.sanerf_head_composite_forward(self.head.data_ptr(), L["deltas"].data_ptr(), L["t_mid"].data_ptr(), N, T,
                                                   self.opaque, float(m.t_thresh), None, L["weights"].data_ptr(),
                                                   L["ws"].data_ptr(), L["depth"].data_ptr(), self.geo_sum.data_ptr(),
                                                   self.n_alive.data_ptr(), st)
        check(rc, "head_composite_forward")
        # ---------------- fork: the proposal branch (proposal losses -> composite backward -> proposal-density backward of
        # both levels) depends only on the three weight tensors and shares nothing with the final-level chain below except
        # the loss scalar (atomic adds), so it runs on a second stream / as a parallel branch of the captured graph
        main = torch.cuda.current_stream(self.dev)
        have_gw2 = lam_d > 0
        side = self.side_stream
        fuse_updates = fuse_updates and self.optimizer.symm is not None
        forked = lam_p > 0 or have_gw2 or fuse_updates
        if forked:
            side.wait_stream(main)
        if self.perturb:                                   # next step's jitter ([N, T+1] uniforms per level, reference order)
            with torch.cuda.stream(side if forked else main):
                with span("uniform_fill", n=self.noise_flat.numel()):
                    rc = lib.sanerf_uniform_fill(self.noise_flat.data_ptr(), self.noise_flat.numel(), self.seed,
                                                 self.rng_state.data_ptr(), None, 0, _lib.current_stream(self.dev))
                check(rc, "uniform_fill")
        if have_gw2:                                       # distortion loss: concurrent with the view head on the main stream
            with torch.cuda.stream(side):
                with span("distortion_loss", N=N, T=T):
                    rc = lib.sanerf_distortion_loss(L["bins"].data_ptr(), L["weights"].data_ptr(), T, N, lam_d,
                                                    self.loss.data_ptr(), L["g_weights"].data_ptr(),
                                                    _lib.current_stream(self.dev))
                check(rc, "distortion_loss")
                self.distort_done.record(side)
        if lam_p > 0:
            with torch.cuda.stream(side):
                st2 = _lib.current_stream(self.dev)
                for li in (0, 1):
                    P = self.lv[li]
                    with span("proposal_loss", N=N, Tp=P["T"]):
                        rc = lib.sanerf_proposal_loss(L["bins"].data_ptr(), L["weights"].data_ptr(), T, P["bins"].data_ptr(),
                                                      P["weights"].data_ptr(), P["T"], N, lam_p, self.loss.data_ptr(),
                                                      P["g_weights"].data_ptr(), st2)
                    check(rc, "proposal_loss")
                # both (short) compositing backwards first: they fit into the window before the field-head backward takes every SM,
                # so that only the two density backwards are left for the tail behind the hash-grid scatter
                for li in (1, 0):
                    P = self.lv[li]
                    with span("composite_backward", N=N, T=P["T"], C=0):
                        rc = lib.sanerf_composite_backward(P["sigma"].data_ptr(), P["deltas"].data_ptr(), P["t_mid"].data_ptr(),
                                                           None, 0, None, N, P["T"], 0, self.opaque, 0.0, P["weights"].data_ptr(),
                                                           P["g_weights"].data_ptr(), None, None, None, P["g_sigma"].data_ptr(),
                                                           None, 0, st2)
                    check(rc, "composite_backward")
                for li in (1, 0):
                    P = self.lv[li]
                    Tp = P["T"]
                    enc, mlp = m.prop_encoders[li], m.prop_mlp[li]
                    with span("prop_density_backward", B=N * Tp, L=enc.num_levels):
                        rc = lib.sanerf_prop_density_backward(P["x01"].data_ptr(), enc.embeddings.data_ptr(),
                                                              enc.offsets.data_ptr(), mlp.net[0].weight.data_ptr(),
                                                              mlp.net[1].weight.data_ptr(), N * Tp, enc.num_levels,
                                                              float(np.log2(enc.per_level_scale)), int(enc.base_resolution),
                                                              P["enc"].data_ptr(), P["g_sigma"].data_ptr(),
                                                              enc.embeddings.grad.data_ptr(), mlp.net[0].weight.grad.data_ptr(),
                                                              mlp.net[1].weight.grad.data_ptr(), st2)
                    check(rc, "prop_density_backward")
        # ---------------- final level: view head + photometric loss (forward and backward), distortion loss
        v1, v2, v3 = (l.weight for l in m.view_mlp.net)
        with span("view_head", N=N):
            rc = lib.sanerf_view_head(self.geo_sum.data_ptr(), L["ws"].data_ptr(), d, self.gt.data_ptr(), v1.data_ptr(),
                                      v2.data_ptr(), v3.data_ptr(), self.bg, 1.0, N, self.image.data_ptr(), self.loss.data_ptr(),
                                      self.g_geo_sum.data_ptr(), self.g_ws.data_ptr(), v1.grad.data_ptr(), v2.grad.data_ptr(),
                                      v3.grad.data_ptr(), st)
        check(rc, "view_head")
        if fuse_updates:                                   # [view_mlp, proposal networks]: complete after view_head + the side branch
            self.view_done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(self.view_done)
                b1 = self.optimizer.ranges[id(m.grid_mlp.net[-1].weight)][1]
                n = self.optimizer.flat_param.numel() if update_proposal else self.prop_range[0]
                self.optimizer.apply_symm(b1, n, channel=2)
        # ---------------- backward: final level
        if have_gw2:
            main.wait_event(self.distort_done)
        with span("head_composite_backward", N=N, T=T):
            rc = lib.sanerf_head_composite_backward(self.head.data_ptr(), L["deltas"].data_ptr(), L["t_mid"].data_ptr(), N, T,
                                                    self.opaque, float(m.t_thresh),
                                                    L["g_weights"].data_ptr() if have_gw2 else None, self.g_ws.data_ptr(),
                                                    None, self.g_geo_sum.data_ptr(), None, self.g_head.data_ptr(), st)
        check(rc, "head_composite_backward")
        with span("field_head_backward", B=B):
            rc = lib.sanerf_field_head_backward(self.enc.data_ptr(), self.h1.data_ptr(), self.h2.data_ptr(), self.g_head.data_ptr(),
                                                w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, self.g_enc.data_ptr(), None, None,
                                                0.0, 0, None, w1.grad.data_ptr(), w2.grad.data_ptr(), w3.grad.data_ptr(),
                                                self.precision, st)
        check(rc, "field_head_backward")
        if fuse_updates:                                   # grid_mlp: complete now; exchanged beside the scatter
            tail = self.tail_stream
            tail.wait_stream(main)
            with torch.cuda.stream(tail):
                a1 = self._main_range()[1]
                self.optimizer.apply_symm(a1, self.optimizer.ranges[id(m.grid_mlp.net[-1].weight)][1], channel=1)
        with span("grid_encode_backward", B=B, L=16, C=2, D=3, half=False):
            rc = lib.sanerf_grid_encode_backward(self.g_enc.data_ptr(), L["x01"].data_ptr(), g.embeddings.data_ptr(),
                                                 g.offsets.data_ptr(), g.embeddings.grad.data_ptr(), B, 3, 2, 16, 16, S, H, None,
                                                 None, 0, 0, 0, _lib.SANERF_F32, _lib.LAYOUT_BLC, st)
        check(rc, "grid_encode_backward")
        if forked:
            main.wait_stream(self.side_stream)            # join
        if fuse_updates:
            main.wait_stream(self.tail_stream)

    # ---- optimizer.  The main hash table is 89 % of the parameters and nothing before the final level's field head reads
    # it, so its Adam update (and, multi-GPU, the all-reduce of its gradient) is DEFERRED to the start of the next step,
    # where it runs on a second stream concurrently with jitter + both proposal levels.  Everything else (MLPs, proposal
    # tables) is updated at the end of the step.  ``flush()`` applies a pending main-table update (call it before
    # reading the parameters: checkpoints, evaluation, tests).
    def _main_range(self):
        return self.optimizer.ranges[id(self.model.grid.embeddings)]

    def _update_main(self):
        """Update of the main table.  Multi-GPU: reduce-scatter of its gradient, Adam on this rank's 1/world shard only
        (so the 28 B/parameter of optimizer traffic shrink by the world size), all-gather of the updated shard — the same
        bytes on the wire as an all-reduce, and every rank ends with identical parameters.  (Measured at 8 GPUs:
        queueing the reduction right after the backward, or chunking it to overlap a full-size Adam, was not faster
        than running it at the start of the next step.)"""
        a, b = self._main_range()
        opt = self.optimizer
        if self.world_size == 1:
            opt.apply(a, b, grad_scale=1.0, zero_grad=True, gated=True)
            return
        if opt.symm is not None:                           # reduce + Adam + broadcast in ONE kernel over NVLink peer memory
            opt.apply_symm(a, b, gated=True)
            return
        world, rank = self.world_size, dist.get_rank()
        if os.environ.get("SANERF_DBG_SKIP_MAIN_NCCL"):    # timing diagnostics only (tools/ab_nccl_g8.sh): wrong gradients
            lo, hi = shard_bounds(a, b, world, rank)
            opt.apply(lo, hi, grad_scale=1.0 / world, zero_grad=True, gated=True)
            opt.flat_grad[a:b].zero_()
            return
        if not self.sharded_update:                        # plain all-reduce + full-size Adam (checker for the sharded form)
            dist.all_reduce(opt.flat_grad[a:b], op=dist.ReduceOp.SUM)
            opt.apply(a, b, grad_scale=1.0 / world, zero_grad=True, gated=True)
            return
        from .parallel import sharded_update
        opt.sharded[(a, b)] = shard_bounds(a, b, world, rank)
        sharded_update(opt.flat_param, opt.flat_grad, a, b,
                       lambda lo, hi: opt.apply(lo, hi, grad_scale=1.0 / world, zero_grad=True, gated=True), world, rank)

    def _update_rest(self, update_proposal=True):
        """MLPs and proposal tables, at the end of the step.  On steps that do not train the proposal networks
        (nerf/utils.py:910-911: 4 of 5 steps after step 3000) their parameters receive NO gradient in the reference
        (``zero_grad(set_to_none=True)`` leaves ``.grad`` None and torch's Adam skips them), so their range — the tail of
        the flat buffer — is left out of the exchange and of the update instead of moving by momentum."""
        a, b = self._main_range()
        n = self.optimizer.flat_param.numel()
        assert a == 0, "the main table is expected to lead the flat parameter buffer"
        if not update_proposal:
            n = self.prop_range[0]
        if self.optimizer.symm is not None:
            self.optimizer.apply_symm(b, n)
            return
        if self.world_size > 1 and not os.environ.get("SANERF_DBG_SKIP_TAIL_NCCL"):
            dist.all_reduce(self.optimizer.flat_grad[b:n], op=dist.ReduceOp.SUM)
        self.optimizer.apply(b, n, grad_scale=1.0 / self.world_size, zero_grad=True)

    def flush(self):
        """Apply a pending deferred table update now (checkpoint, evaluation, tests, switching plans).  The device-side
        gate is cleared afterwards, so the deferred pass at the start of the next step (baked into its CUDA graph) is a
        no-op instead of a momentum-only update on a zero gradient."""
        if self.pending_main:
            with torch.cuda.device(self.dev):
                self._update_main()
                self.optimizer.clear_gate()
            self.pending_main = False

    def gradients_only(self, rays_o, rays_d, gt, update_proposal=True):
        """Forward + backward without the optimizer update (tests): gradients are left in the flat bucket."""
        self.flush()
        self.rays_o.copy_(rays_o); self.rays_d.copy_(rays_d); self.gt.copy_(gt)
        with torch.cuda.device(self.dev):
            self._launch(update_proposal)
        return self.loss[0]

    def _whole_step(self, update_proposal):
        _critical(self, lambda: self._whole_step_body(update_proposal))

    def _whole_step_body(self, update_proposal):
        """One step on the current stream: deferred main-table update || front, then back, then the small update."""
        main = torch.cuda.current_stream(self.dev)
        upd = self.update_stream
        upd.wait_stream(main)
        with torch.cuda.stream(upd):
            self._update_main()                            # previous step's gradient (zero before the first step)
            self.optimizer.schedule()                      # then this step's learning rate / bias corrections
        self._launch_front(update_proposal)
        main.wait_stream(upd)
        # (starting the small ranges' all-reduces inside the backward was measured slower: NCCL's CTAs spin on SMs the
        # persistent one-CTA-per-SM field-head kernels count on, so they stay at the end of the step; a variant that
        # captures ONE all-reduce of the small ranges on the side branch beside the hash-grid scatter was also no faster
        # at 2 GPUs: 0.949-0.959 vs 0.942-0.944 ms.  At 8 GPUs removing this tail reduction saves 75 us and removing the
        # deferred main-table exchange another 74 us (tools/ab_nccl_g8.sh: 1.054 / 0.980 / 0.979 / 0.914 ms), but MOVING the
        # tail reduction beside the hash-grid scatter (a third graph for the scatter, NCCL on the update stream) changes
        # nothing: 1.0528 vs 1.0556 ms at 8 GPUs, 0.950 vs 0.949 at 2 - what the collectives absorb is rank skew.)
        fuse = self.fuse_tail and self.optimizer.symm is not None
        self._launch_back(update_proposal, fuse_updates=fuse)
        if not fuse:
            self._update_rest(update_proposal)

    def _graphs(self, update_proposal):
        """Single GPU: the whole step is one graph.  Multi-GPU: the forward / backward halves are two graphs and the
        NCCL exchanges stay eager between them.  Capturing them into ONE graph (``SANERF_ONE_GRAPH=1``) works and measures
        the same at 2 GPUs (0.947 / 0.955 vs 0.947 / 0.951 ms), so the eager form, which keeps NCCL's watchdog, is the default."""
        key = (bool(update_proposal), bool(self.model.training))    # the captured launches bake in aabb_train / aabb_infer
        if key not in self.graphs:
            if self.world_size == 1 or self.one_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._whole_step(update_proposal)
                self.graphs[key] = (g,)
            else:
                gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf):
                    _critical(self, lambda: self._launch_front(update_proposal))
                with torch.cuda.graph(gb):
                    _critical(self, lambda: self._launch_back(update_proposal))
                self.graphs[key] = (gf, gb)
        return self.graphs[key]

    def __call__(self, rays_o, rays_d, gt):
        """One training step; rays / targets may live on the host (pinned) or on the device.  Returns the loss
        (a view of a static device buffer: read it before the next call)."""
        self.global_step += 1
        update_proposal = self.global_step <= 3000 or self.global_step % 5 == 0      # nerf/utils.py:910-911
        self.rays_o.copy_(rays_o, non_blocking=True)
        self.rays_d.copy_(rays_d, non_blocking=True)
        self.gt.copy_(gt, non_blocking=True)
        with torch.cuda.device(self.dev):
            key = (bool(update_proposal), bool(self.model.training))
            graphed = self.use_graph and self.eager_runs.get(key, 0) >= 1   # first step of each variant runs eagerly (warm-up)
            if not graphed:
                self.eager_runs[key] = self.eager_runs.get(key, 0) + 1
                self._whole_step(update_proposal)
            elif self.world_size == 1 or self.one_graph:
                self._graphs(update_proposal)[0].replay()
            else:
                main = torch.cuda.current_stream(self.dev)
                upd = self.update_stream
                upd.wait_stream(main)
                with torch.cuda.stream(upd):
                    self._update_main()                    # reduce-scatter + Adam shard + all-gather, hidden behind the front
                    self.optimizer.schedule()
                gf, gb = self._graphs(update_proposal)
                gf.replay()
                main.wait_stream(upd)
                gb.replay()
                self._update_rest(update_proposal)
            self.pending_main = True
        return self.loss[0]


class FusedRGBFrame(FusedRGBStep):
    """Forward-only sibling of ``FusedRGBStep`` for inference: rays [N,3] -> image [N,3], depth [N], weights_sum [N] in
    12 launches (3 samplers, 2 proposal densities + 2 compositings, field head on tcgen05 without saved activations,
    fused trunc_exp + compositing, fused SH + view MLP + sigmoid + background), replayed as one CUDA graph.  It is what
    the interactive frame (nerf/utils.py:1647-1712 ``test_gui``, renderer.py:185-219 staged ``render``) reduces to for
    the RGB pass; a SAM model's feature branch is not touched (its outputs are not part of an RGB frame)."""

    def __init__(self, model, n_rays, use_graph=True, bg_color=1.0, chunked=True):
        opt = model.opt
        self.chunked = bool(chunked)       # with model.t_thresh > 0: evaluate the final level in chunks and skip terminated rays
        if opt.with_mask or opt.sum_after_mlp:
            raise UnsupportedConfig("FusedRGBFrame covers deferred shading without mask heads")
        if not field_head_supported(model.grid, model.grid_mlp):
            raise UnsupportedConfig("FusedRGBFrame needs the reference's main grid + grid_mlp shapes")
        for enc, mlp in zip(model.prop_encoders, model.prop_mlp):
            if not prop_density_supported(enc, mlp):
                raise UnsupportedConfig("FusedRGBFrame needs the reference's proposal network shapes")
        vw = [l.weight for l in model.view_mlp.net]
        if [tuple(w.shape) for w in vw] != [(32, 31), (32, 32), (3, 32)] or any(l.bias is not None for l in model.view_mlp.net):
            raise UnsupportedConfig("FusedRGBFrame needs the reference's view_mlp (31 -> 32 -> 32 -> 3, no bias)")
        self.model, self.world_size = model, 1
        self.N, self.steps = int(n_rays), [int(t) for t in opt.num_steps]
        if len(self.steps) != 3:
            raise UnsupportedConfig("FusedRGBFrame expects two proposal levels and one final level")
        self.perturb, self.bg = False, float(bg_color)
        self.precision = PRECISION_IDS[model.mlp_precision]
        self.opaque = int(opt.background == "last_sample")
        self.use_graph = bool(use_graph)
        dev = self.dev = next(model.parameters()).device
        N = self.N
        f32 = dict(device=dev, dtype=torch.float32)
        self.rays_o, self.rays_d = torch.zeros(N, 3, **f32), torch.zeros(N, 3, **f32)
        self.noise = [None] * 3
        self.lv = []
        for t in self.steps:
            self.lv.append(dict(T=t, bins=torch.empty(N, t + 1, **f32), t_mid=torch.empty(N, t, **f32),
                                deltas=torch.empty(N, t, **f32), x01=torch.empty(N, t, 3, **f32),
                                sigma=torch.empty(N, t, **f32), weights=torch.empty(N, t, **f32), enc=None,
                                ws=torch.empty(N, **f32), depth=torch.empty(N, **f32)))
        self.head = torch.empty(N * self.steps[2], 16, **f32)
        # per-ray compositing state in ONE buffer (cleared by one launch before a chunked frame):
        # {optical depth, weights_sum, depth, 15 channel sums, alive count, live-ray counts of chunks 1..}
        self.chunk_len = 8
        self.n_chunks = self.steps[2] // self.chunk_len if self.steps[2] % self.chunk_len == 0 else 0
        self.state = torch.zeros(N * 19 + 16, **f32)
        self.optical, self.ws_acc, self.depth_acc = self.state[:N], self.state[N:2 * N], self.state[2 * N:3 * N]
        self.geo_sum = self.state[3 * N:18 * N].view(N, 15)
        self.n_alive = self.state[18 * N:19 * N].view(torch.int32)
        self.counts = self.state[19 * N:19 * N + 16].view(torch.int32)          # [0] is unused: chunk 0 covers every ray
        self.all_rays = torch.arange(N, device=dev, dtype=torch.int32)
        self.n_rays_dev = torch.tensor([N], device=dev, dtype=torch.int32)
        self.lists = [torch.empty(N, device=dev, dtype=torch.int32) for _ in range(max(self.n_chunks - 1, 0))]
        self.image = torch.empty(N, 3, **f32)
        self.loss = torch.zeros(1, **f32)
        self.rng_state = torch.zeros(4, device=dev, dtype=torch.int32)
        self.graph = {}                                    # model.training -> captured forward (aabb_train / aabb_infer)
        self.eager_runs = 0

    def _launch_render(self):
        m, lib, N = self.model, _lib.load(), self.N
        self._launch_front(False)
        st = _lib.current_stream(self.dev)
        span, check = _lib.stats.span, _lib.check
        L = self.lv[2]
        T, B = L["T"], N * L["T"]
        g = m.grid
        w1, w2, w3 = (l.weight for l in m.grid_mlp.net)
        S, H = float(np.log2(g.per_level_scale)), int(g.base_resolution)
        ws_buf, depth_buf = L["ws"], L["depth"]
        if float(m.t_thresh) > 0.0 and self.n_chunks > 1 and self.chunked:
            # early ray termination that skips the field evaluation behind the termination point: chunks of 8 samples,
            # front to back, for the rays still alive (device-side work lists; no host synchronisation)
            ws_buf, depth_buf = self.ws_acc, self.depth_acc
            with span("clear_state", n=self.state.numel()):
                rc = lib.sanerf_uniform_fill(self.state.data_ptr(), 0, 0, self.rng_state.data_ptr() + 8, self.state.data_ptr(),
                                             self.state.numel(), st)
            check(rc, "clear_state")
            CL = self.chunk_len
            for c in range(self.n_chunks):
                rays = self.all_rays if c == 0 else self.lists[c - 1]
                count = self.n_rays_dev if c == 0 else self.counts[c:c + 1]
                last = c == self.n_chunks - 1
                with span("field_head_forward_chunk", N=N, chunk=c):
                    rc = lib.sanerf_field_head_forward_chunk(L["x01"].data_ptr(), g.embeddings.data_ptr(), g.offsets.data_ptr(), S, H,
                                                             w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), self.head.data_ptr(),
                                                             self.precision, rays.data_ptr(), count.data_ptr(), N, c, CL, T, st)
                check(rc, "field_head_forward_chunk")
                with span("head_composite_chunk", N=N, chunk=c):
                    rc = lib.sanerf_head_composite_chunk(self.head.data_ptr(), L["deltas"].data_ptr(), L["t_mid"].data_ptr(),
                                                         rays.data_ptr(), count.data_ptr(), N, T, c, CL, self.opaque,
                                                         float(m.t_thresh), None if last else self.lists[c].data_ptr(),
                                                         None if last else self.counts[c + 1:c + 2].data_ptr(),
                                                         self.optical.data_ptr(), ws_buf.data_ptr(), depth_buf.data_ptr(),
                                                         self.geo_sum.data_ptr(), self.n_alive.data_ptr(), st)
                check(rc, "head_composite_chunk")
        else:
            with span("field_head_forward", B=B):
                rc = lib.sanerf_field_head_forward(L["x01"].data_ptr(), g.embeddings.data_ptr(), g.offsets.data_ptr(), S, H, None,
                                                   w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, None, None, None,
                                                   self.head.data_ptr(), self.precision, st)
            check(rc, "field_head_forward")
            with span("head_composite_forward", N=N, T=T):
                rc = lib.sanerf_head_composite_forward(self.head.data_ptr(), L["deltas"].data_ptr(), L["t_mid"].data_ptr(), N, T,
                                                       self.opaque, float(m.t_thresh), None, L["weights"].data_ptr(),
                                                       L["ws"].data_ptr(), L["depth"].data_ptr(), self.geo_sum.data_ptr(),
                                                       self.n_alive.data_ptr(), st)
            check(rc, "head_composite_forward")
        self._out_ws, self._out_depth = ws_buf, depth_buf
        v1, v2, v3 = (l.weight for l in m.view_mlp.net)
        with span("view_head", N=N):
            rc = lib.sanerf_view_head(self.geo_sum.data_ptr(), ws_buf.data_ptr(), self.rays_d.data_ptr(), None, v1.data_ptr(),
                                      v2.data_ptr(), v3.data_ptr(), self.bg, 1.0, N, self.image.data_ptr(), None, None, None,
                                      None, None, None, st)
        check(rc, "view_head")

    @torch.no_grad()
    def __call__(self, rays_o, rays_d):
        """Returns views of static buffers (image [N,3], depth [N], weights_sum [N], n_alive [N]): copy before the next call."""
        self.rays_o.copy_(rays_o, non_blocking=True)
        self.rays_d.copy_(rays_d, non_blocking=True)
        with torch.cuda.device(self.dev):
            if self.use_graph and self.eager_runs >= 1:
                mode = bool(self.model.training)
                if mode not in self.graph:
                    self.graph[mode] = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph[mode]):
                        self._launch_render()
                self.graph[mode].replay()
            else:
                self.eager_runs += 1
                self._launch_render()
        return {"image": self.image, "depth": self._out_depth, "weights_sum": self._out_ws, "n_alive": self.n_alive}


class FusedSAMStep:
    """Hand-scheduled stage-2 (SAM feature field) training step (nerf/utils.py:1095-1106 + renderer.py:221-390 with
    ``--with_sam``; stage-1 parameters frozen, main.py:255-262):

      frozen front:  the forward-only RGB plan (``FusedRGBFrame``: samplers, proposal densities, field head on tcgen05,
                     fused compositing, view head) -> final-level samples x01, weights, geo_sum, image, depth
      feature field: ONE kernel  f_sam[r] = sum_i w[r,i] * s_grid(x01[r,i])  (no [N*T,128] feature matrix)
      samvit head:   f = [f_sam, f_image, image, depth] -> samvit_mlp -> resize -> MSE; autograd over the 5-layer
                     SkipConnMLP + LayerNorm only (f is a leaf), gradients accumulate into the flat bucket
      scatter:       ONE kernel  g_table += cw * (w[r,i] * g_f[r,:128])  straight into ``FusedAdam.flat_grad`` (no
                     compositing backward, no gradient matrix, no zeros_like + accumulate pass over the 168 MB table)
      update:        samvit_mlp at the end of the step; the s_grid table (99 % of the parameters, 1.2 GB of optimizer
                     traffic) at the START of the next step on a second stream, hidden behind the frozen front, which
                     does not read it (multi-GPU: reduce-scatter + Adam on the local shard + all-gather there).

    One CUDA graph per step on one GPU; two graphs around the eager NCCL calls otherwise.  ``flush()`` applies a
    pending table update (call before reading parameters)."""

    def __init__(self, model, optimizer, n_rays, h, w, target_shape, world_size=1, use_graph=True):
        from .fused import ray_features_supported
        opt = model.opt
        if not opt.with_sam or opt.sum_after_mlp or opt.with_mask:
            raise UnsupportedConfig("FusedSAMStep covers the stage-2 step (--with_sam, deferred shading, no mask heads)")
        if not ray_features_supported(model.s_grid):
            raise UnsupportedConfig("FusedSAMStep needs a 3-D fp32 hash s_grid with 2, 4 or 8 features per level")
        if h * w != n_rays:
            raise UnsupportedConfig("FusedSAMStep renders one h x w feature map per step")
        self.model, self.optimizer, self.world_size = model, optimizer, world_size
        self.frame = FusedRGBFrame(model, n_rays, use_graph=False, bg_color=1.0, chunked=False)      # launched inside this step's graph
        self.N, self.h, self.w = int(n_rays), int(h), int(w)
        self.use_graph = bool(use_graph)
        dev = self.dev = self.frame.dev
        f32 = dict(device=dev, dtype=torch.float32)
        g = model.s_grid
        self.f_sam = torch.empty(self.N, g.num_levels * g.level_dim, **f32)
        self.target = torch.empty(tuple(target_shape), **f32)
        self.loss = torch.zeros(1, **f32)
        self.update_stream = torch.cuda.Stream(dev, priority=_update_priority(world_size))
        self.side_stream = torch.cuda.Stream(dev)          # weight-gradient GEMMs of the samvit head
        self.copy_stream = torch.cuda.Stream(dev)          # upload of a host target beside the frozen front
        self.copy_done = torch.cuda.Event()
        self.critical_stream = torch.cuda.Stream(dev, priority=int(os.environ.get("SANERF_CRIT_PRIO", -1)))
        self.pending_main = False
        self.sharded_update = True
        self.graphs = {}                                   # model.training -> captured step
        self.eager_runs = 0
        self.last_f = None
        a, b = self._main_range()
        if a != 0 or g.embeddings.grad is None or not g.embeddings.grad.is_contiguous():
            raise RuntimeError("FusedSAMStep needs s_grid.embeddings to lead a FusedAdam flat buffer")
        self._head_params = [p for p in model.samvit_mlp.parameters()]
        # Optional (SANERF_SAM_SPLIT=<level>, default off): split the s_grid update at a level boundary - the scatter runs as
        # two launches (levels complete in order), the Adam pass / exchange of the FIRST part starts as soon as its gradient
        # is complete, beside the scatter of the remaining levels, and only the second part is deferred to the next step's
        # front.  Measured: no gain on one GPU (0.9085 vs 0.9069 ms at level 8) and slower at two (1.097 / 1.164 ms at
        # levels 8 / 10 vs 1.060): the update's CTAs take register-file slots from the scatter as they do from the front.
        split_level = int(os.environ.get("SANERF_SAM_SPLIT", 0))
        nccl_path = world_size > 1 and optimizer.symm is None
        self.split_level = split_level if (0 < split_level < g.num_levels and not nccl_path) else 0
        self.split_at = a + int(g.offsets[self.split_level].item()) * g.level_dim if self.split_level else a
        self.scatter_done = torch.cuda.Event()
        # the autograd-free samvit head (static buffers), when the model is the reference's configuration on the tensor cores
        mlp, ln = model.samvit_mlp[0], model.samvit_mlp[1]
        self.head = None
        feat_ok = g.num_levels * g.level_dim == 128 and (opt.sam_use_view_direction or False)
        if (getattr(mlp, "tc", False) and feat_ok and tuple(target_shape[-2:]) == (self.h, self.w)
                and all(p.grad is not None for p in self._head_params) and not torch.is_autocast_enabled()):
            try:
                self.head = fused.SamvitHead(mlp, ln, self.N, dev)
            except ValueError:
                self.head = None

    def _main_range(self):
        return self.optimizer.ranges[id(self.model.s_grid.embeddings)]

    # ------------------------------------------------------------------------------------------------------
    def _launch_front(self):
        """Frozen stage-1 forward: nothing here reads s_grid."""
        with torch.no_grad():
            if self.head is not None:                      # padded copies of two samvit weights, beside the front
                cur = torch.cuda.current_stream(self.dev)
                self.side_stream.wait_stream(cur)
                with torch.cuda.stream(self.side_stream):
                    self.head.refresh_weights()
            self.frame._launch_render()
            if self.head is not None:
                cur.wait_stream(self.side_stream)          # joined here: the front may be captured as a graph of its own
            if self.head is None:
                self.sh = self.model.view_encoder(self.frame.rays_d)                  # [N,16], once per ray

    def _launch_back(self, fuse_updates=False):
        m, lib, N, fr = self.model, _lib.load(), self.N, self.frame
        span, check = _lib.stats.span, _lib.check
        L = fr.lv[2]
        T = L["T"]
        g = m.s_grid
        S, H, C, nl = float(np.log2(g.per_level_scale)), int(g.base_resolution), int(g.level_dim), int(g.num_levels)
        st = _lib.current_stream(self.dev)
        ws, depth = L["ws"], L["depth"]
        head = self.head
        if head is not None:
            # samvit head without autograd and without torch glue kernels: the features and the rest of the input row are
            # written straight into the head's skip buffer (fused.SamvitHead), five tensor-core GEMMs forward, LayerNorm +
            # MSE forward/backward in one kernel, per layer a weight-gradient GEMM reducing into the flat gradient views + a
            # column sum (bias) on the side stream and a data-gradient GEMM with the activation derivative in its epilogue
            with span("ray_features_forward", N=N, T=T, C=C):
                rc = lib.sanerf_ray_features_forward(L["x01"].data_ptr(), L["weights"].data_ptr(), g.embeddings.data_ptr(),
                                                     g.offsets.data_ptr(), N, T, C, nl, S, H, head.f.data_ptr(), head.LD, st)
            check(rc, "ray_features_forward")
            with span("sam_pack", N=N):
                rc = lib.sanerf_sam_pack(fr.geo_sum.data_ptr(), ws.data_ptr(), fr.rays_d.data_ptr(), fr.image.data_ptr(),
                                         depth.data_ptr(), N, int(bool(m.opt.sam_use_view_direction)),
                                         head.f.data_ptr() + 4 * nl * C, head.LD, st)
            check(rc, "sam_pack")
            head.forward()
            self._clear_loss()
            head.loss_backward(self.target, self.loss)
            g_sam = head.backward(self.side_stream, nl * C)
            g_stride = head.LD
            self.samvit, self.last_f = head.samvit, head.f
        else:
            with span("ray_features_forward", N=N, T=T, C=C):
                rc = lib.sanerf_ray_features_forward(L["x01"].data_ptr(), L["weights"].data_ptr(), g.embeddings.data_ptr(),
                                                     g.offsets.data_ptr(), N, T, C, nl, S, H, self.f_sam.data_ptr(), 0, st)
            check(rc, "ray_features_forward")
            if m.opt.sam_use_view_direction:                                         # renderer.py:380 / :383
                parts = [self.f_sam, fr.geo_sum, ws.unsqueeze(-1) * self.sh, fr.image, depth.unsqueeze(-1)]
            else:
                parts = [self.f_sam, fr.geo_sum, fr.image, depth.unsqueeze(-1)]
            f = torch.cat(parts, dim=-1)
            f.requires_grad_(True)
            with torch.enable_grad():
                samvit = m.samvit_mlp(f)
                pred = samvit.view(self.h, self.w, -1).permute(2, 0, 1).unsqueeze(0)
                if pred.shape[-2:] != self.target.shape[-2:]:
                    pred = torch.nn.functional.interpolate(pred, self.target.shape[-2:], mode="bilinear")
                loss = torch.nn.functional.mse_loss(pred, self.target)
            loss.backward(inputs=[f, *self._head_params])  # one traversal; parameter gradients accumulate into the flat views
            self.loss.copy_(loss.detach().reshape(1))
            self.last_f, self.samvit = f.detach(), samvit.detach()
            g_sam = f.grad[:, :nl * C].contiguous()
            g_stride = 0
        st = _lib.current_stream(self.dev)
        main = torch.cuda.current_stream(self.dev)
        cut = self.split_level if fuse_updates else 0
        for lo, hi in (((0, cut), (cut, nl)) if cut else ((0, nl),)):
            with span("ray_features_backward", N=N, T=T, C=C, levels=hi - lo):
                rc = lib.sanerf_ray_features_backward(L["x01"].data_ptr(), L["weights"].data_ptr(), g_sam.data_ptr(),
                                                      g.offsets.data_ptr(), N, T, C, nl, S, H, g.embeddings.grad.data_ptr(), lo, hi,
                                                      g_stride, st)
            check(rc, "ray_features_backward")
            if cut and hi == cut:                          # levels [0, cut) are complete: their Adam pass starts now
                self.scatter_done.record(main)
                with torch.cuda.stream(self.update_stream):
                    self.update_stream.wait_event(self.scatter_done)
                    self._update_early()
        if cut:
            main.wait_stream(self.update_stream)
        if head is not None:                               # weight-gradient GEMMs ran beside the data-gradient chain + scatter
            main.wait_stream(self.side_stream)

    def _clear_loss(self):
        with _lib.stats.span("clear_loss"):
            rc = _lib.load().sanerf_uniform_fill(self.loss.data_ptr(), 0, 0, self.frame.rng_state.data_ptr(), self.loss.data_ptr(), 1,
                                                 _lib.current_stream(self.dev))
        _lib.check(rc, "clear_loss")

    # ---- optimizer ----------------------------------------------------------------------------------------
    def _update_early(self):
        """Adam of the table's first part [a, split_at) — issued on the update stream beside the scatter of the later levels."""
        a, _ = self._main_range()
        opt = self.optimizer
        if self.split_at <= a:
            return
        if opt.symm is not None:
            opt.apply_symm(a, self.split_at, channel=1)
        else:
            opt.apply(a, self.split_at, grad_scale=1.0, zero_grad=True)

    def _update_main(self):
        a, b = self._main_range()
        a = self.split_at                                  # the deferred part (everything when the split is off)
        opt = self.optimizer
        if self.world_size == 1:
            opt.apply(a, b, grad_scale=1.0, zero_grad=True, gated=True)
            return
        if opt.symm is not None:
            opt.apply_symm(a, b, gated=True)
            return
        world, rank = self.world_size, dist.get_rank()
        if not self.sharded_update:
            dist.all_reduce(opt.flat_grad[a:b], op=dist.ReduceOp.SUM)
            opt.apply(a, b, grad_scale=1.0 / world, zero_grad=True, gated=True)
            return
        from .parallel import sharded_update
        opt.sharded[(a, b)] = shard_bounds(a, b, world, rank)
        sharded_update(opt.flat_param, opt.flat_grad, a, b,
                       lambda lo, hi: opt.apply(lo, hi, grad_scale=1.0 / world, zero_grad=True, gated=True), world, rank)

    def _update_rest(self):
        b, n = self._main_range()[1], self.optimizer.flat_param.numel()
        if self.optimizer.symm is not None:
            self.optimizer.apply_symm(b, n, channel=2)
            return
        if self.world_size > 1:
            dist.all_reduce(self.optimizer.flat_grad[b:n], op=dist.ReduceOp.SUM)
        self.optimizer.apply(b, n, grad_scale=1.0 / self.world_size, zero_grad=True)

    def flush(self):
        """Apply a pending deferred table update now (checkpoint, evaluation, tests, switching plans).  The device-side
        gate is cleared afterwards, so the deferred pass at the start of the next step (baked into its CUDA graph) is a
        no-op instead of a momentum-only update on a zero gradient."""
        if self.pending_main:
            with torch.cuda.device(self.dev):
                self._update_main()
                self.optimizer.clear_gate()
            self.pending_main = False

    def _deferred_update(self):
        main, upd = torch.cuda.current_stream(self.dev), self.update_stream
        upd.wait_stream(main)
        with torch.cuda.stream(upd):
            self._update_main()                            # previous step's table gradient (zero before the first step)
            self.optimizer.schedule()                      # then this step's learning rate / bias corrections
        return main, upd

    def _whole_step(self):
        _critical(self, self._whole_step_body)

    def _whole_step_body(self):
        main, upd = self._deferred_update()
        self._launch_front()
        main.wait_stream(upd)
        self._launch_back(fuse_updates=True)
        self._update_rest()

    def gradients_only(self, rays_o, rays_d, target):
        """Forward + backward without any optimizer update (tests): gradients are left in the flat bucket."""
        self.flush()
        self.frame.rays_o.copy_(rays_o); self.frame.rays_d.copy_(rays_d); self.target.copy_(target)
        with torch.cuda.device(self.dev):
            self._launch_front()
            self._launch_back()
        return self.loss[0]

    def _front_half(self):
        """deferred s_grid update || frozen stage-1 forward (nothing here reads the target feature map)."""
        def body():
            main, upd = self._deferred_update()
            self._launch_front()
            main.wait_stream(upd)
        _critical(self, body)

    def _back_half(self):
        def body():
            self._launch_back(fuse_updates=True)
            self._update_rest()
        _critical(self, body)

    def __call__(self, rays_o, rays_d, target):
        """One training step; inputs may live on the host (pinned) or the device.  Returns the loss (static buffer).
        A HOST target ([1,256,h,w]: 4 MB at 64 x 64) is uploaded on a copy stream while the frozen front runs — the step is
        then replayed as two graphs around the wait for that copy; a device target is copied in stream order and the step
        is one graph."""
        fr = self.frame
        fr.rays_o.copy_(rays_o, non_blocking=True)
        fr.rays_d.copy_(rays_d, non_blocking=True)
        cur = torch.cuda.current_stream(self.dev)
        split = (not target.is_cuda) and self.use_graph and self.eager_runs >= 1 and \
            (self.world_size == 1 or self.optimizer.symm is not None)
        if split:
            cs = self.copy_stream
            cs.wait_stream(cur)                            # the previous step has finished reading self.target
            with torch.cuda.stream(cs):
                self.target.copy_(target, non_blocking=True)
                self.copy_done.record(cs)
        else:
            self.target.copy_(target, non_blocking=True)
        with torch.cuda.device(self.dev):
            if not self.use_graph or self.eager_runs < 1:
                self.eager_runs += 1
                self._whole_step()
            elif self.world_size == 1 or self.optimizer.symm is not None:
                key = (bool(self.model.training), split)
                if key not in self.graphs:
                    if split:
                        gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gf):
                            self._front_half()
                        with torch.cuda.graph(gb):
                            self._back_half()
                        self.graphs[key] = (gf, gb)
                    else:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._whole_step()
                        self.graphs[key] = (g,)
                if split:
                    self.graphs[key][0].replay()
                    cur.wait_event(self.copy_done)
                    self.graphs[key][1].replay()
                else:
                    self.graphs[key][0].replay()
            else:
                mode = bool(self.model.training)
                if mode not in self.graphs:
                    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gf):
                        _critical(self, self._launch_front)
                    with torch.cuda.graph(gb):
                        _critical(self, self._launch_back)
                    self.graphs[mode] = (gf, gb)
                main, upd = self._deferred_update()
                self.graphs[mode][0].replay()
                main.wait_stream(upd)
                self.graphs[mode][1].replay()
                self._update_rest()
            self.pending_main = True
        return self.loss[0]
