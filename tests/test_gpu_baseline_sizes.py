"""Parity of the hand-scheduled steps AT THE SIZES BASELINE.json names (configs[1]: 8192 rays x 128/64/32 samples;
configs[2]: 64x64 rays, 256-d target), not only at the small sizes of tests/test_gpu_step.py:

* against the repository's own autograd path (same kernels underneath, different scheduling), and
* against the ORACLE (oracle/render_torch.py: the statement-by-statement torch restatement of nerf/renderer.py:221-390 +
  nerf/network.py:221-259 + nerf/utils.py:897-930, 1095-1106) evaluated with plain torch ops on the same device, from the
  same state_dict — loss, image, and every gradient.

Tolerances: BASELINE.json north_star (1e-3 relative fp32 outputs; atomic-order gradients 1e-4 relative against the same
kernels; against the oracle the sample positions differ by ulps — torch cumsum / searchsorted vs the sampler kernel — so
table gradients are compared in the L2 norm, as in tests/test_gpu_render.py)."""
import pytest
import torch

from oracle import render_torch as R

pytestmark = pytest.mark.gpu


def _smooth_tables(mods, grid_type):
    with torch.no_grad():
        for mod in mods:
            if isinstance(mod, grid_type):
                offs = mod.offsets.tolist()
                for l in range(len(offs) - 1):
                    mod.embeddings[offs[l]:offs[l + 1]].uniform_(-0.5, 0.5).mul_(1.0 / mod.per_level_scale ** l)


def _pair(with_sam, seed):
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import default_opt
    torch.manual_seed(seed)
    ref = R.NeRFNetworkRef(with_sam=with_sam)
    _smooth_tables(ref.modules(), R.GridEncoderRef)
    model = NeRFNetwork(default_opt(with_sam=with_sam))
    missing, unexpected = model.load_state_dict(ref.state_dict(), strict=False)
    assert not [m for m in missing if "aabb_infer" not in m] and not unexpected, (missing, unexpected)
    return ref.cuda().train(), model.cuda().train()


def _rays(n, seed):
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    return o, d, g


def _assert_grads_vs_oracle(model, ref, names=None):
    ref_p = dict(ref.named_parameters())
    for name, p in model.named_parameters():
        if not p.requires_grad or (names is not None and name not in names):
            continue
        got, exp = p.grad.detach().double(), ref_p[name].grad.detach().double()
        assert float(exp.abs().max()) > 0, name
        rel = ((got - exp).norm() / exp.norm()).item()
        assert rel < (5e-3 if name.endswith("embeddings") else 3e-3), f"{name}: relative L2 error {rel:.3e}"


def test_rgb_step_8192_rays_matches_autograd_and_oracle(cuda):
    """configs[1]: 8192 rays, 2^20 / 2^19 / 2^18 samples per level."""
    from sanerf_b200.step import FusedRGBStep
    from sanerf_b200.train import RGBTrainer
    n = 8192
    ref, model = _pair(False, seed=11)
    o, d, g = _rays(n, 12)
    gt = torch.rand(n, 3, generator=g).cuda()

    trainer = RGBTrainer(model, fused_step=False)
    loss_auto, out_auto = trainer.loss(o, d, gt, update_proposal=True, perturb=False)
    loss_auto.backward()
    auto = {k: p.grad.clone() for k, p in model.named_parameters()}
    trainer.optimizer.zero_grad()

    plan = FusedRGBStep(model, trainer.optimizer, n, use_graph=False, perturb=False)
    loss = plan.gradients_only(o, d, gt, update_proposal=True)
    torch.cuda.synchronize()
    # (i) same kernels, autograd-scheduled: outputs to rounding, gradients to atomic order
    torch.testing.assert_close(loss, loss_auto.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(plan.image, out_auto["image"].detach(), rtol=1e-5, atol=1e-6)
    for k, p in model.named_parameters():
        assert ((p.grad - auto[k]).norm() / auto[k].norm()).item() < 1e-4, k

    # (ii) the oracle on the same rays / parameters
    loss_ref, out_ref = ref.rgb_loss(o, d, gt, update_proposal=True, perturb=False)
    loss_ref.backward()
    torch.testing.assert_close(loss, loss_ref.detach(), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(plan.image, out_ref["image"].detach(), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(plan.lv[2]["weights"], out_ref["weights"].detach(), rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(plan.lv[2]["depth"], out_ref["depth"].detach(), rtol=1e-3, atol=1e-3)
    _assert_grads_vs_oracle(model, ref)


def test_rgb_step_1024_rays_jittered_matches_oracle(cuda):
    """Same comparison with the training jitter on: the oracle consumes the very uniforms the step's static noise
    buffers hold (reference order: [N, T+1] per level, renderer.py:269 and :101)."""
    from sanerf_b200.step import FusedRGBStep
    from sanerf_b200.train import RGBTrainer
    n = 1024
    ref, model = _pair(False, seed=21)
    o, d, g = _rays(n, 22)
    gt = torch.rand(n, 3, generator=g).cuda()
    trainer = RGBTrainer(model, fused_step=False)
    plan = FusedRGBStep(model, trainer.optimizer, n, use_graph=False, perturb=True)
    noise = [t.clone() for t in plan.noise]
    loss = plan.gradients_only(o, d, gt, update_proposal=True)
    torch.cuda.synchronize()
    loss_ref, out_ref = ref.rgb_loss(o, d, gt, update_proposal=True, perturb=True, noise=noise)
    loss_ref.backward()
    torch.testing.assert_close(loss, loss_ref.detach(), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(plan.image, out_ref["image"].detach(), rtol=1e-3, atol=1e-4)
    _assert_grads_vs_oracle(model, ref)


def test_sam_step_64x64_matches_autograd_and_oracle(cuda):
    """configs[2]: 4096 rays (64 x 64), s_grid L16 F8 T2^19, samvit_mlp 163 -> 256 x 5 + LayerNorm, [1,256,64,64] target."""
    from sanerf_b200.step import FusedSAMStep
    from sanerf_b200.train import SAMTrainer
    h = w = 64
    ref, model = _pair(True, seed=31)
    o, d, g = _rays(h * w, 32)
    target = torch.randn(1, 256, h, w, generator=g).cuda()

    trainer = SAMTrainer(model, fused_step=False, use_graph=False)
    loss_auto = trainer._forward_backward(o, d, target, h, w)
    auto = {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}
    assert {k.split(".")[0] for k in auto} == {"s_grid", "samvit_mlp"}
    trainer.optimizer.zero_grad()

    plan = FusedSAMStep(model, trainer.optimizer, h * w, h, w, target.shape, use_graph=False)
    loss = plan.gradients_only(o, d, target)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss, loss_auto, rtol=1e-5, atol=1e-7)
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert ((p.grad - auto[k]).norm() / auto[k].norm()).item() < 1e-4, k

    res = ref.run(o, d, perturb=False, update_proposal=False, return_feats=1, H=h, W=w)
    pred = res["samvit"].permute(2, 0, 1).unsqueeze(0)
    loss_ref = torch.nn.functional.mse_loss(pred, target)                       # nerf/utils.py:1100-1106
    names = [k for k in auto]
    grads = torch.autograd.grad(loss_ref, [dict(ref.named_parameters())[k] for k in names])
    for k, gr in zip(names, grads):
        dict(ref.named_parameters())[k].grad = gr
    torch.testing.assert_close(loss, loss_ref.detach(), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(plan.samvit, res["samvit"].detach().reshape(h * w, 256), rtol=1e-3, atol=2e-4)
    _assert_grads_vs_oracle(model, ref, names=set(names))


def test_deferred_update_is_idempotent_after_flush(cuda):
    """step, flush, step == step, step: a flush applies the pending main-table update early and the deferred pass baked
    into the next step's graph must then be a no-op (not a momentum-only move on a zero gradient)."""
    import copy

    from sanerf_b200.train import RGBTrainer
    from tests.test_gpu_step import _setup
    model_a, _, o, d, gt = _setup(256, seed=7)
    model_b = copy.deepcopy(model_a)
    ta, tb = RGBTrainer(model_a), RGBTrainer(model_b)
    for t in (ta, tb):
        t.plan(256).perturb = False
    for i in range(6):
        la, lb = ta.step(o, d, gt).clone(), tb.step(o, d, gt).clone()
        torch.testing.assert_close(la, lb, rtol=2e-4, atol=1e-6)
        if i % 2 == 0:
            ta.flush()                         # e.g. a checkpoint / evaluation between steps
    ta.flush(); tb.flush()
    # (Adam with eps = 1e-15 amplifies atomic-order noise of near-zero gradients: 1.2e-3 measured after these six steps
    # for two correct runs; a momentum-only pass moves every touched entry by ~lr = 1e-2, i.e. ||diff|| / ||table|| ~ 7e-2)
    for (n, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
        assert ((p - q).norm() / q.norm().clamp_min(1e-12)).item() < 5e-3, n
    # and the optimizer state: a spurious pass would have decayed exp_avg of the main table by 0.9 per flush
    a, b = ta.optimizer.ranges[id(model_a.grid.embeddings)]
    ma, mb = ta.optimizer.exp_avg[a:b], tb.optimizer.exp_avg[a:b]
    assert ((ma - mb).norm() / mb.norm()).item() < 2e-2


def test_proposal_networks_do_not_move_without_gradient(cuda):
    """Steps with update_proposal=False (nerf/utils.py:910-911) leave the proposal networks and their Adam state alone,
    as torch's Adam does for parameters whose .grad is None."""
    from sanerf_b200.train import RGBTrainer
    from tests.test_gpu_step import _setup
    model, _, o, d, gt = _setup(256, seed=9)
    tr = RGBTrainer(model)
    for _ in range(3):
        tr.step(o, d, gt)
    tr.flush()
    tr.global_step = 3001                      # -> update_proposal only every 5th step
    before = {n: p.detach().clone() for n, p in model.named_parameters() if n.startswith("prop_")}
    lo, hi = tr._prop_range
    m_before = tr.optimizer.exp_avg[lo:hi].clone()
    main_before = model.grid_mlp.net[0].weight.detach().clone()
    for _ in range(3):                         # global steps 3002, 3003, 3004: none divisible by 5
        tr.step(o, d, gt)
    tr.flush()
    for n, p in model.named_parameters():
        if n.startswith("prop_"):
            assert torch.equal(p, before[n]), n
    assert torch.equal(tr.optimizer.exp_avg[lo:hi], m_before)
    assert not torch.equal(model.grid_mlp.net[0].weight, main_before)


def test_ema_follows_torch_ema_semantics(cuda):
    """EMA(0.95) as torch_ema.ExponentialMovingAverage keeps it (nerf/utils.py:616): shadow -= (1 - min(decay, (1+k)/(10+k))) *
    (shadow - param) at every ``update()`` — which the reference calls once per EPOCH (nerf/utils.py:1862) = ``end_epoch()``
    here; ``ema_every_step`` folds the same update into every Adam pass, including the deferred main table's."""
    from sanerf_b200.fused import FusedAdam
    from sanerf_b200.train import RGBTrainer
    from tests.test_gpu_step import _setup
    model, _, o, d, gt = _setup(128, seed=13)
    tr = RGBTrainer(model, ema_decay=0.95)
    opt = tr.optimizer
    shadow = opt.flat_param.clone()
    for k in range(1, 4):                                  # three "epochs" of two steps
        tr.step(o, d, gt); tr.step(o, d, gt)
        torch.testing.assert_close(opt.ema, shadow, rtol=1e-5, atol=1e-7)   # steps alone do not touch the shadow
        tr.end_epoch()
        decay = min(0.95, (1 + k) / (10 + k))
        shadow -= (1 - decay) * (shadow - opt.flat_param)
        torch.testing.assert_close(opt.ema, shadow, rtol=1e-5, atol=1e-7)
    state = opt.ema_state_dict()
    assert state["num_updates"] == 3 and len(state["shadow_params"]) == len(opt.params)
    live = opt.flat_param.clone()
    opt.ema_store(); opt.ema_copy_to()
    torch.testing.assert_close(model.grid.embeddings.reshape(-1), shadow[:model.grid.embeddings.numel()], rtol=1e-5, atol=1e-7)
    opt.ema_restore()
    assert torch.equal(opt.flat_param, live)
    # per-step variant: the update rides in the Adam kernels
    model2, _, o, d, gt = _setup(128, seed=14)
    tr2 = RGBTrainer(model2)
    tr2.optimizer = FusedAdam([p for p in model2.parameters()], lr=1e-2, eps=1e-15, decay_iters=20000, ema_decay=0.95,
                              ema_every_step=True)
    tr2._prop_range = tr2.optimizer.range_of([*model2.prop_encoders.parameters(), *model2.prop_mlp.parameters()])
    tr2._plans.clear()
    opt2 = tr2.optimizer
    shadow = opt2.flat_param.clone()
    for t in range(1, 4):
        tr2.step(o, d, gt); tr2.flush()
        decay = min(0.95, (1 + t) / (10 + t))
        shadow -= (1 - decay) * (shadow - opt2.flat_param)
        torch.testing.assert_close(opt2.ema, shadow, rtol=1e-5, atol=1e-7)
