"""The hand-scheduled CUDA-graph training step (sanerf_b200/step.py) against the autograd path of the same model."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n_rays, seed=0):
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import RGBTrainer, default_opt
    torch.manual_seed(seed)
    model = NeRFNetwork(default_opt()).cuda()
    with torch.no_grad():   # a non-trivial field: the default +-1e-4 tables give near-constant outputs
        for enc in [model.grid, *model.prop_encoders]:
            offs = enc.offsets.tolist()
            for l in range(len(offs) - 1):
                enc.embeddings[offs[l]:offs[l + 1]].uniform_(-0.5, 0.5).mul_(1.0 / enc.per_level_scale ** l)
    g = torch.Generator().manual_seed(seed + 1)
    o = (torch.rand(n_rays, 3, generator=g) - 0.5).cuda()
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).cuda()
    gt = torch.rand(n_rays, 3, generator=g).cuda()
    return model, RGBTrainer, o, d, gt


@pytest.mark.parametrize("update_proposal", [True, False])
def test_fused_step_gradients_match_autograd(cuda, update_proposal):
    from sanerf_b200.step import FusedRGBStep
    model, RGBTrainer, o, d, gt = _setup(200)
    trainer = RGBTrainer(model, fused_step=False)
    loss_ref, out = trainer.loss(o, d, gt, update_proposal=update_proposal, perturb=False)
    loss_ref.backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    trainer.optimizer.zero_grad()
    plan = FusedRGBStep(model, trainer.optimizer, 200, use_graph=False, perturb=False)
    loss = plan.gradients_only(o, d, gt, update_proposal=update_proposal)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss, loss_ref.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(plan.image, out["image"].detach(), rtol=1e-5, atol=1e-6)
    for n, p in model.named_parameters():
        a, b = p.grad, ref[n]
        if not update_proposal and n.startswith("prop_"):
            assert float(a.abs().max()) == 0.0, n
            continue
        scale = b.abs().max().item()
        assert scale > 0, n
        # same kernels on the table / field-head path (atomic order only); the view head re-associates fp32 sums
        assert ((a - b).norm() / b.norm()).item() < 1e-4, n
        torch.testing.assert_close(a, b, rtol=1e-3, atol=2e-4 * scale, msg=lambda m, n=n: f"{n}: {m}")


def test_fused_step_graph_replay_equals_eager(cuda):
    """Two identical models, same rays, no jitter: three steps of graph replay == three eager steps."""
    from sanerf_b200.step import FusedRGBStep
    model_a, RGBTrainer, o, d, gt = _setup(256, seed=3)
    model_b = copy.deepcopy(model_a)
    ta, tb = RGBTrainer(model_a), RGBTrainer(model_b)
    pa = FusedRGBStep(model_a, ta.optimizer, 256, use_graph=True, perturb=False)
    pb = FusedRGBStep(model_b, tb.optimizer, 256, use_graph=False, perturb=False)
    for _ in range(4):
        la, lb = pa(o, d, gt).clone(), pb(o, d, gt).clone()
        assert torch.isfinite(la)
        torch.testing.assert_close(la, lb, rtol=2e-4, atol=1e-6)
    pa.flush(); pb.flush()          # the main-table update of the last step is deferred to the next one
    assert (True, True) in pa.graphs and int(ta.optimizer.step_count) == 4
    for (n, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
        assert ((p - q).norm() / q.norm().clamp_min(1e-12)).item() < 1e-3, n


def test_trainer_uses_fused_step_and_learns(cuda):
    model, RGBTrainer, o, d, gt = _setup(512, seed=5)
    trainer = RGBTrainer(model)
    assert trainer.plan(512) is not None
    gt = gt * 0 + torch.tensor([0.2, 0.5, 0.8], device="cuda")
    losses = [float(trainer.step(o, d, gt)) for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses
    before = model.grid.embeddings.detach().clone()
    trainer.flush()                 # applies the deferred main-table update exactly once
    assert not torch.equal(before, model.grid.embeddings)
    after = model.grid.embeddings.detach().clone()
    trainer.flush()
    assert torch.equal(after, model.grid.embeddings)


@pytest.mark.parametrize("with_sam", [False, True])
def test_fused_frame_equals_staged_render(cuda, with_sam):
    """Forward-only hand-scheduled frame (one CUDA graph) == the renderer's staged evaluation, RGB / depth / weights_sum."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.step import FusedRGBFrame
    from sanerf_b200.train import default_opt, render_frame
    torch.manual_seed(2)
    model = NeRFNetwork(default_opt(with_sam=with_sam, max_ray_batch=300)).cuda().eval()
    with torch.no_grad():
        for enc in [model.grid, *model.prop_encoders]:
            enc.embeddings.uniform_(-0.5, 0.5)
    g = torch.Generator().manual_seed(9)
    o = (torch.rand(1000, 3, generator=g) - 0.5).cuda()
    d = (torch.randn(1000, 3, generator=g)).cuda()                      # un-normalised, as get_rays produces them
    with torch.no_grad():
        ref = model.render(o, d, staged=True, bg_color=1, perturb=False)
    plan = FusedRGBFrame(model, 1000)
    for _ in range(3):                                                   # eager, capture, replay
        out = plan(o, d)
        torch.testing.assert_close(out["image"], ref["image"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out["depth"], ref["depth"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out["weights_sum"], ref["weights_sum"], rtol=1e-5, atol=1e-6)
    assert False in plan.graph                 # captured in eval mode (aabb_infer)
    res = render_frame(model, o, d, o[:64] if with_sam else None, d[:64] if with_sam else None, 8, 8)
    torch.testing.assert_close(res["image"], ref["image"], rtol=1e-5, atol=1e-6)
    if with_sam:
        assert res["samvit"].shape == (8, 8, 256)


def test_sam_trainer_graph_replay_equals_eager(cuda):
    """Stage-2 step (autograd + cuBLAS MLP) captured as a CUDA graph == the same step run eagerly."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import SAMTrainer, default_opt
    torch.manual_seed(4)
    model_a = NeRFNetwork(default_opt(with_sam=True)).cuda()
    with torch.no_grad():
        for enc in [model_a.grid, *model_a.prop_encoders, model_a.s_grid]:
            enc.embeddings.uniform_(-0.5, 0.5)
    model_b = copy.deepcopy(model_a)
    ta, tb = SAMTrainer(model_a, use_graph=True, fused_step=False), SAMTrainer(model_b, use_graph=False, fused_step=False)
    g = torch.Generator().manual_seed(5)
    o = (torch.rand(64, 3, generator=g) - 0.5).cuda()
    d = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1).cuda()
    target = torch.randn(1, 256, 8, 8, generator=g).cuda()
    for i in range(5):
        la, lb = ta.step(o, d, target, 8, 8).clone(), tb.step(o, d, target, 8, 8).clone()
        torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-6)
    assert any(isinstance(v, tuple) for v in ta._graphs.values()) and int(ta.optimizer.step_count) == 5
    for (n, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
        assert ((p - q).norm() / q.norm().clamp_min(1e-12)).item() < 1e-3, n
    assert not model_a.grid.embeddings.requires_grad and model_a.s_grid.embeddings.requires_grad


# ------------------------------------------------------------------------------------- stage 2 (SAM feature field)
def _setup_sam(h, w, seed=0):
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import default_opt
    torch.manual_seed(seed)
    model = NeRFNetwork(default_opt(with_sam=True)).cuda()
    with torch.no_grad():
        for enc in [model.grid, model.s_grid, *model.prop_encoders]:
            offs = enc.offsets.tolist()
            for l in range(len(offs) - 1):
                enc.embeddings[offs[l]:offs[l + 1]].uniform_(-0.5, 0.5).mul_(1.0 / enc.per_level_scale ** l)
    n = h * w
    g = torch.Generator().manual_seed(seed + 1)
    o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    target = torch.randn(1, 256, h, w, generator=g).cuda()
    return model, o, d, target


@pytest.mark.parametrize("C", [2, 4, 8])
def test_ray_features_equal_encode_then_weighted_sum(cuda, C):
    """fused.ray_features == sum_i w_i * GridEncoder(x_i) (renderer.py:302-303, 377), forward and table gradient."""
    from gridencoder import GridEncoder
    from sanerf_b200 import fused
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=8, level_dim=C, base_resolution=8, log2_hashmap_size=12,
                      desired_resolution=256).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    N, T = 300, 32
    o = torch.rand(N, 1, 3, device="cuda") * 0.6 + 0.2
    step = torch.randn(N, 1, 3, device="cuda") * 0.01
    x01 = (o + step * torch.arange(T, device="cuda").view(1, T, 1)).clamp(0, 1).contiguous()   # ray-ordered: shared cells
    x01[5, 3] = 1.5                                                                            # out of range: encodes to zero
    w = torch.rand(N, T, device="cuda")
    w[:, 20:] = 0.0                                                                            # terminated samples
    g_out = torch.randn(N, 8 * C, device="cuda")

    got = fused.ray_features(x01, w, enc)
    got.backward(g_out)
    g_got = enc.embeddings.grad.clone()
    enc.embeddings.grad = None
    feats = enc(x01 * 2 - 1, bound=1)                                                          # maps back to [0,1]
    exp = (w.unsqueeze(-1) * feats).sum(-2)
    exp.backward(g_out)
    torch.testing.assert_close(got, exp, rtol=1e-4, atol=1e-5)
    g_exp = enc.embeddings.grad
    assert ((g_got - g_exp).norm() / g_exp.norm()).item() < 1e-5
    torch.testing.assert_close(g_got, g_exp, rtol=1e-3, atol=1e-5 * float(g_exp.abs().max()))


def test_fused_sam_step_gradients_match_autograd(cuda):
    from sanerf_b200.step import FusedSAMStep
    from sanerf_b200.train import SAMTrainer
    h = w = 12
    model, o, d, target = _setup_sam(h, w)
    trainer = SAMTrainer(model, fused_step=False, use_graph=False)
    loss_ref = trainer._forward_backward(o, d, target, h, w)
    ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    assert set(n.split(".")[0] for n in ref) == {"s_grid", "samvit_mlp"}
    trainer.optimizer.zero_grad()
    plan = FusedSAMStep(model, trainer.optimizer, h * w, h, w, target.shape, use_graph=False)
    loss = plan.gradients_only(o, d, target)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss, loss_ref, rtol=1e-5, atol=1e-7)
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        a, b = p.grad, ref[n]
        assert float(b.abs().max()) > 0, n
        assert ((a - b).norm() / b.norm()).item() < 1e-4, n


def test_fused_sam_step_graph_replay_equals_autograd_steps(cuda):
    """Four optimizer steps: hand-scheduled step replayed as one CUDA graph == the autograd SAM step."""
    from sanerf_b200.train import SAMTrainer
    h = w = 16
    model_a, o, d, target = _setup_sam(h, w, seed=2)
    model_b = copy.deepcopy(model_a)
    ta, tb = SAMTrainer(model_a), SAMTrainer(model_b, fused_step=False, use_graph=False)
    assert ta.plan(h * w, h, w, tuple(target.shape)) is not None
    la = lb = None
    for _ in range(4):
        la, lb = ta.step(o, d, target, h, w).clone(), tb.step(o, d, target, h, w).clone()
        torch.testing.assert_close(la, lb, rtol=2e-4, atol=1e-6)
    host_target = target.cpu().pin_memory()                 # a host target: uploaded beside the frozen front, two graphs
    for _ in range(2):
        la, lb = ta.step(o, d, host_target, h, w).clone(), tb.step(o, d, target, h, w).clone()
        torch.testing.assert_close(la, lb, rtol=2e-4, atol=1e-6)
    assert (True, True) in ta._plans[(h * w, h, w, tuple(target.shape))].graphs
    ta.flush()
    assert (True, False) in ta._plans[(h * w, h, w, tuple(target.shape))].graphs
    for (n, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
        assert ((p - q).norm() / q.norm().clamp_min(1e-12)).item() < 1e-3, n


def test_chunked_frame_with_termination_matches_unchunked(cuda):
    """Early ray termination that skips work (front-to-back chunks of 8 samples, device-side live-ray lists) == the
    un-chunked frame with the same threshold (same rule): image, depth, weights_sum, and the alive counts
    exactly; with threshold ~0 it reproduces the frame without termination."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.step import FusedRGBFrame
    from sanerf_b200.train import default_opt
    torch.manual_seed(2)
    model = NeRFNetwork(default_opt()).cuda().eval()
    with torch.no_grad():
        for enc in [model.grid, *model.prop_encoders]:
            enc.embeddings.uniform_(-0.5, 0.5)
        model.grid_mlp.net[2].weight[0].abs_().mul_(30.0)            # strong densities: rays terminate inside the 32 samples
    g = torch.Generator().manual_seed(9)
    n = 3000
    o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
    d = torch.randn(n, 3, generator=g).cuda()
    for thresh in (1e-2, 1e-4, 1e-30):
        model.t_thresh = thresh
        ref = {k: v.clone() for k, v in FusedRGBFrame(model, n, chunked=False)(o, d).items()}
        plan = FusedRGBFrame(model, n, chunked=True)
        for _ in range(3):                                           # eager, capture, replay
            out = plan(o, d)
            torch.testing.assert_close(out["image"], ref["image"], rtol=1e-5, atol=2e-6)
            torch.testing.assert_close(out["depth"], ref["depth"], rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(out["weights_sum"], ref["weights_sum"], rtol=1e-5, atol=2e-6)
            same = (out["n_alive"] == ref["n_alive"]).float().mean().item()
            assert same > 0.999, same                                # prefix sums associate differently: ties at the threshold only
        if thresh == 1e-2:
            assert float(ref["n_alive"].float().mean()) < 30         # the scene really terminates rays
        live = plan.counts[1:plan.n_chunks].tolist()
        assert all(a >= b for a, b in zip([n] + live, live)), live   # live-ray lists shrink front to back
    model.t_thresh = 0.0
