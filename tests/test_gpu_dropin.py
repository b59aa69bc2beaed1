"""Drop-in proof (INTEGRATION.md path (a)).

The reference's OWN Python modules (oracle/_ref_py: unmodified copies staged by oracle/stage_ref_py.py) are run in a child
interpreter, from one reference-format checkpoint, twice:

  * with ``_gridencoder`` / ``_shencoder`` / ``_freqencoder`` resolved to THIS repository's shims, and
  * with them resolved to the reference's CUDA extensions rebuilt for sm_100 (oracle/_ref) — the true reference result,

and both are compared with this repository's ``NeRFNetwork`` / hand-scheduled steps loaded from the same checkpoint:
``render()`` outputs, losses and every gradient.  Also: reference-format checkpoint round trip (nerf/utils.py:2041-2166)
and the freeze-by-key warm start of main.py:255-262.
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import stage_ref_py

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "dropin_child.py")


def _smooth(model):
    with torch.no_grad():
        for enc in [m for m in model.modules() if hasattr(m, "embeddings") and hasattr(m, "offsets")]:
            offs = enc.offsets.tolist()
            for l in range(len(offs) - 1):
                enc.embeddings[offs[l]:offs[l + 1]].uniform_(-0.5, 0.5).mul_(1.0 / enc.per_level_scale ** l)


def _run_child(backend, state_path, out_path, with_sam, rays=512, hw=16):
    if not stage_ref_py.available():
        pytest.skip("oracle/_ref_py not staged (python -m oracle.stage_ref_py where /root/reference exists)")
    if backend == "ref" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "_ref_gridencoder.so")):
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, CHILD, "--backend", backend, "--state", state_path, "--out", out_path,
                        "--with-sam", str(int(with_sam)), "--rays", str(rays), "--hw", str(hw)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return torch.load(out_path, map_location="cpu")


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def stage1(cuda, tmp_path_factory):
    """A stage-1 model of ours, saved in the reference's checkpoint layout."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.checkpoint import save_checkpoint
    from sanerf_b200.train import default_opt
    torch.manual_seed(0)
    model = NeRFNetwork(default_opt()).cuda()
    _smooth(model)
    d = tmp_path_factory.mktemp("dropin")
    path = str(d / "ngp_ep0001.pth")
    save_checkpoint(path, model, epoch=1)
    ck = torch.load(path)
    ck["ray_seed"] = 77
    torch.save(ck, path)
    return model, path, d


def test_reference_python_runs_on_our_shims_rgb(cuda, stage1):
    from sanerf_b200.step import FusedRGBStep
    from sanerf_b200.train import RGBTrainer
    model, path, d = stage1
    n = 512
    on_ours = _run_child("ours", path, str(d / "rgb_ours.pt"), False, rays=n)
    on_ref = _run_child("ref", path, str(d / "rgb_ref.pt"), False, rays=n)
    assert on_ours["keys"] == list(model.state_dict().keys())        # same state_dict keys, same order

    # (1) the reference's Python on our kernels == the reference's Python on the reference's kernels
    for k in ("image", "depth", "weights_sum", "weights"):
        torch.testing.assert_close(on_ours[k], on_ref[k], rtol=1e-3, atol=1e-5, msg=lambda m, k=k: f"{k}: {m}")
    torch.testing.assert_close(on_ours["loss"], on_ref["loss"], rtol=1e-4, atol=1e-7)
    for k, g in on_ref["grads"].items():
        assert _rel(on_ours["grads"][k], g) < 1e-4, k                # same torch glue, same positions: atomic order only

    # (2) this repository's model and hand-scheduled step from the same checkpoint == the true reference result
    g = torch.Generator().manual_seed(77)
    o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
    dr = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    gt = torch.rand(n, 3, generator=g).cuda()
    model.train()
    trainer = RGBTrainer(model, fused_step=False)
    plan = FusedRGBStep(model, trainer.optimizer, n, use_graph=False, perturb=False)
    loss = plan.gradients_only(o, dr, gt, update_proposal=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.cpu(), on_ref["loss"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(plan.image.cpu(), on_ref["image"], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(plan.lv[2]["depth"].cpu(), on_ref["depth"], rtol=1e-3, atol=1e-3)
    for k, p in model.named_parameters():
        rel = _rel(p.grad.cpu(), on_ref["grads"][k])
        assert rel < (5e-3 if k.endswith("embeddings") else 2e-3), f"{k}: {rel:.3e}"
    trainer.optimizer.zero_grad()


def test_reference_python_runs_on_our_shims_sam(cuda, stage1, tmp_path):
    """Stage 2: warm start from the stage-1 checkpoint (main.py:255-262), 16x16 feature rays."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.checkpoint import save_checkpoint, warm_start
    from sanerf_b200.step import FusedSAMStep
    from sanerf_b200.train import SAMTrainer, default_opt
    _, stage1_path, _ = stage1
    torch.manual_seed(5)
    model = NeRFNetwork(default_opt(with_sam=True))
    frozen = warm_start(model, stage1_path)
    assert {k.split(".")[0] for k in frozen} == {"grid", "grid_mlp", "view_mlp", "prop_encoders", "prop_mlp"}
    assert all(not p.requires_grad for k, p in model.named_parameters() if k in frozen)
    assert all(p.requires_grad for k, p in model.named_parameters() if k not in frozen)
    model = model.cuda()
    with torch.no_grad():
        offs = model.s_grid.offsets.tolist()
        for l in range(len(offs) - 1):
            model.s_grid.embeddings[offs[l]:offs[l + 1]].uniform_(-0.5, 0.5).mul_(1.0 / model.s_grid.per_level_scale ** l)
    path = str(tmp_path / "sam_ep0001.pth")
    save_checkpoint(path, model, epoch=1)
    ck = torch.load(path)
    ck["ray_seed"] = 78
    torch.save(ck, path)

    hw = 16
    on_ours = _run_child("ours", path, str(tmp_path / "sam_ours.pt"), True, hw=hw)
    on_ref = _run_child("ref", path, str(tmp_path / "sam_ref.pt"), True, hw=hw)
    torch.testing.assert_close(on_ours["samvit"], on_ref["samvit"], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(on_ours["loss"], on_ref["loss"], rtol=1e-4, atol=1e-7)

    g = torch.Generator().manual_seed(78)
    o = (torch.rand(hw * hw, 3, generator=g) - 0.5).cuda()
    dr = torch.nn.functional.normalize(torch.randn(hw * hw, 3, generator=g), dim=-1).cuda()
    target = torch.randn(1, 256, hw, hw, generator=g).cuda()
    trainer = SAMTrainer(model, use_graph=False)                       # respects the warm start's requires_grad flags
    assert {k.split(".")[0] for k, p in model.named_parameters() if p.requires_grad} == {"s_grid", "samvit_mlp"}
    plan = FusedSAMStep(model, trainer.optimizer, hw * hw, hw, hw, target.shape, use_graph=False)
    loss = plan.gradients_only(o, dr, target)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.cpu(), on_ref["loss"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(plan.samvit.cpu().view(hw, hw, 256), on_ref["samvit"], rtol=1e-3, atol=2e-4)
    for k, p in model.named_parameters():
        if p.requires_grad:
            # the reference back-propagates into the frozen stage-1 field as well; only the trained keys are compared
            rel = _rel(p.grad.cpu(), on_ref["grads"][k])
            assert rel < (5e-3 if k.endswith("embeddings") else 2e-3), f"{k}: {rel:.3e}"


def test_reference_format_checkpoint_round_trip(cuda, tmp_path):
    """save (full=True) -> load into a fresh model + trainer -> identical parameters, Adam moments, step count and EMA;
    the continued run matches an uninterrupted one."""
    import copy

    from nerf.network import NeRFNetwork
    from sanerf_b200.checkpoint import load_checkpoint, save_checkpoint
    from sanerf_b200.train import RGBTrainer, default_opt
    torch.manual_seed(3)
    model = NeRFNetwork(default_opt()).cuda()
    _smooth(model)
    twin = copy.deepcopy(model)
    g = torch.Generator().manual_seed(4)
    o = (torch.rand(256, 3, generator=g) - 0.5).cuda()
    dr = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=-1).cuda()
    gt = torch.rand(256, 3, generator=g).cuda()

    a = RGBTrainer(model, ema_decay=0.95)
    a.plan(256).perturb = False
    for _ in range(3):
        a.step(o, dr, gt)
    a.end_epoch()                                           # the per-epoch EMA update (nerf/utils.py:1862)
    path = str(tmp_path / "ngp_ep0003.pth")
    save_checkpoint(path, model, a, epoch=3, full=True)
    ck = torch.load(path)
    assert set(ck) >= {"epoch", "global_step", "stats", "model", "optimizer", "lr_scheduler", "ema"}
    assert list(ck["model"].keys()) == list(model.state_dict().keys()) and ck["global_step"] == 3
    assert set(ck["optimizer"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}

    b = RGBTrainer(twin, ema_decay=0.95)
    b.plan(256).perturb = False
    missing, unexpected = load_checkpoint(path, twin, b)
    assert not missing and not unexpected
    for (k, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        assert torch.equal(p, q), k
    assert torch.equal(a.optimizer.exp_avg, b.optimizer.exp_avg) and torch.equal(a.optimizer.ema, b.optimizer.ema)
    assert int(b.optimizer.step_count) == 3 and b.global_step == 3 and b.optimizer.ema_updates == 1
    assert not torch.equal(a.optimizer.ema, a.optimizer.flat_param)
    for _ in range(2):
        la, lb = a.step(o, dr, gt).clone(), b.step(o, dr, gt).clone()
        torch.testing.assert_close(la, lb, rtol=2e-4, atol=1e-6)
    a.flush(); b.flush()
    for (k, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        assert _rel(p, q) < 2e-3, k                        # atomic-order noise through Adam's sign-like early updates


TRAINER_CHILD = os.path.join(ROOT, "tests", "dropin_trainer_child.py")


def _run_trainer_child(mode, state_path, out_path, workspace, steps=3, rays=512):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref_py", "nerf", "utils.py")):
        pytest.skip("oracle/_ref_py/nerf/utils.py not staged (python -m oracle.stage_ref_py where /root/reference exists)")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    env["WANDB_MODE"] = "disabled"
    r = subprocess.run([sys.executable, TRAINER_CHILD, "--mode", mode, "--state", state_path, "--out", out_path,
                        "--workspace", workspace, "--steps", str(steps), "--rays", str(rays)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-5000:]
    return torch.load(out_path, map_location="cpu", weights_only=False)


def test_reference_trainer_runs_on_our_operators_and_checkpoints_interoperate(cuda, stage1, tmp_path):
    """the reference's own ``Trainer`` (nerf/utils.py, unmodified) — ``train_one_epoch`` with ``train_step``,
    backward, ``torch.optim.Adam(model.get_params(lr), eps=1e-15)``, LambdaLR per step, EMA(0.95) per epoch — on THIS
    repository's NeRFNetwork / operators, against this repository's own trainer from the same checkpoint, rays and jitter
    stream; then checkpoints in both directions: the reference's ``save_checkpoint(full=True)`` file restores FusedAdam, and
    a file written here is restored by the reference's ``load_checkpoint`` into its torch optimizer / scheduler / EMA."""
    from nerf.network import NeRFNetwork
    from sanerf_b200.checkpoint import load_checkpoint, save_checkpoint
    from sanerf_b200.train import RGBTrainer, default_opt
    _, path, _ = stage1
    steps, n = 3, 512
    ref = _run_trainer_child("train", path, str(tmp_path / "trainer_ref.pt"), str(tmp_path / "ws"), steps, n)
    assert ref["global_step"] == steps

    # (1) this repository's trainer (autograd path: same torch.rand draws in the reference's order), same everything
    twin = NeRFNetwork(default_opt(lambda_distort=0.02))
    load_checkpoint(path, twin)
    twin = twin.cuda()
    tr = RGBTrainer(twin, fused_step=False, ema_decay=0.95)
    g = torch.Generator().manual_seed(77)
    batches = []
    for _ in range(steps):
        o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
        rgb = torch.rand(n, 3, generator=g).cuda()
        batches.append((o, d, rgb))
    torch.manual_seed(1234)
    losses = [float(tr.step(*b)) for b in batches]
    tr.end_epoch()
    assert abs(sum(losses) / steps - ref["avg_loss"]) < 1e-4 * abs(ref["avg_loss"])
    for k, v in twin.state_dict().items():
        if v.dtype.is_floating_point and "aabb" not in k:
            assert _rel(v.cpu(), ref["params"][k]) < 1e-3, k
    for p, s in zip(tr.optimizer.params, ref["ema"]):
        a, _ = tr.optimizer.ranges[id(p)]
        assert _rel(tr.optimizer.ema[a:a + p.numel()].cpu(), s.reshape(-1)) < 1e-3
    assert abs(float(tr.optimizer.dyn[0]) * 0.1 ** (1 / 20000) - ref["lr"]) < 1e-7 * ref["lr"] + 1e-9    # LambdaLR after 3 steps

    # (2) the file the reference's Trainer wrote -> this repository's loader
    fresh = NeRFNetwork(default_opt()).cuda()
    tf = RGBTrainer(fresh, fused_step=False, ema_decay=0.95)
    missing, unexpected = load_checkpoint(ref["ckpt"], fresh, tf)
    assert not missing and not unexpected
    for k, v in fresh.state_dict().items():
        assert torch.equal(v.cpu(), ref["params"][k]), k
    assert int(tf.optimizer.step_count) == steps and tf.global_step == steps and tf.optimizer.ema_updates == 1
    ck = torch.load(ref["ckpt"], map_location="cpu", weights_only=False)
    a, _ = tf.optimizer.ranges[id(fresh.grid.embeddings)]
    assert torch.equal(tf.optimizer.exp_avg[a:a + fresh.grid.embeddings.numel()].cpu(),
                       ck["optimizer"]["state"][0]["exp_avg"].reshape(-1))
    for p, s in zip(tf.optimizer.params, ref["ema"]):
        a, _ = tf.optimizer.ranges[id(p)]
        assert torch.equal(tf.optimizer.ema[a:a + p.numel()].cpu(), s.reshape(-1))

    # (3) a file written here -> the reference's own Trainer.load_checkpoint
    ours = str(tmp_path / "ours_ep0001.pth")
    save_checkpoint(ours, twin, tr, epoch=1, full=True)
    back = _run_trainer_child("load", ours, str(tmp_path / "trainer_back.pt"), str(tmp_path / "ws2"))
    text = "\n".join(back["logs"])
    for needle in ("loaded model", "loaded EMA", "loaded optimizer", "loaded scheduler"):
        assert needle in text and "Failed" not in text and "failed" not in text, text
    assert back["global_step"] == steps and back["epoch"] == 1 and back["ema_updates"] == 1
    for k, v in twin.state_dict().items():
        assert torch.equal(v.cpu(), back["params"][k]), k
    assert back["opt_state_keys"] == list(range(13))                  # 13 trained tensors, numbered through the 5 groups
    a, _ = tr.optimizer.ranges[id(twin.grid.embeddings)]
    assert torch.equal(back["opt_state"][0]["exp_avg"].reshape(-1), tr.optimizer.exp_avg[a:a + twin.grid.embeddings.numel()].cpu())
    assert abs(back["lr"] - float(tr.optimizer.dyn[0])) < 1e-9
