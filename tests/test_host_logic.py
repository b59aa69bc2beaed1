"""Host-side logic that needs no GPU: slice arithmetic of the fused multi-GPU update, flat-buffer ranges, the reference's
checkpoint layout and warm start, the oracle's torch_ema restatement, the Philox reference of the jitter kernel."""
import types

import pytest
import torch

from sanerf_b200 import symm


@pytest.mark.parametrize("start,stop,world", [(0, 12599936, 8), (12599936, 14236288, 8), (0, 7168, 8), (64, 64 + 4 * 13, 4),
                                              (0, 40, 2), (0, 12, 8)])
def test_slice_bounds_partition_the_range(start, stop, world):
    """Every element of [start, stop) belongs to exactly one rank's slice; slices are float4-aligned (symm_adam_kernel)."""
    spans = [symm.slice_bounds(start, stop, world, r) for r in range(world)]
    covered = 0
    for (lo, hi) in spans:
        assert lo % 4 == 0 and hi % 4 == 0 and start <= lo <= hi <= stop
        covered += hi - lo
    assert covered == stop - start
    starts = sorted(lo for lo, hi in spans if hi > lo)
    ends = sorted(hi for lo, hi in spans if hi > lo)
    assert starts[0] == start and ends[-1] == stop and starts[1:] == ends[:-1]


def _cpu_model(with_sam=False):
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import default_opt
    torch.manual_seed(0)
    return NeRFNetwork(default_opt(with_sam=with_sam))


def test_flat_ranges_and_reference_param_groups():
    """FusedAdam slots (multiples of 32, proposal networks at the tail) and the torch.optim.Adam(model.get_params(lr))
    numbering the checkpoint's optimizer entry uses (network.py:278-308)."""
    from sanerf_b200.checkpoint import _param_groups
    from sanerf_b200.fused import FusedAdam
    model = _cpu_model()
    opt = FusedAdam(list(model.parameters()))
    for p in model.parameters():
        a, b = opt.ranges[id(p)]
        assert a % 32 == 0 and (b - a) % 32 == 0 and b - a >= p.numel()
        assert p.data_ptr() == opt.flat_param[a:].data_ptr() and p.grad.data_ptr() == opt.flat_grad[a:].data_ptr()
    lo, hi = opt.range_of([*model.prop_encoders.parameters(), *model.prop_mlp.parameters()])
    assert hi == opt.flat_param.numel() and lo > opt.ranges[id(model.grid.embeddings)][1]
    with pytest.raises(ValueError):
        opt.range_of([model.grid.embeddings, *model.prop_mlp.parameters()])          # not contiguous
    groups, order = _param_groups(model, opt)
    assert [len(g) for g in groups] == [1, 3, 3, 2, 4] and sum(groups, []) == list(range(13))
    assert order[0] is model.grid.embeddings and order[-1] is model.prop_mlp[1].net[1].weight


def test_checkpoint_layout_and_warm_start_on_cpu(tmp_path):
    from sanerf_b200.checkpoint import checkpoint_state, load_checkpoint, save_checkpoint, warm_start
    stage1 = _cpu_model()
    state = checkpoint_state(stage1, epoch=7)
    assert set(state) == {"epoch", "global_step", "stats", "model"} and state["epoch"] == 7
    assert set(state["stats"]) == {"loss", "valid_loss", "results", "checkpoints", "best_result"}     # nerf/utils.py:612-618
    path = str(tmp_path / "ngp_ep0007.pth")
    save_checkpoint(path, stage1, epoch=7)
    sam = _cpu_model(with_sam=True)
    frozen = warm_start(sam, path)                                                                   # main.py:255-262
    assert {k.split(".")[0] for k in frozen} == {"grid", "grid_mlp", "view_mlp", "prop_encoders", "prop_mlp"}
    assert torch.equal(sam.grid.embeddings, stage1.grid.embeddings) and not sam.grid.embeddings.requires_grad
    assert sam.s_grid.embeddings.requires_grad and sam.samvit_mlp[0].net[0].weight.requires_grad
    other = _cpu_model()
    with torch.no_grad():
        other.grid.embeddings.add_(1.0)
    missing, unexpected = load_checkpoint(path, other)
    assert not missing and not unexpected and torch.equal(other.grid.embeddings, stage1.grid.embeddings)
    bare = str(tmp_path / "bare.pth")
    torch.save(stage1.state_dict(), bare)                                                            # nerf/utils.py:2118-2121
    assert load_checkpoint(bare, other) == ([], [])


def test_ema_restatement_matches_published_formula():
    from oracle.ema_ref import ExponentialMovingAverage
    p = torch.nn.Parameter(torch.tensor([1.0, 2.0]))
    ema = ExponentialMovingAverage([p], decay=0.95)
    shadow = p.detach().clone()
    for k in range(1, 30):
        with torch.no_grad():
            p.add_(0.5)
        ema.update()
        decay = min(0.95, (1 + k) / (10 + k))
        shadow = shadow - (1 - decay) * (shadow - p.detach())
        torch.testing.assert_close(ema.shadow_params[0], shadow)
    live = p.detach().clone()
    ema.store(); ema.copy_to()
    torch.testing.assert_close(p.detach(), shadow)
    ema.restore()
    assert torch.equal(p.detach(), live)
    st = ema.state_dict()
    other = ExponentialMovingAverage([p], decay=0.5)
    other.load_state_dict(st)
    assert other.decay == 0.95 and other.num_updates == 29 and torch.equal(other.shadow_params[0], ema.shadow_params[0])


def test_multi_gpu_trainer_requires_equal_shards():
    """_check_equal_shards is a no-op on one rank (the multi-rank branch is exercised by tests/test_gpu_multi.py)."""
    from sanerf_b200 import train
    t = types.SimpleNamespace(world_size=1, _checked_counts=set())
    train._check_equal_shards(t, 123)
    assert not t._checked_counts
