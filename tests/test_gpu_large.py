"""BASELINE configs[4] path: half-precision table with fp32 master weights (sanerf_adam_step_half) and the large-scene step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_adam_step_half_matches_torch_adam(cuda):
    """Three steps of sanerf_adam_step_half == torch.optim.Adam(eps=1e-15) on the fp32 master with the half gradients
    up-cast (main.py:296), and the half table is the rounded master."""
    from sanerf_b200 import _lib
    lib = _lib.load()
    n = 8 * 1000
    g = torch.Generator().manual_seed(0)
    master = torch.randn(n, generator=g).cuda()
    ref = master.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15)
    p16, m, v = master.half(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step, dyn = torch.zeros(1, device="cuda", dtype=torch.int32), torch.tensor([1e-2, 1.0, 1.0, 0.0], device="cuda")
    st = _lib.current_stream(master.device)
    for _ in range(3):
        grad = (torch.randn(n, generator=g) * 0.1).cuda().half()
        ref.grad = grad.float() * 0.5
        opt.step()
        g16 = grad.clone()
        _lib.check(lib.sanerf_adam_schedule(step.data_ptr(), dyn.data_ptr(), 1e-2, 0.9, 0.999, 0.0, None, 0.0, st), "schedule")
        _lib.check(lib.sanerf_adam_step_half(master.data_ptr(), p16.data_ptr(), g16.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                             dyn.data_ptr(), 0.9, 0.999, 1e-15, 0.5, 1, st), "adam_step_half")
        torch.testing.assert_close(master, ref.detach(), rtol=1e-5, atol=1e-6)
        assert torch.equal(p16, master.half()) and float(g16.abs().max()) == 0.0


def test_half_scatter_aggregated_matches_fp32_scatter(cuda):
    """The warp-aggregated fp16 scatter (sums kept in fp32, rounded once per run) against the fp32 scatter of the same
    gradient, ray-ordered samples: level 0's 4096 rows receive hundreds of contributions each."""
    import numpy as np

    from gridencoder import GridEncoder
    from gridencoder.grid import grid_encode
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                      desired_resolution=4096).cuda()
    N, T = 512, 64
    o = torch.rand(N, 1, 3, device="cuda") * 0.6 + 0.2
    step = torch.randn(N, 1, 3, device="cuda") * 0.004
    x = (o + step * torch.arange(T, device="cuda").view(1, T, 1)).clamp(0, 1).reshape(-1, 3).contiguous()
    grad = torch.randn(N * T, 32, device="cuda") * 0.01
    t32 = enc.embeddings.detach().clone().requires_grad_(True)
    grid_encode(x, t32, enc.offsets, enc.per_level_scale, enc.base_resolution, False, 0, False, 0, None).backward(grad)
    t16 = enc.embeddings.detach().half().requires_grad_(True)
    grid_encode(x, t16, enc.offsets, enc.per_level_scale, enc.base_resolution, False, 0, False, 0, None).backward(grad.half())
    exact, got = t32.grad.double(), t16.grad.double()
    rel = ((got - exact).norm() / exact.norm()).item()
    assert rel < 1e-2, rel
    offs = enc.offsets.tolist()
    lvl0 = ((got[:offs[1]] - exact[:offs[1]]).norm() / exact[:offs[1]].norm()).item()
    assert lvl0 < 1e-2, lvl0


def test_large_scene_step_runs_and_updates(cuda):
    from sanerf_b200.large import LargeSceneStep
    step = LargeSceneStep(cuda, n_rays=256, T=32, log2_hashmap_size=16)
    before = step.table16.clone()
    step()
    step()
    torch.cuda.synchronize()
    assert torch.isfinite(step.master).all() and not torch.equal(before, step.table16)
    assert float(step.grad16.abs().max()) == 0.0                        # consumed and cleared by the update
    assert torch.equal(step.table16, step.master.half())
