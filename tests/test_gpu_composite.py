"""Per-ray compositing kernel (forward + backward) vs the torch restatement of
nerf/renderer.py:309-338 (oracle/render_torch.py) — dense, packed/ragged, early termination,
edge cases.  fp32 bar: 1e-3 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import render_torch as R
from sanerf_b200.ops import composite

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def make_rays(N, T, C, seed=0, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    sig = torch.exp(torch.randn(N, T, generator=g) * 1.5)
    bins = torch.sort(torch.rand(N, T + 1, generator=g), dim=-1).values * 6 + 0.2
    deltas = bins[:, 1:] - bins[:, :-1]
    ts = (bins[:, 1:] + bins[:, :-1]) / 2
    feats = torch.randn(N, T, C, generator=g) if C else None
    mv = lambda t: None if t is None else t.to(dev)  # noqa: E731
    return mv(sig), mv(deltas), mv(ts), mv(feats)


@pytest.mark.parametrize("T", [2, 7, 32, 64, 128, 200])  # the reference's own statements need T >= 2
@pytest.mark.parametrize("C", [0, 3, 8, 15, 31, 128, 159, 256])
@pytest.mark.parametrize("opaque", [True, False])
def test_forward_backward_dense(cuda, T, C, opaque):
    if C > 31 and T not in (32, 128):
        pytest.skip("large-C cases only at the reference's sample counts")
    N = 257
    sig, deltas, ts, feats = make_rays(N, T, C, seed=T * 1000 + C)
    sig.requires_grad_(True)
    if C:
        feats.requires_grad_(True)
    w, ws, dp, out, alive = composite(sig, deltas, ts, feats, last_sample_opaque=opaque)
    sig_r = sig.detach().clone().requires_grad_(True)
    feats_r = feats.detach().clone().requires_grad_(True) if C else None
    w_r, ws_r, dp_r, out_r, alive_r = R.composite(sig_r, deltas, ts, feats_r, last_sample_opaque=opaque)
    torch.testing.assert_close(w, w_r, rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(ws, ws_r, rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(dp, dp_r, rtol=RTOL, atol=1e-4)
    assert torch.equal(alive, alive_r) and int(alive.min()) == T
    gen = torch.Generator(device="cuda").manual_seed(5)
    gw = torch.randn(N, T, device="cuda", generator=gen)
    gws, gdp = torch.randn(N, device="cuda", generator=gen), torch.randn(N, device="cuda", generator=gen)
    loss = (w * gw).sum() + (ws * gws).sum() + (dp * gdp).sum()
    loss_r = (w_r * gw).sum() + (ws_r * gws).sum() + (dp_r * gdp).sum()
    if C:
        torch.testing.assert_close(out, out_r, rtol=RTOL, atol=1e-4)
        go = torch.randn(N, C, device="cuda", generator=gen)
        loss = loss + (out * go).sum()
        loss_r = loss_r + (out_r * go).sum()
    loss.backward()
    loss_r.backward()
    scale = sig_r.grad.abs().max().item()
    torch.testing.assert_close(sig.grad, sig_r.grad, rtol=RTOL, atol=1e-4 * scale)
    if C:
        torch.testing.assert_close(feats.grad, feats_r.grad, rtol=RTOL, atol=1e-6)
    if opaque:
        assert torch.all(sig.grad[:, -1] == 0)          # the opaque sample's sigma is replaced by a constant
        torch.testing.assert_close(ws, torch.ones_like(ws), rtol=0, atol=1e-5)


def test_packed_ragged_matches_dense_per_ray(cuda):
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(0, 97, (300,), generator=g)
    counts[:4] = torch.tensor([0, 1, 32, 33])        # empty ray, single sample, chunk boundaries
    offs = torch.zeros(301, dtype=torch.int32)
    offs[1:] = torch.cumsum(counts, 0).int()
    M, C = int(offs[-1]), 31
    sig = torch.exp(torch.randn(M, generator=g)).cuda().requires_grad_(True)
    deltas = (torch.rand(M, generator=g) * 0.2 + 0.01).cuda()
    ts = torch.rand(M, generator=g).cuda()
    feats = torch.randn(M, C, generator=g).cuda().requires_grad_(True)
    w, ws, dp, out, alive = composite(sig, deltas, ts, feats, ray_offsets=offs.cuda(), max_count=96)
    go = torch.randn(300, C, device="cuda")
    ((out * go).sum() + ws.sum() * 0.3 + (dp * dp).sum()).backward()
    sig_r = sig.detach().clone().requires_grad_(True)
    feats_r = feats.detach().clone().requires_grad_(True)
    tot = 0
    for r in range(300):
        a, b = int(offs[r]), int(offs[r + 1])
        if a == b:
            assert ws[r] == 0 and dp[r] == 0 and torch.all(out[r] == 0) and alive[r] == 0
            continue
        if b - a == 1:  # a single (opaque) sample: alpha = 1, T = 1 (the torch statements cannot express T = 1)
            wr, wsr = torch.ones(1, 1, device="cuda"), torch.ones(1, device="cuda")
            dpr, outr = ts[None, a:b].sum(-1), feats_r[None, a:b].sum(1)
        else:
            wr, wsr, dpr, outr, _ = R.composite(sig_r[None, a:b], deltas[None, a:b], ts[None, a:b],
                                                feats_r[None, a:b])
        torch.testing.assert_close(w[a:b], wr[0], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(out[r], outr[0], rtol=RTOL, atol=1e-4)
        torch.testing.assert_close(dp[r], dpr[0], rtol=RTOL, atol=1e-5)
        tot = tot + (outr[0] * go[r]).sum() + wsr.sum() * 0.3 + (dpr * dpr).sum()
    tot.backward()
    torch.testing.assert_close(sig.grad, sig_r.grad, rtol=RTOL, atol=1e-4 * sig_r.grad.abs().max().item())
    torch.testing.assert_close(feats.grad, feats_r.grad, rtol=RTOL, atol=1e-6)


def test_long_rays_use_the_sequential_fallback(cuda):
    sig, deltas, ts, feats = make_rays(40, 300, 4, seed=8)
    deltas = deltas * 0.05
    sig.requires_grad_(True)
    w, ws, dp, out, _ = composite(sig, deltas, ts, feats)
    (out.sum() + dp.sum()).backward()
    sig_r = sig.detach().clone().requires_grad_(True)
    w_r, ws_r, dp_r, out_r, _ = R.composite(sig_r, deltas, ts, feats)
    (out_r.sum() + dp_r.sum()).backward()
    torch.testing.assert_close(w, w_r, rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(sig.grad, sig_r.grad, rtol=2e-3, atol=2e-4 * sig_r.grad.abs().max().item())


@pytest.mark.parametrize("t_thresh", [1e-4, 1e-2, 0.3])
def test_early_termination_counts_and_truncation_bound(cuda, t_thresh):
    """n_alive == #{k : T_k >= t_thresh} with T_k accumulated sequentially in fp32 (ties reported
    separately); outputs differ from the un-terminated ones by at most T_k * max|c|."""
    N, T, C = 2000, 128, 31
    sig, deltas, ts, feats = make_rays(N, T, C, seed=21)
    w, ws, dp, out, alive = composite(sig, deltas, ts, feats, t_thresh=t_thresh)
    counts, ties = R.n_alive_sequential(sig, deltas, t_thresh)
    got = alive.cpu().numpy()
    assert np.array_equal(got[~ties], counts[~ties])
    assert np.abs(got[ties] - counts[ties]).max(initial=0) <= 1 and ties.mean() < 0.01
    w_r, ws_r, dp_r, out_r, alive_r = R.composite(sig, deltas, ts, feats, t_thresh=t_thresh)
    mism = (alive != alive_r).float().mean().item()
    assert mism < 0.01                                   # parallel-scan vs cumsum ties only
    same = alive == alive_r
    torch.testing.assert_close(out[same], out_r[same], rtol=RTOL, atol=1e-4)
    w0, ws0, dp0, out0, _ = composite(sig, deltas, ts, feats, t_thresh=0.0)
    assert (out - out0).abs().max().item() <= t_thresh * feats.abs().max().item() * 1.01 + 1e-5
    assert torch.all(w[torch.arange(T, device="cuda")[None, :] >= alive[:, None]] == 0)
    assert int(alive.max()) <= T and int(alive.min()) >= 1


def test_edge_cases(cuda):
    # rays that miss the box: near = far = 1e9 -> zero-length intervals; weights = [0,..,0,1] (last sample opaque)
    N, T = 8, 32
    sig = torch.rand(N, T, device="cuda") * 5
    deltas = torch.zeros(N, T, device="cuda")
    ts = torch.full((N, T), 1e9, device="cuda")
    w, ws, dp, out, alive = composite(sig, deltas, ts, torch.ones(N, T, 3, device="cuda"))
    assert torch.all(w[:, :-1] == 0) and torch.all(w[:, -1] == 1) and torch.all(ws == 1)
    torch.testing.assert_close(out, torch.ones(N, 3, device="cuda"))
    # 0 * inf = NaN -> nan_to_num(0) (renderer.py:326): sigma = inf with delta = 0 mid-ray
    sig2 = torch.ones(N, T, device="cuda"); sig2[:, 5] = float("inf")
    d2 = torch.full((N, T), 0.01, device="cuda"); d2[:, 5] = 0.0
    w2, ws2, _, _, _ = composite(sig2, d2, ts, None, last_sample_opaque=False)
    w2_r, _ = R.sigma_to_weights(sig2, d2, False)
    assert torch.isfinite(w2).all()
    torch.testing.assert_close(w2, w2_r, rtol=RTOL, atol=1e-7, equal_nan=True)
    # N = 0
    e = torch.empty(0, 16, device="cuda")
    w0, ws0, dp0, out0, al0 = composite(e, e, e, torch.empty(0, 16, 4, device="cuda"))
    assert w0.shape == (0, 16) and out0.shape == (0, 4)
    with pytest.raises(RuntimeError, match="at most 256 channels"):
        composite(sig, deltas, ts, torch.ones(N, T, 300, device="cuda"))


def test_golden_run_weights(cuda, ref_cpu):
    """The weights the REAL reference renderer produced for its final level (stub field, perturb=False): feed the
    same sigma/deltas through the kernel.  Pins sigma->weights against renderer.py itself."""
    from oracle.make_golden import stub_sigma
    w_ref = torch.from_numpy(ref_cpu["run_weights"])          # [96, 32]
    assert w_ref.shape[1] == 32
    # recompute the reference's final-level inputs with the oracle sampler (validated on CPU against the same golden)
    from tests.test_oracle_render import run_stub_oracle
    res = run_stub_oracle(ref_cpu, want_internals=True)
    sig, deltas, ts = (res[k].cuda() for k in ("sigmas", "deltas", "rays_t"))
    w, ws, dp, _, _ = composite(sig, deltas, ts, None)
    torch.testing.assert_close(w.cpu(), w_ref, rtol=RTOL, atol=2e-6)
    torch.testing.assert_close(dp.cpu(), torch.from_numpy(ref_cpu["run_depth"]), rtol=RTOL, atol=1e-4)
