"""Fused kernels of the step (sampling chain, proposal density, head compositing, losses, Adam) against the
pure-torch oracle of the reference statements they replace."""
import numpy as np
import pytest
import torch

from oracle import render_torch as R
from sanerf_b200 import fused
from sanerf_b200.ops import composite

pytestmark = pytest.mark.gpu


def make_rays(N, seed=0, far_out=True):
    g = torch.Generator().manual_seed(seed)
    o = torch.rand(N, 3, generator=g) - 0.5
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    if far_out:
        o[:5] = o[:5] * 1000          # outside the box: some hit it, some miss (near = far = 1e9)
        d[5] = torch.tensor([0.0, 0.0, 1.0])   # axis-aligned direction: division by 1e-15
    return o, d


def oracle_level(o, d, aabb, min_near, bins, contract, bound, cnf=None):
    near, far = R.near_far_from_aabb(o, d, aabb, min_near)
    if cnf is not None:
        near, far = torch.maximum(near, cnf[:, [0]]), torch.minimum(far, cnf[:, [1]])
    s_near, s_far = R.spacing_fn(near), R.spacing_fn(far)
    real = R.spacing_fn_inv(s_near * (1 - bins) + s_far * bins)
    t = (real[..., 1:] + real[..., :-1]) / 2
    deltas = real[..., 1:] - real[..., :-1]
    xyz = o.unsqueeze(1) + d.unsqueeze(1) * t.unsqueeze(2)
    if contract:
        xyz = R.contract(xyz)
    return t, deltas, (xyz + bound) / (2 * bound)


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.cpu(), b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("T", [128, 64, 33])
@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("contract", [True, False])
def test_sample_uniform(cuda, T, perturb, contract):
    N = 300
    o, d = make_rays(N, T)
    bound = 2.0 if contract else 128.0
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3)
    noise = torch.rand(N, T + 1, generator=torch.Generator().manual_seed(9)) if perturb else None
    cnf = torch.tensor([[0.3, 50.0]]).expand(N, 2).contiguous() if T == 64 else None
    bins, t_mid, deltas, x01 = fused.sample_uniform(o.cuda(), d.cuda(), aabb.cuda(), 0.2, T,
                                                    None if noise is None else noise.cuda(),
                                                    None if cnf is None else cnf.cuda(), contract, bound)
    eb = torch.linspace(0, 1, T + 1).unsqueeze(0).expand(N, -1)
    if perturb:
        eb = (eb + (noise - 0.5) / T).clamp(0, 1)
    et, ed, ex = oracle_level(o, d, aabb, 0.2, eb, contract, bound, cnf)
    close(bins, eb, atol=1e-7)
    hit = (et[:, 0] < 1e8)
    assert (~hit).any() and hit.sum() > N - 6
    close(t_mid[hit], et[hit], rtol=1e-5, atol=1e-5)
    close(deltas[hit], ed[hit], rtol=1e-4, atol=1e-6)
    close(x01[hit], ex[hit], rtol=1e-5, atol=2e-6)
    # rays that miss the box: near = far = 1e9 -> s = 1 -> 1/(2-2s) = inf edges, inf - inf = NaN intervals,
    # exactly like the reference statements (the NaN weights are zeroed by nan_to_num in the compositing)
    torch.testing.assert_close(t_mid[~hit].cpu(), et[~hit], rtol=1e-6, atol=0, equal_nan=True)
    torch.testing.assert_close(deltas[~hit].cpu(), ed[~hit], rtol=1e-4, atol=0, equal_nan=True)


@pytest.mark.parametrize("T0,T", [(128, 64), (64, 32), (32, 16), (40, 100)])
@pytest.mark.parametrize("perturb", [False, True])
def test_sample_pdf(cuda, T0, T, perturb):
    N = 257
    o, d = make_rays(N, T0 + T, far_out=False)
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3)
    g = torch.Generator().manual_seed(T)
    pbins = torch.sort(torch.rand(N, T0 + 1, generator=g), -1).values
    pbins[:, 0], pbins[:, -1] = 0.0, 1.0
    w = torch.rand(N, T0, generator=g) ** 4
    w[:3] = 0                                             # all-zero weights: uniform pdf through the +0.01
    w[3, 5] = 50.0                                        # one dominant interval: cdf clamps at 1
    noise = torch.rand(N, T + 1, generator=g) if perturb else None
    bins, t_mid, deltas, x01 = fused.sample_pdf(o.cuda(), d.cuda(), aabb.cuda(), 0.2, pbins.cuda(), w.cuda(), T,
                                                None if noise is None else noise.cuda(), None, True, 2.0)
    eb = R.sample_pdf(pbins, w, T + 1, perturb, noise)
    # ulp-level cdf differences (parallel vs sequential prefix sums) are divided by the pdf of the interval:
    # a few edges in near-empty intervals move by ~1e-5 of the [0,1] range, the bulk agrees to 1e-6
    close(bins, eb, rtol=0, atol=3e-5)
    assert (bins.cpu() - eb).abs().mean().item() < 5e-7
    et, ed, ex = oracle_level(o, d, aabb, 0.2, bins.cpu(), True, 2.0)   # downstream math from OUR bins
    close(t_mid, et, rtol=1e-5, atol=1e-5)
    close(deltas, ed, rtol=1e-4, atol=1e-6)
    close(x01, ex, rtol=1e-5, atol=2e-6)
    assert torch.all(bins[:, 1:] >= bins[:, :-1])        # sortedness of the resampled edges


def test_sample_pdf_golden_from_reference(cuda, ref_cpu):
    """Bins the REAL reference sample_pdf produced (oracle/make_golden.py --cpu)."""
    pbins, w = torch.from_numpy(ref_cpu["pdf_bins"]), torch.from_numpy(ref_cpu["pdf_w"])
    N = pbins.shape[0]
    o, d = make_rays(N, 1, far_out=False)
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3)
    bins, _, _, _ = fused.sample_pdf(o.cuda(), d.cuda(), aabb.cuda(), 0.2, pbins.cuda(), w.cuda(), 16, None, None,
                                     True, 2.0)
    close(bins, torch.from_numpy(ref_cpu["pdf_out"]), rtol=0, atol=3e-5)
    assert (bins.cpu() - torch.from_numpy(ref_cpu["pdf_out"])).abs().mean().item() < 5e-7


def test_prop_density_forward_backward(cuda):
    from gridencoder import GridEncoder
    from nerf.network import MLP
    from activation import trunc_exp
    torch.manual_seed(0)
    for finest in (128, 256):
        enc = GridEncoder(input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17, desired_resolution=finest).cuda()
        mlp = MLP(10, 1, 16, 2, bias=False).cuda()
        with torch.no_grad():
            enc.embeddings.uniform_(-1, 1)
        assert fused.prop_density_supported(enc, mlp)
        x01 = torch.rand(4000, 3, device="cuda")
        x01[0] = 1.5                                     # out of range -> zero encoding
        sig = fused.prop_density(x01, enc, mlp)
        ref = trunc_exp(mlp(enc(x01 * 2 - 1, bound=1)).squeeze(-1))
        torch.testing.assert_close(sig, ref, rtol=1e-4, atol=1e-6)
        g = torch.randn(4000, device="cuda")
        (sig * g).sum().backward()
        got = [p.grad.clone() for p in (enc.embeddings, mlp.net[0].weight, mlp.net[1].weight)]
        for p in (enc.embeddings, mlp.net[0].weight, mlp.net[1].weight):
            p.grad = None
        (ref * g).sum().backward()
        for a, p in zip(got, (enc.embeddings, mlp.net[0].weight, mlp.net[1].weight)):
            torch.testing.assert_close(a, p.grad, rtol=1e-3, atol=1e-4 * p.grad.abs().max().item())
    # CPU oracle of the same chain
    ref_enc = R.GridEncoderRef(input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17, desired_resolution=256)
    ref_enc.load_state_dict({k: v.cpu() for k, v in enc.state_dict().items()})
    ref_mlp = R.MLP(10, 1, 16, 2, bias=False)
    ref_mlp.load_state_dict({k: v.cpu() for k, v in mlp.state_dict().items()})
    exp = R.trunc_exp(ref_mlp(ref_enc(x01.cpu() * 2 - 1, bound=1)).squeeze(-1))
    torch.testing.assert_close(sig.detach().cpu(), exp.detach(), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("T,t_thresh", [(32, 0.0), (70, 0.0), (32, 1e-2), (9, 0.0)])
def test_head_composite_equals_separate_ops(cuda, T, t_thresh):
    """W = 16 runs the fused lane-per-sample kernels (csrc/head_composite.cu); trunc_exp + composite are the checker."""
    from activation import trunc_exp
    N = 200
    g = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(N, T, 16, device="cuda", generator=g).requires_grad_(True)
    bins = torch.sort(torch.rand(N, T + 1, device="cuda", generator=g), -1).values * 5 + 0.2
    deltas, ts = bins[:, 1:] - bins[:, :-1], (bins[:, 1:] + bins[:, :-1]) / 2
    sig, w, ws, dp, out, alive = fused.head_composite(f, deltas, ts, True, t_thresh)
    f2 = f.detach().clone().requires_grad_(True)
    sig2 = trunc_exp(f2[..., 0])
    w2, ws2, dp2, out2, alive2 = composite(sig2, deltas, ts, f2[..., 1:].contiguous(), t_thresh=t_thresh)
    assert torch.equal(alive, alive2)
    for a, b in ((sig, sig2), (w, w2), (ws, ws2), (dp, dp2), (out, out2)):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    go, gw = torch.randn(N, 15, device="cuda", generator=g), torch.randn(N, T, device="cuda", generator=g)
    ((out * go).sum() + (w * gw).sum() + dp.sum() + (sig * 0.1).sum()).backward()
    ((out2 * go).sum() + (w2 * gw).sum() + dp2.sum() + (sig2 * 0.1).sum()).backward()
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-4, atol=1e-5)   # suffix = total - prefix vs a reverse scan


def test_losses_vs_oracle(cuda):
    N = 300
    g = torch.Generator().manual_seed(4)
    mk_bins = lambda T: torch.sort(torch.rand(N, T + 1, generator=g), -1).values  # noqa: E731
    mk_w = lambda T: torch.softmax(torch.randn(N, T, generator=g) * 2, -1)        # noqa: E731
    b0, b1, b2 = mk_bins(128), mk_bins(64), mk_bins(32)
    b0[:, 0] = b1[:, 0] = b2[:, 0] = 0
    b0[:, -1] = b1[:, -1] = b2[:, -1] = 1
    w0, w1, w2 = (mk_w(T).requires_grad_(True) for T in (128, 64, 32))
    exp_p = R.proposal_loss([b0, b1, b2], [w0, w1, w2])
    exp_d = R.distort_loss(b2, w2)
    (exp_p * 1.7 + exp_d * 0.3).backward()
    c = lambda t: t.detach().cuda().requires_grad_(t.requires_grad)  # noqa: E731
    cw0, cw1, cw2 = c(w0), c(w1), c(w2)
    got_p = fused.proposal_loss([b0.cuda(), b1.cuda(), b2.cuda()], [cw0, cw1, cw2])
    got_d = fused.distort_loss(b2.cuda(), cw2)
    (got_p * 1.7 + got_d * 0.3).backward()
    torch.testing.assert_close(got_p.cpu(), exp_p.detach(), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(got_d.cpu(), exp_d.detach(), rtol=1e-4, atol=1e-8)
    for a, b in ((cw0, w0), (cw1, w1), (cw2, w2)):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-3, atol=1e-5 * b.grad.abs().max().item())


def test_fused_adam_matches_torch_adam(cuda):
    torch.manual_seed(0)
    shapes = [(1001, 2), (16, 10), (7,), (64, 64)]
    ps = [torch.randn(*s, device="cuda").requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    ref = torch.optim.Adam(qs, lr=1e-2, eps=1e-15)
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lambda it: 0.1 ** min(it / 50, 1))
    opt = fused.FusedAdam(ps, lr=1e-2, eps=1e-15, decay_iters=50)
    for it in range(80):
        grads = [torch.randn(*s, device="cuda") * (0.0 if (it == 3 and i == 1) else 1.0) for i, s in enumerate(shapes)]
        for p, q, gr in zip(ps, qs, grads):
            p.grad.copy_(gr)
            q.grad = gr.clone()
        opt.step(grad_scale=1.0, zero_grad=True)
        ref.step(); sched.step()
        assert all(float(p.grad.abs().max()) == 0 for p in ps)
    for p, q in zip(ps, qs):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-4, atol=1e-5)
    assert int(opt.step_count.item()) == 80


def test_fused_adam_large_pass_matches_torch_adam(cuda):
    """A large pass (whole table sized, with a ragged tail): torch.optim.Adam semantics over several steps of a schedule,
    gradients cleared; a range whose gate is clear is left untouched."""
    torch.manual_seed(1)
    n = (1 << 22) + 768 * 5 + 333                       # whole chunks + a ragged tail
    p = torch.randn(n, device="cuda").requires_grad_(True)
    q = p.detach().clone().requires_grad_(True)
    ref = torch.optim.Adam([q], lr=1e-2, eps=1e-15)
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lambda it: 0.1 ** min(it / 50, 1))
    opt = fused.FusedAdam([p], lr=1e-2, eps=1e-15, decay_iters=50)
    for it in range(6):
        gr = torch.randn(n, device="cuda") * (0.0 if it == 2 else 1.0)
        p.grad.copy_(gr)
        q.grad = gr.clone()
        opt.step(grad_scale=1.0, zero_grad=True)
        ref.step(); sched.step()
        assert float(p.grad.abs().max()) == 0
    torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-4, atol=1e-5)
    # a range whose gate is clear is not touched (deferred update after flush())
    before = p.detach().clone()
    p.grad.fill_(1.0)
    opt.clear_gate()
    opt.apply(0, n, gated=True)
    torch.cuda.synchronize()
    assert torch.equal(p.detach(), before) and float(p.grad.min()) == 1.0


def test_fused_renderer_equals_generic_renderer(cuda):
    """Same parameters, same random draws: the fused fast path and the op-by-op path agree."""
    from tests.test_gpu_render import build_pair, rays
    _, model = build_pair(seed=11)
    model.train()
    o, d = rays(200, seed=12)
    o, d = o.cuda(), d.cuda()
    outs = []
    for flag in (True, False):
        model.fused = flag
        model.zero_grad(set_to_none=True)
        torch.manual_seed(123)
        out = model.render(o, d, staged=False, perturb=True, update_proposal=True)
        loss = out["image"].square().mean() + out["proposal_loss"] + 0.02 * out["distort_loss"]
        loss.backward()
        outs.append((out, {n: p.grad.clone() for n, p in model.named_parameters()}))
    (a, ga), (b, gb) = outs
    for k in ("image", "depth", "weights_sum", "weights", "proposal_loss", "distort_loss"):
        torch.testing.assert_close(a[k], b[k], rtol=1e-3, atol=1e-4, msg=lambda m, k=k: f"{k}: {m}")
    from tests.test_gpu_render import assert_grad_close
    for n in ga:
        assert_grad_close(n, ga[n], gb[n].cpu())


@pytest.mark.parametrize("N", [8192, 77])
def test_view_head_kernel_matches_autograd(cuda, N):
    """Fused deferred-shading head (SH + view MLP + sigmoid + background + MSE, forward and backward) vs torch."""
    from shencoder import SHEncoder
    from sanerf_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    geo = torch.randn(N, 15, device="cuda", generator=g).requires_grad_(True)
    ws = torch.rand(N, device="cuda", generator=g).requires_grad_(True)
    d = torch.randn(N, 3, device="cuda", generator=g) * 3
    gt = torch.rand(N, 3, device="cuda", generator=g)
    w1 = (torch.randn(32, 31, device="cuda", generator=g) / 31 ** 0.5).requires_grad_(True)
    w2 = (torch.randn(32, 32, device="cuda", generator=g) / 32 ** 0.5).requires_grad_(True)
    w3 = (torch.randn(3, 32, device="cuda", generator=g) / 32 ** 0.5).requires_grad_(True)
    sh = SHEncoder(degree=4)(d)
    f = torch.cat([geo, ws.unsqueeze(-1) * sh], dim=-1)
    rgb = torch.sigmoid(torch.relu(torch.relu(f @ w1.t()) @ w2.t()) @ w3.t())
    image_ref = rgb + (1 - ws).unsqueeze(-1) * 1.0
    loss_ref = 0.7 * torch.nn.functional.mse_loss(image_ref, gt)
    loss_ref.backward()

    lib = _lib.load()
    image, loss = torch.empty(N, 3, device="cuda"), torch.zeros(1, device="cuda")
    g_geo, g_ws = torch.empty(N, 15, device="cuda"), torch.empty(N, device="cuda")
    gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
    rc = lib.sanerf_view_head(geo.data_ptr(), ws.data_ptr(), d.data_ptr(), gt.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                              w3.data_ptr(), 1.0, 0.7, N, image.data_ptr(), loss.data_ptr(), g_geo.data_ptr(), g_ws.data_ptr(),
                              gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), _lib.current_stream(geo.device))
    _lib.check(rc, "view_head")
    torch.testing.assert_close(image, image_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss[0], loss_ref.detach(), rtol=1e-5, atol=1e-8)
    scale = lambda t: t.abs().max().item()  # noqa: E731
    torch.testing.assert_close(g_geo, geo.grad, rtol=1e-4, atol=1e-5 * scale(geo.grad))
    torch.testing.assert_close(g_ws, ws.grad, rtol=1e-4, atol=1e-5 * scale(ws.grad))
    for got, ref in zip(gw, (w1, w2, w3)):
        torch.testing.assert_close(got, ref.grad, rtol=1e-4, atol=2e-5 * scale(ref.grad))
    # forward only (inference): same image, nothing else touched
    image2 = torch.empty(N, 3, device="cuda")
    rc = lib.sanerf_view_head(geo.data_ptr(), ws.data_ptr(), d.data_ptr(), None, w1.data_ptr(), w2.data_ptr(), w3.data_ptr(),
                              1.0, 1.0, N, image2.data_ptr(), None, None, None, None, None, None,
                              _lib.current_stream(geo.device))
    _lib.check(rc, "view_head")
    torch.testing.assert_close(image2, image_ref.detach(), rtol=1e-5, atol=1e-6)


def test_sample_pdf_with_fused_compositing_equals_two_kernels(cuda):
    """sanerf_sample_pdf fed with (sigma, deltas) of the previous level == composite (C = 0) followed by sample_pdf."""
    from sanerf_b200 import _lib
    N, T0, T = 257, 128, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    o = torch.rand(N, 3, device="cuda", generator=g) - 0.5
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=g), dim=-1)
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3, device="cuda")
    noise0, noise1 = torch.rand(N, T0 + 1, device="cuda", generator=g), torch.rand(N, T + 1, device="cuda", generator=g)
    bins0, t0, d0, _ = fused.sample_uniform(o, d, aabb, 0.2, T0, noise0)
    sigma0 = torch.randn(N, T0, device="cuda", generator=g).exp()
    w0 = composite(sigma0, d0, t0, None, last_sample_opaque=True)[0]
    ref = fused.sample_pdf(o, d, aabb, 0.2, bins0, w0, T, noise1)
    lib = _lib.load()
    outs = (torch.empty(N, T + 1, device="cuda"), torch.empty(N, T, device="cuda"), torch.empty(N, T, device="cuda"),
            torch.empty(N, T, 3, device="cuda"))
    w_out = torch.full((N, T0), float("nan"), device="cuda")
    rc = lib.sanerf_sample_pdf(o.data_ptr(), d.data_ptr(), aabb.data_ptr(), 0.2, None, 0, bins0.data_ptr(), None, T0,
                               noise1.data_ptr(), N, T, 1, 2.0, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                               outs[3].data_ptr(), sigma0.data_ptr(), d0.data_ptr(), 1, w_out.data_ptr(),
                               _lib.current_stream(o.device))
    _lib.check(rc, "sample_pdf")
    assert torch.equal(w_out, w0)                          # same arithmetic as the compositing kernel
    for a, b in zip(outs, ref):
        assert torch.equal(a, b)


def test_uniform_fill_is_uniform_and_advances(cuda):
    """sanerf_uniform_fill: [0,1) values, flat histogram, fresh numbers on every launch (also when replayed in a CUDA graph),
    optional accumulator clear."""
    from sanerf_b200 import _lib
    lib = _lib.load()
    n = 1 << 20
    out = torch.empty(n + 3, device="cuda")[:n + 1]           # odd length: exercises the tail
    state = torch.zeros(2, device="cuda", dtype=torch.int32)
    acc = torch.ones(1, device="cuda")
    st = _lib.current_stream(out.device)
    _lib.check(lib.sanerf_uniform_fill(out.data_ptr(), n + 1, 1234, state.data_ptr(), acc.data_ptr(), 1, st), "uniform_fill")
    a = out.clone()
    assert float(acc) == 0.0 and float(a.min()) >= 0.0 and float(a.max()) < 1.0
    hist = torch.histc(a, bins=64, min=0, max=1) / a.numel()
    assert (hist - 1 / 64).abs().max().item() < 2e-3 and abs(float(a.mean()) - 0.5) < 2e-3
    assert abs(float((a[:-1] * a[1:]).mean()) - 0.25) < 2e-3                      # no lag-1 correlation
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            _lib.check(lib.sanerf_uniform_fill(out.data_ptr(), n + 1, 1234, state.data_ptr(), None, 0,
                                               _lib.current_stream(out.device)), "uniform_fill")
    g.replay(); b = out.clone(); g.replay(); c = out.clone()
    torch.cuda.synchronize()
    assert int(state[0]) == 3 and int(state[1]) == 0
    assert not torch.equal(a, b) and not torch.equal(b, c)
