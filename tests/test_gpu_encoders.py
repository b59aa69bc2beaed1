"""SH / frequency encoders and trunc_exp on sm_100a vs the oracle, the reference extensions
(oracle/_ref) and the golden vectors.  fp32 bar: 1e-3 relative (north_star)."""
import numpy as np
import pytest
import torch

import _freqencoder
import _shencoder
from activation import trunc_exp
from freqencoder import FreqEncoder
from oracle import encoders_np
from shencoder import SHEncoder

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def unit_dirs(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    d[:6] = torch.tensor([[1.0, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]])  # poles / axes
    return d


@pytest.mark.parametrize("degree", range(1, 9))
def test_sh_forward_and_jacobian_vs_oracle(cuda, degree):
    d = unit_dirs(2000, degree)
    out = torch.empty(2000, degree ** 2, device="cuda")
    jac = torch.empty(2000, 3 * degree ** 2, device="cuda")
    _shencoder.sh_encode_forward(d.cuda(), out, 2000, 3, degree, jac)
    exp, exp_j = encoders_np.sh_encode(d.numpy(), degree, want_jacobian=True)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(jac.cpu().numpy().reshape(2000, 3, -1), exp_j, rtol=RTOL, atol=5e-4)


@pytest.mark.parametrize("degree", [1, 4, 8])
def test_sh_vs_reference_extension(cuda, ref_ext, degree):
    ref = ref_ext("shencoder")
    d = unit_dirs(3000, 40 + degree).cuda()
    n = degree ** 2
    exp, exp_j = torch.empty(3000, n, device="cuda"), torch.empty(3000, 3 * n, device="cuda")
    got, got_j = torch.empty_like(exp), torch.empty_like(exp_j)
    ref.sh_encode_forward(d, exp, 3000, 3, degree, exp_j)
    _shencoder.sh_encode_forward(d, got, 3000, 3, degree, got_j)
    torch.testing.assert_close(got, exp, rtol=RTOL, atol=2e-5)
    torch.testing.assert_close(got_j, exp_j, rtol=RTOL, atol=5e-4)
    grad = torch.randn(3000, n, device="cuda")
    gi_exp, gi_got = torch.zeros(3000, 3, device="cuda"), torch.zeros(3000, 3, device="cuda")
    ref.sh_encode_backward(grad, d, 3000, 3, degree, exp_j, gi_exp)
    _shencoder.sh_encode_backward(grad, d, 3000, 3, degree, exp_j, gi_got)
    torch.testing.assert_close(gi_got, gi_exp, rtol=RTOL, atol=1e-4 * max(1.0, gi_exp.abs().max().item()))


def test_sh_module_normalises_and_broadcasts(cuda):
    enc = SHEncoder(degree=4).cuda()
    raw = torch.randn(5, 32, 3, device="cuda") * 3
    y = enc(raw)
    assert y.shape == (5, 32, 16)
    exp = encoders_np.sh_encode(torch.nn.functional.normalize(raw, dim=-1).reshape(-1, 3).cpu().numpy(), 4)
    np.testing.assert_allclose(y.reshape(-1, 16).cpu().numpy(), exp, rtol=RTOL, atol=2e-5)
    # gradient w.r.t. the direction through dy_dx
    r = raw[0].clone().requires_grad_(True)
    enc(r).square().sum().backward()
    assert r.grad is not None and torch.isfinite(r.grad).all()
    # one evaluation per ray broadcast to its samples (ray_stride) == per-sample evaluation
    from sanerf_b200 import _lib
    dirs = unit_dirs(64, 3).cuda()
    per_sample = dirs[:, None, :].expand(64, 32, 3).reshape(-1, 3).contiguous()
    a = torch.empty(64 * 32, 16, device="cuda"); b = torch.empty_like(a)
    lib = _lib.load()
    _lib.check(lib.sanerf_sh_encode_forward(dirs.data_ptr(), a.data_ptr(), 64 * 32, 3, 4, None, 32,
                                            _lib.current_stream()), "sh")
    _lib.check(lib.sanerf_sh_encode_forward(per_sample.data_ptr(), b.data_ptr(), 64 * 32, 3, 4, None, 0,
                                            _lib.current_stream()), "sh")
    assert torch.equal(a, b)
    assert enc(torch.empty(0, 3, device="cuda")).shape == (0, 16)
    with pytest.raises(AssertionError):
        SHEncoder(degree=9)


def test_freq_vs_oracle_and_reference(cuda, ref_ext):
    enc = FreqEncoder(input_dim=3, degree=6).cuda()
    x = (torch.rand(4000, 3, device="cuda") * 2 - 1).requires_grad_(True)
    y = enc(x)
    assert y.shape == (4000, 39)
    exp = encoders_np.freq_encode(x.detach().cpu().numpy(), 6)
    # __sinf: absolute error grows with |2^f x| -> abs tolerance
    np.testing.assert_allclose(y.detach().cpu().numpy(), exp, rtol=RTOL, atol=5e-5)
    g = torch.randn_like(y)
    y.backward(g)
    exp_g = encoders_np.freq_backward(g.cpu().numpy(), y.detach().cpu().numpy(), 3, 6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), exp_g, rtol=RTOL, atol=1e-4)
    ref = ref_ext("freqencoder")
    r_out = torch.empty(4000, 39, device="cuda")
    ref.freq_encode_forward(x.detach(), 4000, 3, 6, 39, r_out)
    assert torch.equal(y.detach(), r_out)  # same intrinsics, same order: bit-exact
    r_gi, m_gi = torch.zeros(4000, 3, device="cuda"), torch.zeros(4000, 3, device="cuda")
    ref.freq_encode_backward(g, r_out, 4000, 3, 6, 39, r_gi)
    _freqencoder.freq_encode_backward(g, r_out, 4000, 3, 6, 39, m_gi)
    torch.testing.assert_close(m_gi, r_gi, rtol=1e-5, atol=1e-5)
    # D = 2 and a 1-D prefix shape
    e2 = FreqEncoder(input_dim=2, degree=3).cuda()
    x2 = torch.rand(7, 5, 2, device="cuda")
    np.testing.assert_allclose(e2(x2).reshape(-1, 14).cpu().numpy(),
                               encoders_np.freq_encode(x2.reshape(-1, 2).cpu().numpy(), 3), rtol=RTOL, atol=5e-5)


def test_trunc_exp(cuda, ref_cpu):
    x = torch.from_numpy(ref_cpu["texp_x"]).cuda().requires_grad_(True)
    y = trunc_exp(x)
    torch.testing.assert_close(y.detach().cpu(), torch.from_numpy(ref_cpu["texp_y"]), rtol=1e-5, atol=0)
    y.backward(torch.from_numpy(ref_cpu["texp_g"]).cuda())
    torch.testing.assert_close(x.grad.cpu(), torch.from_numpy(ref_cpu["texp_dx"]), rtol=1e-5, atol=1e-30)
    # clamp in the backward only: huge inputs give inf forward but a bounded gradient
    big = torch.tensor([100.0, -100.0, 15.0, -15.0], device="cuda", requires_grad=True)
    trunc_exp(big).sum().backward()
    np.testing.assert_allclose(big.grad.cpu().numpy(), np.exp(np.float32([15, -15, 15, -15])), rtol=1e-5)
    # strided column form used on the 16-wide grid_mlp output (network.py:226)
    from sanerf_b200 import _lib
    f = torch.randn(1000, 16, device="cuda")
    s = torch.empty(1000, device="cuda")
    lib = _lib.load()
    _lib.check(lib.sanerf_trunc_exp_forward(f.data_ptr(), s.data_ptr(), 1000, 16, 0, _lib.current_stream()), "te")
    torch.testing.assert_close(s, torch.exp(f[:, 0]), rtol=1e-6, atol=0)


def test_golden_sh_and_freq(cuda, ref_cpu, ref_gpu):
    d = torch.from_numpy(ref_cpu["sh_dirs"]).float().cuda()
    n = d.shape[0]
    out, jac = torch.empty(n, 64, device="cuda"), torch.empty(n, 192, device="cuda")
    _shencoder.sh_encode_forward(d, out, n, 3, 8, jac)
    np.testing.assert_allclose(out.cpu().numpy(), ref_cpu["sh_out"], rtol=RTOL, atol=3e-5)
    np.testing.assert_allclose(jac.cpu().numpy().reshape(n, 3, 64), ref_cpu["sh_jac"], rtol=RTOL, atol=6e-4)
    dirs = torch.from_numpy(ref_gpu["sh.dirs"]).cuda()
    for deg in (1, 4, 8):
        o = torch.empty(dirs.shape[0], deg * deg, device="cuda")
        _shencoder.sh_encode_forward(dirs, o, dirs.shape[0], 3, deg, None)
        np.testing.assert_allclose(o.cpu().numpy(), ref_gpu[f"sh.{deg}.out"], rtol=RTOL, atol=2e-5)
    xf = torch.from_numpy(ref_gpu["freq_x"]).cuda()
    of = torch.empty(xf.shape[0], 39, device="cuda")
    _freqencoder.freq_encode_forward(xf, xf.shape[0], 3, 6, 39, of)
    np.testing.assert_array_equal(of.cpu().numpy(), ref_gpu["freq_out"])
