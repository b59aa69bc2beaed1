"""tcgen05 (UMMA) tile products and the fused tensor-core MLP head against plain fp32 matmuls."""
import pytest
import torch

from sanerf_b200 import _lib

pytestmark = pytest.mark.gpu


def _selftest(mode, M, N, K, A, B):
    lib = _lib.load()
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.sanerf_umma_selftest(mode, M, N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(),
                                  _lib.current_stream(A.device))
    _lib.check(rc, "umma_selftest")
    torch.cuda.synchronize()
    return D


def _ints(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-8, 9, shape, generator=g).float().cuda()      # exact in tf32, products exact in fp32


@pytest.mark.parametrize("N,K", [(64, 32), (64, 64), (16, 64), (32, 8), (128, 64), (96, 96)])
def test_umma_k_major_exact(cuda, N, K):
    A, B = _ints(128, K, seed=1), _ints(N, K, seed=2)
    D = _selftest(0, 128, N, K, A, B)
    assert torch.equal(D, A @ B.t())


@pytest.mark.parametrize("M,N,K", [(64, 32, 128), (64, 64, 128), (64, 16, 128), (128, 64, 64), (64, 16, 8)])
def test_umma_mn_major_exact(cuda, M, N, K):
    At, Bt = _ints(K, M, seed=3), _ints(K, N, seed=4)
    D = _selftest(1, M, N, K, At, Bt)
    assert torch.equal(D, At.t() @ Bt)


@pytest.mark.parametrize("N,K", [(64, 64), (16, 64), (64, 32)])
def test_umma_a_from_tmem_exact(cuda, N, K):
    A, B = _ints(128, K, seed=5), _ints(N, K, seed=6)
    D = _selftest(2, 128, N, K, A, B)
    assert torch.equal(D, A @ B.t())


def test_umma_3xtf32_is_fp32_accurate(cuda):
    g = torch.Generator().manual_seed(7)
    A, B = torch.randn(128, 64, generator=g).cuda(), torch.randn(64, 64, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    one = _selftest(0, 128, 64, 64, A, B).double()
    three = _selftest(3, 128, 64, 64, A, B).double()
    err1 = ((one - ref).abs().max() / ref.abs().max()).item()
    err3 = ((three - ref).abs().max() / ref.abs().max()).item()
    assert err1 < 5e-3                      # single-pass tf32: ~2^-10 relative operand truncation
    assert err3 < 2e-6, (err1, err3)        # split product: fp32-level


# ------------------------------------------------------------------------------------------ fused field head
def _head_setup(B, seed=0, table_scale=1.0):
    from gridencoder import GridEncoder
    torch.manual_seed(seed)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                      desired_resolution=4096).cuda()
    with torch.no_grad():
        enc.embeddings.normal_(0, table_scale)
    w1 = (torch.randn(64, 32) / 32 ** 0.5).cuda().requires_grad_(True)
    w2 = (torch.randn(64, 64) / 64 ** 0.5).cuda().requires_grad_(True)
    w3 = (torch.randn(16, 64) / 64 ** 0.5).cuda().requires_grad_(True)
    x01 = torch.rand(B, 3, device="cuda")
    x01[::97] = 1.5            # out-of-range samples encode to zero (gridencoder.cu:105-130)
    return enc, w1, w2, w3, x01


def _head_call(x01, enc, w1, w2, w3, precision, enc_in=None, want_enc=True, want_hidden=False):
    import numpy as np
    lib = _lib.load()
    from sanerf_b200.fused import tcm_rows, tcm_to_rows
    B = x01.shape[0] if x01 is not None else enc_in.shape[0]
    Bp = tcm_rows(B)                           # saved activations: tile-chunk-major, whole tiles
    out = torch.full((B, 16), float("nan"), device="cuda")
    enc_out = torch.full((Bp, 32), float("nan"), device="cuda") if want_enc else None
    h1 = torch.full((Bp, 64), float("nan"), device="cuda") if want_hidden else None
    h2 = torch.full((Bp, 64), float("nan"), device="cuda") if want_hidden else None
    rc = lib.sanerf_field_head_forward(_lib.ptr(x01), enc.embeddings.data_ptr(), enc.offsets.data_ptr(),
                                       float(np.log2(enc.per_level_scale)), int(enc.base_resolution), _lib.ptr(enc_in),
                                       w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B, _lib.ptr(enc_out), _lib.ptr(h1),
                                       _lib.ptr(h2), out.data_ptr(), precision, _lib.current_stream(out.device))
    _lib.check(rc, "field_head_forward")
    torch.cuda.synchronize()
    if want_enc:
        enc_out = tcm_to_rows(enc_out, B, 32)
    if want_hidden:
        return out, enc_out, h1, h2          # h1 / h2 stay in the kernels' private layout
    return out, enc_out


def _mlp64(e, w1, w2, w3):
    h = torch.relu(e.double() @ w1.double().t())
    h = torch.relu(h @ w2.double().t())
    return h @ w3.double().t()


@pytest.mark.parametrize("B", [1000, 128 * 148 * 3 + 77])
def test_field_head_forward_matches_fp32_path(cuda, B):
    from gridencoder.grid import grid_encode
    enc, w1, w2, w3, x01 = _head_setup(B)
    ref_enc = grid_encode(x01, enc.embeddings, enc.offsets, enc.per_level_scale, enc.base_resolution).detach()
    ref = _mlp64(ref_enc, w1, w2, w3)
    out, enc_out = _head_call(x01, enc, w1.detach(), w2.detach(), w3.detach(), 0)
    assert torch.equal(enc_out, ref_enc)                       # the gather is the reference kernel's arithmetic
    scale = ref.abs().max().item()
    assert ((out.double() - ref).abs().max().item()) < 4e-6 * scale     # 3xTF32: fp32-level (three chained layers)
    fast, _ = _head_call(x01, enc, w1.detach(), w2.detach(), w3.detach(), 1, want_enc=False)
    assert ((fast.double() - ref).abs().max().item()) < 1e-2 * scale    # single tf32 pass
    # encoding supplied by the caller instead of gathered
    out2, _ = _head_call(None, enc, w1.detach(), w2.detach(), w3.detach(), 0, enc_in=ref_enc, want_enc=False)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("B", [300, 128 * 148 * 2 + 5])
def test_field_head_backward_matches_autograd(cuda, B):
    enc, w1, w2, w3, x01 = _head_setup(B, seed=1)
    g = torch.Generator(device="cuda").manual_seed(3)
    e = torch.randn(B, 32, device="cuda", generator=g)
    g_out = torch.randn(B, 16, device="cuda", generator=g)
    lib = _lib.load()
    # the forward saves the hidden activations the backward consumes
    _, _, h1, h2 = _head_call(None, enc, w1.detach(), w2.detach(), w3.detach(), 0, enc_in=e, want_enc=False,
                              want_hidden=True)
    W1, W2, W3 = (w.detach().double() for w in (w1, w2, w3))
    h1_ref = torch.relu(e.double() @ W1.t())
    h2_ref = torch.relu(h1_ref @ W2.t())
    from sanerf_b200.fused import rows_to_tcm, tcm_to_rows
    h1_tcm, h2_tcm, e_tcm = h1, h2, rows_to_tcm(e)
    h1, h2 = tcm_to_rows(h1_tcm, B, 64), tcm_to_rows(h2_tcm, B, 64)
    assert ((h1.double() - h1_ref).abs().max() / h1_ref.abs().max()).item() < 2e-6
    assert ((h2.double() - h2_ref).abs().max() / h2_ref.abs().max()).item() < 3e-6
    # fp64 backward THROUGH THE SAME ReLU sign pattern (a pre-activation within rounding of zero may legitimately
    # land on either side in any fp32 evaluation; that is not what this test is about)
    G3 = g_out.double()
    G2 = (G3 @ W3) * (h2 > 0)
    G1 = (G2 @ W2) * (h1 > 0)
    ref_enc, ref_w = G1 @ W1, (G1.t() @ e.double(), G2.t() @ h1.double(), G3.t() @ h2.double())
    for precision, tol_e, tol_w in ((0, 5e-6, 1e-5), (1, 5e-3, 5e-3)):
        g_enc = torch.full((B, 32), float("nan"), device="cuda")
        gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
        rc = lib.sanerf_field_head_backward(e_tcm.data_ptr(), h1_tcm.data_ptr(), h2_tcm.data_ptr(), g_out.data_ptr(), w1.data_ptr(),
                                            w2.data_ptr(), w3.data_ptr(), B, g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(),
                                            gw[1].data_ptr(), gw[2].data_ptr(), precision, _lib.current_stream(e.device))
        _lib.check(rc, "field_head_backward")
        torch.cuda.synchronize()
        def rel(a, b):
            return ((a.double() - b).abs().max() / b.abs().max()).item()
        assert rel(g_enc, ref_enc) < tol_e, precision
        for got, ref, name in zip(gw, ref_w, ("w1", "w2", "w3")):
            assert rel(got, ref) < tol_w, (precision, name)


def test_field_head_autograd_equals_unfused(cuda):
    """The drop-in check: the model's head with the fused kernel == GridEncoder + nn.Linear MLP, values and grads."""
    from nerf.network import MLP
    from sanerf_b200 import fused
    enc, _, _, _, x01 = _head_setup(5000, seed=4, table_scale=0.3)
    mlp = MLP(32, 16, 64, 3, bias=False).cuda()
    x = x01.view(50, 100, 3)
    out = fused.field_head(x, enc, mlp)
    from gridencoder.grid import grid_encode
    ref = mlp(grid_encode(x.reshape(-1, 3), enc.embeddings, enc.offsets, enc.per_level_scale,
                          enc.base_resolution).view(50, 100, 32))
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5 * ref.abs().max().item())
    go = torch.randn_like(ref)
    params = [enc.embeddings] + [l.weight for l in mlp.net]
    g_ref = torch.autograd.grad(ref, params, go)
    g_new = torch.autograd.grad(out, params, go)
    for a, b in zip(g_new, g_ref):
        assert ((a - b).norm() / b.norm()).item() < 1e-4


def test_field_head_backward_fused_scatter_equals_separate_kernel(cuda):
    """Optional mode of the backward kernel: the hash-grid scatter runs inside it (gradient of the encoding taken from
    tensor memory) instead of sanerf_grid_encode_backward on a [B,32] buffer."""
    import numpy as np
    from sanerf_b200.fused import rows_to_tcm
    B = 128 * 148 * 2 + 37
    enc, w1, w2, w3, x01 = _head_setup(B, seed=6)
    # ray-like ordering so that the warp aggregation actually merges lanes
    x01 = (torch.rand(B // 32 + 1, 1, 3, device="cuda") + torch.linspace(0, 0.05, 32, device="cuda").view(1, 32, 1)).reshape(-1, 3)[:B]
    x01 = x01.clamp(0, 1).contiguous()
    x01[::97] = 1.5
    lib = _lib.load()
    _, enc_rows, h1, h2 = _head_call(x01, enc, w1.detach(), w2.detach(), w3.detach(), 0, want_hidden=True)
    e_tcm = rows_to_tcm(enc_rows.contiguous())
    g_out = torch.randn(B, 16, device="cuda")
    S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
    st = _lib.current_stream(x01.device)
    g_enc = torch.empty(B, 32, device="cuda")
    gw_a = [torch.zeros_like(w) for w in (w1, w2, w3)]
    rc = lib.sanerf_field_head_backward(e_tcm.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(),
                                        w2.data_ptr(), w3.data_ptr(), B, g_enc.data_ptr(), None, None, 0.0, 0, None,
                                        gw_a[0].data_ptr(), gw_a[1].data_ptr(), gw_a[2].data_ptr(), 0, st)
    _lib.check(rc, "field_head_backward")
    gt_a = torch.zeros_like(enc.embeddings)
    rc = lib.sanerf_grid_encode_backward(g_enc.data_ptr(), x01.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(),
                                         gt_a.data_ptr(), B, 3, 2, 16, 16, S, H, None, None, 0, 0, 0, _lib.SANERF_F32,
                                         _lib.LAYOUT_BLC, st)
    _lib.check(rc, "grid_encode_backward")
    gt_b = torch.zeros_like(enc.embeddings)
    gw_b = [torch.zeros_like(w) for w in (w1, w2, w3)]
    rc = lib.sanerf_field_head_backward(e_tcm.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(),
                                        w2.data_ptr(), w3.data_ptr(), B, None, x01.data_ptr(), enc.offsets.data_ptr(), S, H,
                                        gt_b.data_ptr(), gw_b[0].data_ptr(), gw_b[1].data_ptr(), gw_b[2].data_ptr(), 0, st)
    _lib.check(rc, "field_head_backward")
    torch.cuda.synchronize()
    assert ((gt_a - gt_b).norm() / gt_a.norm()).item() < 1e-5          # same contributions, different summation order
    torch.testing.assert_close(gt_b, gt_a, rtol=1e-3, atol=1e-5 * gt_a.abs().max().item())
    for a, b in zip(gw_a, gw_b):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-5 * a.abs().max().item())


# ------------------------------------------------------------------------ wide MLP head (csrc/gemm_tc.cu)
@pytest.mark.parametrize("M,N,K", [(300, 70, 45), (128, 64, 32), (4096, 256, 163), (257, 419, 256)])
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, True)])
def test_gemm_tc_fp32_parity(cuda, M, N, K, a_trans, b_trans):
    """3xTF32 tensor-core GEMM against an fp64 product: far inside the 1e-3 fp32 bar, for ragged / unaligned shapes."""
    from sanerf_b200 import fused
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = (A.double() @ B.double().t())
    As = A.t().contiguous() if a_trans else A
    Bs = B.t().contiguous() if b_trans else B
    C = torch.full((M, N), 7.0, device="cuda")
    fused.gemm_tc(As, Bs, C, M, N, K, a_trans=a_trans, b_trans=b_trans, bias=bias, act=True, slope=0.01)
    exp = torch.nn.functional.leaky_relu(ref + bias.double(), 0.01)
    scale = float(ref.abs().max())
    assert float((C.double() - exp).abs().max()) < 2e-6 * scale * K ** 0.5
    # one tf32 pass: the 1e-2 class
    C1 = torch.empty(M, N, device="cuda")
    fused.gemm_tc(As, Bs, C1, M, N, K, a_trans=a_trans, b_trans=b_trans, precision=1)
    assert float((C1.double() - ref).abs().max()) < 2e-3 * scale
    # masked data-gradient epilogue and split accumulating epilogue
    mask = torch.randn(M, N, generator=g).cuda()
    C2 = torch.empty(M, N, device="cuda")
    fused.gemm_tc(As, Bs, C2, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=1, mask=mask, mask_cols=N // 2, slope=0.01)
    exp2 = ref.clone()
    exp2[:, :N // 2] *= torch.where(mask[:, :N // 2] > 0, 1.0, 0.01).double()
    assert float((C2.double() - exp2).abs().max()) < 2e-6 * scale * K ** 0.5
    C3 = torch.ones(M, N, device="cuda")
    fused.gemm_tc(As, Bs, C3, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=2, k_splits=3)
    assert float((C3.double() - (ref + 1)).abs().max()) < 2e-6 * scale * K ** 0.5


def test_skip_mlp_tensor_core_equals_torch(cuda):
    """samvit_mlp shape (network.py:120-123): forward, input gradient and every parameter gradient vs nn.Linear."""
    import copy
    from nerf.network import SkipConnMLP
    torch.manual_seed(0)
    ref = SkipConnMLP(163, 256, 256, 5, skip_layers=[2], bias=True).cuda()
    tc = copy.deepcopy(ref)
    tc.tc = True
    x = torch.randn(500, 163, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), tc(xb)
    torch.testing.assert_close(yb, ya, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    torch.testing.assert_close(xb.grad, xa.grad, rtol=1e-4, atol=1e-5)
    for (n, p), (_, q) in zip(ref.named_parameters(), tc.named_parameters()):
        assert ((p.grad - q.grad).norm() / p.grad.norm()).item() < 1e-5, n


def test_layernorm_mse_fused_equals_torch(cuda):
    """LayerNorm(256) + permute + mse_loss (network.py:122, utils.py:1100-1106), forward and backward, in one kernel."""
    from sanerf_b200 import fused
    torch.manual_seed(1)
    h, w = 12, 10
    ln = torch.nn.LayerNorm(256).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    x = torch.randn(h * w, 256, device="cuda") * 3 + 1
    target = torch.randn(1, 256, h, w, device="cuda")
    xr = x.clone().requires_grad_(True)
    y_ref = ln(xr)
    loss_ref = torch.nn.functional.mse_loss(y_ref.view(h, w, -1).permute(2, 0, 1).unsqueeze(0), target)
    loss_ref.backward()
    gw_ref, gb_ref = ln.weight.grad.clone(), ln.bias.grad.clone()
    ln.weight.grad.zero_(); ln.bias.grad.zero_()
    loss = torch.zeros(1, device="cuda")
    y = torch.empty_like(x)
    g_x = fused.layernorm_mse(x, ln, target, loss, y)
    torch.testing.assert_close(y, y_ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(loss[0], loss_ref.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(g_x, xr.grad, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(ln.weight.grad, gw_ref, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(ln.bias.grad, gb_ref, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("M,N,K,lda,ldb", [(4096, 256, 256, 256, 256), (300, 256, 256, 256, 256), (4096, 256, 419, 420, 420),
                                           (128, 64, 40, 40, 40), (1000, 200, 163, 420, 164)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_gemm_tma_path_equals_register_path(cuda, M, N, K, lda, ldb, precision):
    """The TMA-fed forward GEMM (csrc/gemm_tma.cu: cp.async.bulk.tensor boxes into 128-byte-swizzled stages, hi / lo planes
    derived in shared memory) against the register-staged kernel (csrc/gemm_tc.cu; precision bit 7 forces it) and fp64:
    row / column / K tails are zero-filled by the TMA, strided operands (lda, ldb > K) are addressed through the tensor map."""
    from sanerf_b200 import fused
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, lda, generator=g).cuda()[:, :K]
    B = (torch.randn(N, ldb, generator=g) / K ** 0.5).cuda()[:, :K]
    bias = torch.randn(N, generator=g).cuda()
    prec = fused.PRECISION_IDS[precision]
    out_tma, out_reg = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    fused.gemm_tc(A, B, out_tma, M, N, K, bias=bias, act=True, precision=prec)
    fused.gemm_tc(A, B, out_reg, M, N, K, bias=bias, act=True, precision=prec | 128)
    ref = torch.nn.functional.leaky_relu(A.double() @ B.double().t() + bias.double(), 0.01)
    tol = 1e-5 if precision == "fp32" else 3e-3
    scale = float(ref.abs().max())
    torch.testing.assert_close(out_tma.double(), ref, rtol=tol, atol=tol * scale)
    torch.testing.assert_close(out_tma, out_reg, rtol=1e-6 if precision == "fp32" else 1e-5, atol=1e-6 * scale)


@pytest.mark.parametrize("M,N,K,ldb", [(4096, 256, 256, 256), (4096, 419, 256, 420), (300, 128, 256, 164), (200, 64, 40, 64)])
@pytest.mark.parametrize("masked", [False, True])
def test_gemm_tma_data_gradient_equals_register_path(cuda, M, N, K, ldb, masked):
    """Data-gradient products C = (A . Bt) * act'(mask) with Bt = an nn.Linear weight as stored ([K, N] row-major): the TMA
    path loads it as an MN-major operand (32 x 32 boxes, 128-byte swizzle with 32-byte atoms) — against the register-staged
    kernel and fp64; N / K / M tails and a padded leading dimension included."""
    from sanerf_b200 import fused
    g = torch.Generator().manual_seed(M + N + K + ldb)
    A = torch.randn(M, K, generator=g).cuda()
    Bt = (torch.randn(K, ldb, generator=g) / K ** 0.5).cuda()[:, :N]
    mask = torch.randn(M, N, generator=g).cuda()
    kw = dict(b_trans=True, epilogue=1, mask=mask, mask_cols=N - 3) if masked else dict(b_trans=True)
    out_tma, out_reg = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    fused.gemm_tc(A, Bt, out_tma, M, N, K, precision=0, **kw)
    fused.gemm_tc(A, Bt, out_reg, M, N, K, precision=128, **kw)
    ref = A.double() @ Bt.double()
    if masked:
        d = torch.where(mask.double() > 0, 1.0, 0.01)
        d[:, N - 3:] = 1.0
        ref = ref * d
    scale = float(ref.abs().max())
    torch.testing.assert_close(out_tma.double(), ref, rtol=1e-5, atol=1e-5 * scale)
    torch.testing.assert_close(out_tma, out_reg, rtol=1e-6, atol=1e-6 * scale)
