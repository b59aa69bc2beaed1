"""Parity of the sm_100a grid encoder (through the C ABI / operator surface) against
(a) the numpy oracle, (b) the UNMODIFIED reference CUDA extension rebuilt for sm_100 (oracle/_ref),
(c) the committed golden vectors, and (d) size-independent properties at BASELINE sizes.

Bars (BASELINE.json north_star): integer indices bit-exact; fp32 outputs <= 1e-3 relative (we get
bit-exact vs the reference kernel); fp16 <= 1e-2; atomically accumulated gradients <= 1e-4 relative.
"""
import numpy as np
import pytest
import torch

import _gridencoder
from gridencoder import GridEncoder
from gridencoder.grid import grid_encode
from oracle import grid_np
from sanerf_b200 import _lib
from sanerf_b200.ops import grid_dump_indices

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-3   # north_star: encoder outputs within 1e-3 relative in fp32
FP16_RTOL = 1e-2   # ... and 1e-2 in fp16
ATOMIC_RTOL = 1e-4  # atomic-order gradient differences


def make_case(D, L, C, base, log2T, desired, dtype=torch.float32, B=3000, seed=0, table_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    scale = grid_np.per_level_scale(desired, base, L)
    offs = grid_np.level_offsets(D, L, scale, base, log2T)
    table = ((torch.rand(int(offs[-1]), C, generator=g) * 2 - 1) * table_scale).to(dtype)
    x = torch.rand(B, D, generator=g)
    if B >= 16:
        x[:6] = torch.tensor([0.0, 1.0, 0.5, 1.0 - 1e-7, 1e-7, 0.999])[:, None]
        x[6] = -0.01   # out-of-range samples (inclusive bounds, gridencoder.cu:109)
        x[7] = 1.0001
    return float(np.log2(scale)), scale, offs, table, x


def run_forward(x, table, offs, S, H, *, layout, gridtype=0, align=False, interp=0, max_level=None,
                want_dydx=False, zero_tail=1, dev="cuda"):
    B, D = x.shape
    L, C = len(offs) - 1, table.shape[1]
    max_level = L if max_level is None else max_level
    xd, td, od = x.to(dev), table.to(dev), torch.from_numpy(offs).to(dev)
    shape = (B, L * C) if layout == _lib.LAYOUT_BLC else (L, B, C)
    out = torch.full(shape, 7.0, device=dev, dtype=table.dtype)
    dydx = torch.empty(B, L * D * C, device=dev, dtype=table.dtype) if want_dydx else None
    lib = _lib.load()
    rc = lib.sanerf_grid_encode_forward(xd.data_ptr(), td.data_ptr(), od.data_ptr(), out.data_ptr(), B, D, C, L,
                                        max_level, S, H, _lib.ptr(dydx), gridtype, int(align), interp,
                                        _lib.SANERF_F16 if table.dtype == torch.float16 else _lib.SANERF_F32,
                                        layout, zero_tail, _lib.current_stream())
    _lib.check(rc, "fwd")
    torch.cuda.synchronize()
    return out, dydx


CASES = [
    # D, L, C, base, log2T, desired, gridtype, align, interp
    (3, 8, 2, 4, 12, 96, 0, False, 0),
    (3, 6, 8, 4, 11, 64, 0, False, 0),
    (3, 6, 2, 4, 10, 64, 1, False, 1),
    (2, 6, 4, 4, 10, 128, 0, True, 0),
    (3, 5, 1, 4, 10, 40, 0, False, 0),
    (3, 4, 16, 4, 9, 32, 0, False, 0),
    (3, 3, 32, 4, 9, 16, 0, False, 1),
    (4, 4, 2, 4, 10, 16, 0, False, 0),
    (5, 3, 1, 3, 10, 8, 0, False, 0),
    (2, 5, 1, 8, 8, 64, 1, True, 1),
    (3, 5, 4, 16, 17, 128, 0, False, 0),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("layout", [_lib.LAYOUT_BLC, _lib.LAYOUT_LBC])
def test_forward_fp32_vs_oracle(cuda, case, layout):
    D, L, C, base, log2T, desired, gridtype, align, interp = case
    S, _, offs, table, x = make_case(D, L, C, base, log2T, desired, B=1500)
    out, dydx = run_forward(x, table, offs, S, base, layout=layout, gridtype=gridtype, align=align, interp=interp,
                            want_dydx=True)
    exp, exp_j = grid_np.grid_encode_forward(x.numpy(), table.numpy(), offs, S, base, gridtype, align, interp,
                                             want_dy_dx=True)
    got = out.cpu().numpy()
    if layout == _lib.LAYOUT_LBC:
        got = got.transpose(1, 0, 2).reshape(x.shape[0], L * C)
    # fp64-emulated FMA in the oracle can differ from a true FMA by 1 ulp: tolerance far below 1e-3
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=2e-6)
    scale = max(1.0, float(np.abs(exp_j).max()))
    np.testing.assert_allclose(dydx.cpu().numpy(), exp_j, rtol=FP32_RTOL, atol=1e-5 * scale)
    assert np.all(got[6] == 0) and np.all(got[7] == 0)  # OOB rows are exactly zero


@pytest.mark.parametrize("case", CASES[:5])
def test_indices_bit_exact_vs_oracle(cuda, case):
    D, L, C, base, log2T, desired, gridtype, align, interp = case
    S, scale, offs, table, x = make_case(D, L, C, base, log2T, desired, B=4000, seed=3)
    rows, geom = grid_dump_indices(x.cuda(), torch.from_numpy(offs).cuda(), scale, base, gridtype, align)
    exp_rows, exp_geom = grid_np.dump_indices(x.numpy(), offs, S, base, gridtype, align)
    np.testing.assert_array_equal(geom.cpu().numpy().astype(np.uint32), exp_geom)
    np.testing.assert_array_equal(rows.cpu().numpy().astype(np.uint32), exp_rows)


def test_level_geometry_of_reference_table_shapes(cuda):
    """Device-side resolution / dense-vs-hash decision for every table the reference builds (network.py:102-216),
    against the fp32 emulation — including the levels where host fp64 and device fp32 disagree."""
    from oracle.make_golden import TABLE_SHAPES
    for name, kw in TABLE_SHAPES.items():
        L, base = kw["num_levels"], kw["base_resolution"]
        scale = grid_np.per_level_scale(kw["desired_resolution"], base, L)
        offs = grid_np.level_offsets(3, L, scale, base, kw["log2_hashmap_size"])
        x = torch.rand(64, 3)
        rows, geom = grid_dump_indices(x.cuda(), torch.from_numpy(offs).cuda(), scale, base, 0, False)
        exp_rows, exp_geom = grid_np.dump_indices(x.numpy(), offs, float(np.log2(scale)), base, 0, False)
        np.testing.assert_array_equal(geom.cpu().numpy().astype(np.uint32), exp_geom, err_msg=name)
        np.testing.assert_array_equal(rows.cpu().numpy().astype(np.uint32), exp_rows, err_msg=name)
    # known answers: main grid offsets and the fp32-vs-fp64 resolution quirk of s_grid
    scale = grid_np.per_level_scale(4096, 16, 16)
    offs = grid_np.level_offsets(3, 16, scale, 16, 19)
    assert list(offs[:6]) == [0, 4096, 17920, 57224, 174880, 532792] and offs[-1] == 6299960
    s = float(np.log2(grid_np.per_level_scale(512, 16, 16)))
    assert [grid_np.device_resolution(l, s, 16) for l in (6, 9, 12, 15)] == [64, 128, 256, 512]


@pytest.mark.parametrize("case", CASES[:7])
def test_forward_bit_exact_vs_reference_extension(cuda, ref_ext, case):
    """fp32 outputs are bit-identical to the unmodified reference kernel (same FMA order)."""
    ref = ref_ext("gridencoder")
    D, L, C, base, log2T, desired, gridtype, align, interp = case
    S, _, offs, table, x = make_case(D, L, C, base, log2T, desired, B=5000, seed=5)
    xd, td, od = x.cuda(), table.cuda(), torch.from_numpy(offs).cuda()
    B = x.shape[0]
    exp = torch.empty(L, B, C, device="cuda")
    exp_j = torch.empty(B, L * D * C, device="cuda")
    ref.grid_encode_forward(xd, td, od, exp, B, D, C, L, L, S, base, exp_j, gridtype, align, interp)
    got = torch.empty(L, B, C, device="cuda")
    got_j = torch.empty(B, L * D * C, device="cuda")
    _gridencoder.grid_encode_forward(xd, td, od, got, B, D, C, L, L, S, base, got_j, gridtype, align, interp)
    torch.cuda.synchronize()
    assert torch.equal(got, exp), f"max abs diff {(got - exp).abs().max().item()}"
    torch.testing.assert_close(got_j, exp_j, rtol=FP32_RTOL, atol=1e-5 * max(1.0, exp_j.abs().max().item()))
    # operator surface ([B, L*C] layout) == permuted reference output
    blc, _ = run_forward(x, table, offs, S, base, layout=_lib.LAYOUT_BLC, gridtype=gridtype, align=align,
                         interp=interp)
    assert torch.equal(blc, exp.permute(1, 0, 2).reshape(B, L * C))


def test_indices_vs_reference_extension_with_probe_tables(cuda, ref_ext):
    """a table whose row r holds (r mod 251, r mod 241) turns the reference kernel's output into
    a checksum of the rows it read; compare with sum_k w_k * probe[rows_new[k]] from OUR dumped indices."""
    ref = ref_ext("gridencoder")
    for (L, base, log2T, desired) in [(16, 16, 19, 4096), (16, 16, 19, 512), (5, 16, 17, 128), (5, 16, 17, 256)]:
        scale = grid_np.per_level_scale(desired, base, L)
        S = float(np.log2(scale))
        offs = grid_np.level_offsets(3, L, scale, base, log2T)
        od = torch.from_numpy(offs).cuda()
        probe = torch.zeros(int(offs[-1]), 2, device="cuda")
        for l in range(L):
            r = torch.arange(int(offs[l + 1] - offs[l]), device="cuda", dtype=torch.float32)
            probe[int(offs[l]):int(offs[l + 1]), 0] = torch.remainder(r, 251.0)
            probe[int(offs[l]):int(offs[l + 1]), 1] = torch.remainder(r, 241.0)
        B = 20000
        x = torch.rand(B, 3, generator=torch.Generator().manual_seed(L + desired)).cuda()
        exp = torch.empty(L, B, 2, device="cuda")
        ref.grid_encode_forward(x, probe, od, exp, B, 3, 2, L, L, S, base, None, 0, False, 0)
        rows, geom = grid_dump_indices(x, od, scale, base, 0, False)
        rows = rows.cpu().numpy().astype(np.uint32).astype(np.int64)     # [B, L, 8]
        xs = x.cpu().numpy()
        exp = exp.cpu().numpy().astype(np.float64)
        for l in range(L):
            res = int(geom[l, 0])
            _, frac, _ = grid_np.locate(xs, res, False, 0)
            w = grid_np.corner_weights(frac).astype(np.float64)        # [B, 8]
            r = rows[:, l]                                             # rows relative to the level
            for ch, p in enumerate((251, 241)):
                mine = (w * (r % p)).sum(1)
                err = np.abs(mine - exp[l, :, ch])
                # a wrong corner with weight >= 1e-3 shifts the checksum by >= 1e-3 unless both moduli collide
                assert err.max() < 2e-3, (desired, l, err.max())


@pytest.mark.parametrize("case", CASES[:8])
def test_backward_vs_oracle_and_reference(cuda, ref_ext, case):
    D, L, C, base, log2T, desired, gridtype, align, interp = case
    S, _, offs, table, x = make_case(D, L, C, base, log2T, desired, B=4000, seed=9)
    B = x.shape[0]
    g = torch.Generator().manual_seed(1)
    grad = torch.rand(B, L * C, generator=g) * 2 - 1
    xd, td, od, gd = x.cuda(), table.cuda(), torch.from_numpy(offs).cuda(), grad.cuda()
    lib = _lib.load()
    got = torch.zeros_like(td)
    rc = lib.sanerf_grid_encode_backward(gd.data_ptr(), xd.data_ptr(), td.data_ptr(), od.data_ptr(), got.data_ptr(),
                                         B, D, C, L, L, S, base, None, None, gridtype, int(align), interp,
                                         _lib.SANERF_F32, _lib.LAYOUT_BLC, _lib.current_stream())
    _lib.check(rc, "bwd")
    exp = grid_np.grid_encode_backward(grad.numpy(), x.numpy(), offs, table.shape[0], C, S, base, gridtype, align,
                                       interp)
    scale = float(np.abs(exp).max())
    np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=ATOMIC_RTOL, atol=ATOMIC_RTOL * scale)
    # reference extension ([L,B,C] gradient layout), including grad_inputs through dy_dx
    ref = ref_ext("gridencoder")
    g_lbc = gd.view(B, L, C).permute(1, 0, 2).contiguous()
    dydx = torch.empty(B, L * D * C, device="cuda")
    tmp = torch.empty(L, B, C, device="cuda")
    ref.grid_encode_forward(xd, td, od, tmp, B, D, C, L, L, S, base, dydx, gridtype, align, interp)
    exp_t, exp_in = torch.zeros_like(td), torch.zeros(B, D, device="cuda")
    ref.grid_encode_backward(g_lbc, xd, td, od, exp_t, B, D, C, L, L, S, base, dydx, exp_in, gridtype, align, interp)
    got_t, got_in = torch.zeros_like(td), torch.zeros(B, D, device="cuda")
    _gridencoder.grid_encode_backward(g_lbc, xd, td, od, got_t, B, D, C, L, L, S, base, dydx, got_in, gridtype,
                                      align, interp)
    torch.cuda.synchronize()
    torch.testing.assert_close(got_t, exp_t, rtol=ATOMIC_RTOL, atol=ATOMIC_RTOL * exp_t.abs().max().item())
    torch.testing.assert_close(got_in, exp_in, rtol=FP32_RTOL, atol=1e-4 * max(1.0, exp_in.abs().max().item()))
    torch.testing.assert_close(got_t, got, rtol=ATOMIC_RTOL, atol=ATOMIC_RTOL * scale)  # both layouts agree


@pytest.mark.parametrize("C", [2, 4, 8])
def test_fp16_tables(cuda, ref_ext, C):
    """Half tables: 1e-2 relative against the fp32 oracle and against the reference's half kernels."""
    ref = ref_ext("gridencoder")
    D, L, base, log2T, desired = 3, 8, 4, 12, 96
    S, _, offs, table, x = make_case(D, L, C, base, log2T, desired, dtype=torch.float16, B=4000, seed=2)
    B = x.shape[0]
    out, _ = run_forward(x, table, offs, S, base, layout=_lib.LAYOUT_BLC)
    exp = grid_np.grid_encode_forward(x.numpy(), table.float().numpy(), offs, S, base)
    np.testing.assert_allclose(out.float().cpu().numpy(), exp, rtol=FP16_RTOL, atol=FP16_RTOL)
    xd, td, od = x.cuda(), table.cuda(), torch.from_numpy(offs).cuda()
    ref_out = torch.empty(L, B, C, device="cuda", dtype=torch.half)
    ref.grid_encode_forward(xd, td, od, ref_out, B, D, C, L, L, S, base, None, 0, False, 0)
    torch.testing.assert_close(out.float(), ref_out.permute(1, 0, 2).reshape(B, L * C).float(), rtol=FP16_RTOL,
                               atol=FP16_RTOL)
    # backward with half2 / vector-half2 reductions
    grad = ((torch.rand(B, L * C, generator=torch.Generator().manual_seed(4)) * 2 - 1) * 0.01).half().cuda()
    got = torch.zeros_like(td)
    lib = _lib.load()
    rc = lib.sanerf_grid_encode_backward(grad.data_ptr(), xd.data_ptr(), td.data_ptr(), od.data_ptr(), got.data_ptr(),
                                         B, D, C, L, L, S, base, None, None, 0, 0, 0, _lib.SANERF_F16,
                                         _lib.LAYOUT_BLC, _lib.current_stream())
    _lib.check(rc, "bwd16")
    exp_t = grid_np.grid_encode_backward(grad.float().cpu().numpy(), x.numpy(), offs, table.shape[0], C, S, base)
    scale = float(np.abs(exp_t).max())
    np.testing.assert_allclose(got.float().cpu().numpy(), exp_t, rtol=5e-2, atol=FP16_RTOL * scale)


def test_tv_and_weight_decay_vs_reference(cuda, ref_ext):
    ref = ref_ext("gridencoder")
    D, L, C, base, log2T, desired = 3, 6, 2, 4, 11, 64
    S, scale, offs, table, x = make_case(D, L, C, base, log2T, desired, B=6000, seed=12)
    xd, td, od = x.cuda(), table.cuda(), torch.from_numpy(offs).cuda()
    B = x.shape[0]
    exp_tv, got_tv = torch.zeros_like(td), torch.zeros_like(td)
    ref.grad_total_variation(xd, td, exp_tv, od, 0.37, B, D, C, L, S, base, 0, False)
    _gridencoder.grad_total_variation(xd, td, got_tv, od, 0.37, B, D, C, L, S, base, 0, False)
    torch.testing.assert_close(got_tv, exp_tv, rtol=1e-3, atol=1e-4 * exp_tv.abs().max().item())
    inc = grid_np.grad_total_variation(x.numpy(), table.numpy(), offs, 0.37, S, base)
    np.testing.assert_allclose(got_tv.cpu().numpy(), inc, rtol=1e-3, atol=1e-4 * np.abs(inc).max())
    exp_wd, got_wd = torch.ones_like(td), torch.ones_like(td)
    ref.grad_weight_decay(td, exp_wd, od, 0.1, td.shape[0], C, L)
    _gridencoder.grad_weight_decay(td, got_wd, od, 0.1, td.shape[0], C, L)
    torch.testing.assert_close(got_wd, exp_wd, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got_wd.cpu().numpy() - 1.0, grid_np.grad_weight_decay(table.numpy(), offs, 0.1),
                               rtol=1e-3, atol=1e-7)


def test_module_autograd_and_edge_cases(cuda):
    enc = GridEncoder(input_dim=3, num_levels=8, level_dim=2, base_resolution=4, log2_hashmap_size=12,
                      desired_resolution=96).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    x = (torch.rand(7, 11, 3, device="cuda") * 2 - 1) * 2  # bound = 2 like the contracted scene
    y = enc(x, bound=2)
    assert y.shape == (7, 11, 16)
    exp = grid_np.grid_encode_forward(((x + 2) / 4).reshape(-1, 3).cpu().numpy(), enc.embeddings.detach().cpu().numpy(),
                                      enc.offsets.cpu().numpy(), float(np.log2(enc.per_level_scale)), 4)
    np.testing.assert_allclose(y.detach().cpu().numpy().reshape(-1, 16), exp, rtol=1e-5, atol=2e-6)
    (y * torch.arange(16, device="cuda")).sum().backward()
    assert enc.embeddings.grad is not None and enc.embeddings.grad.abs().sum() > 0
    # max_level: the tail is exactly zero and gets no gradient
    y4 = enc(x, bound=2, max_level=4)
    assert torch.equal(y4[..., :8], y[..., :8]) and torch.all(y4[..., 8:] == 0)
    # empty batch
    assert enc(torch.empty(0, 3, device="cuda")).shape == (0, 16)
    # input gradient path (dy_dx) against finite differences of the oracle-validated forward
    xg = (torch.rand(64, 3, device="cuda") * 1.6 - 0.8).requires_grad_(True)
    enc(xg, bound=1).sum().backward()
    eps = 2e-4
    num = torch.zeros_like(xg)
    for d in range(3):
        dx = torch.zeros_like(xg); dx[:, d] = eps
        num[:, d] = (enc(xg.detach() + dx).sum(-1) - enc(xg.detach() - dx).sum(-1)) / (2 * eps)
    ok = (num - xg.grad).abs() < 0.05 * num.abs().max() + 1e-2
    assert ok.float().mean() > 0.8  # piecewise-linear field: finite differences straddle cell borders sometimes
    # AMP: half tables when autocast is on and C is even (grid.py:43-46)
    with torch.autocast("cuda", dtype=torch.float16):
        yh = enc(x, bound=2)
    assert yh.dtype == torch.float16
    torch.testing.assert_close(yh.float(), y, rtol=FP16_RTOL, atol=FP16_RTOL)
    # TV / WD hooks mutate .grad in place
    before = enc.embeddings.grad.clone()
    enc.grad_weight_decay(0.1)
    enc.grad_total_variation(1e-3, B=1000)
    assert not torch.equal(before, enc.embeddings.grad)
    # unsupported sizes raise like the reference's std::runtime_error
    with pytest.raises(RuntimeError, match="C must be"):
        grid_encode(torch.rand(4, 3, device="cuda"), torch.zeros(64, 3, device="cuda"),
                    torch.tensor([0, 64], dtype=torch.int32, device="cuda"), 2.0, 4)


def test_golden_vectors_from_reference_extension(cuda, ref_gpu):
    """Committed fixtures produced by the reference extension on a B200 (oracle/make_golden.py --gpu)."""
    names = sorted({k.split(".")[1] for k in ref_gpu.files if k.startswith("grid.")})
    assert names
    for name in names:
        D, L, C, base, log2T, desired, gridtype, align, interp, is_half = ref_gpu[f"grid.{name}.meta"].tolist()
        S = float(ref_gpu[f"grid.{name}.S"])
        offs = ref_gpu[f"grid.{name}.offsets"]
        x = torch.from_numpy(ref_gpu[f"grid.{name}.x"])
        table = torch.from_numpy(ref_gpu[f"grid.{name}.table"])
        out, dydx = run_forward(x, table, offs, S, base, layout=_lib.LAYOUT_LBC, gridtype=gridtype, align=bool(align),
                                interp=interp, want_dydx=True)
        exp = torch.from_numpy(ref_gpu[f"grid.{name}.out_LBC"]).cuda()
        if is_half:
            torch.testing.assert_close(out.float(), exp.float(), rtol=FP16_RTOL, atol=FP16_RTOL)
        else:
            assert torch.equal(out, exp), name
        B = x.shape[0]
        grad = torch.from_numpy(ref_gpu[f"grid.{name}.grad_LBC"]).cuda()
        xd, td, od = x.cuda(), table.cuda(), torch.from_numpy(offs).cuda()   # keep the device buffers alive
        got = torch.zeros_like(td)
        lib = _lib.load()
        rc = lib.sanerf_grid_encode_backward(grad.data_ptr(), xd.data_ptr(), td.data_ptr(),
                                             od.data_ptr(), got.data_ptr(), B, D, C, L, L, S,
                                             base, None, None, gridtype, align, interp,
                                             _lib.SANERF_F16 if is_half else _lib.SANERF_F32, _lib.LAYOUT_LBC,
                                             _lib.current_stream())
        _lib.check(rc, "bwd")
        torch.cuda.synchronize()
        exp_t = torch.from_numpy(ref_gpu[f"grid.{name}.grad_table"]).cuda().float()
        tol = 3e-2 if is_half else ATOMIC_RTOL
        torch.testing.assert_close(got.float(), exp_t, rtol=tol, atol=tol * exp_t.abs().max().item())


def test_properties_at_baseline_size(cuda):
    """BASELINE cfg1/cfg2 shape: main grid L16 F2 T2^19, B = 262,144.  Size-independent checks:
    linearity in the table, and <encode(T), G> == <T, backward(G)> (the scatter is the adjoint of the gather)."""
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, log2_hashmap_size=19, desired_resolution=4096).cuda()
    B = 262144
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(B, 3, device="cuda", generator=g)
    T1 = torch.randn(enc.embeddings.shape, device="cuda", generator=g)
    T2 = torch.randn(enc.embeddings.shape, device="cuda", generator=g)
    args = (enc.offsets, enc.per_level_scale, enc.base_resolution)
    y1, y2 = grid_encode(x, T1, *args), grid_encode(x, T2, *args)
    y12 = grid_encode(x, 2.0 * T1 - T2, *args)
    torch.testing.assert_close(y12, 2.0 * y1 - y2, rtol=1e-4, atol=1e-4)
    G = torch.randn(B, 32, device="cuda", generator=g)
    T1r = T1.clone().requires_grad_(True)
    grid_encode(x, T1r, *args).backward(G)
    lhs = (y1.double() * G.double()).sum()
    rhs = (T1.double() * T1r.grad.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-4
    # checksum of checksums: the gradient mass equals the mass of G per level (weights sum to 1)
    mass = T1r.grad.double().sum(0)
    torch.testing.assert_close(mass, G.double().view(B, 16, 2).sum((0, 1)), rtol=1e-5, atol=1e-2)
