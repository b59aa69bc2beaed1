"""CPU tests: numpy grid / SH / freq oracles against known answers from the reference source, against each
other (numpy vs torch restatement), and against the golden vectors from the reference (CPU + GPU runs)."""
import math

import numpy as np
import pytest
import torch

from oracle import encoders_np, grid_np
from oracle import render_torch as R


def test_known_answers_from_reference_constants():
    # primes and hash of gridencoder.cu:49-55 (uint32 wrap-around)
    assert [int(p) for p in grid_np.PRIMES[:3]] == [1, 2654435761, 805459861]
    pg = np.array([[3, 5, 7]], dtype=np.uint64)
    mult = grid_np.PRIMES[:3]
    rows = grid_np.corner_rows(pg, 1 << 20, 1 << 19, mult, True)
    h = (3 * 1) ^ ((5 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)
    assert int(rows[0, 0]) == h % (1 << 19)
    # main-grid offsets (grid.py:124-134) and table sizes (§8 header)
    scale = grid_np.per_level_scale(4096, 16, 16)
    offs = grid_np.level_offsets(3, 16, scale, 16, 19)
    assert list(offs[:7]) == [0, 4096, 17920, 57224, 174880, 532792, 1057080] and int(offs[-1]) == 6299960
    offs_s = grid_np.level_offsets(3, 16, grid_np.per_level_scale(512, 16, 16), 16, 19)
    assert int(offs_s[-1]) == 5258512
    # dense vs hashed levels: main grid 0-4 dense, 5-15 hashed; s_grid 0-6 dense
    S = float(np.log2(scale))
    hashed = [grid_np.level_geometry(offs, l, S, 16, 3, 0)[3] for l in range(16)]
    assert hashed == [False] * 5 + [True] * 11
    Ss = float(np.log2(grid_np.per_level_scale(512, 16, 16)))
    hashed_s = [grid_np.level_geometry(offs_s, l, Ss, 16, 3, 0)[3] for l in range(16)]
    assert hashed_s == [False] * 7 + [True] * 9
    # fp32 device resolution vs fp64 host resolution quirk
    assert [grid_np.device_resolution(l, Ss, 16) for l in (6, 9, 12, 15)] == [64, 128, 256, 512]
    assert grid_np.device_resolution(15, S, 16) == 4096
    assert [grid_np.device_resolution(l, S, 16) for l in range(5)] == [16, 24, 34, 49, 71]
    # uint32 stride wrap (SURVEY §7 hard parts): T=2^22, res=1956 stays hashed
    res, rows_, mult_, hashed_, covered = 1956, 1 << 22, None, None, None
    stride = 1
    for d in range(3):
        if stride <= rows_:
            stride = (stride * res) & 0xFFFFFFFF
    assert stride == (1956 ** 3) % (1 << 32) and stride > rows_


def test_offsets_match_reference_module(ref_cpu):
    from oracle.make_golden import TABLE_SHAPES
    keys = [k for k in ref_cpu.files if k.startswith("offsets.")]
    if not keys:
        pytest.skip("offset goldens need oracle/_ref at generation time")
    for k in keys:
        name = k.split(".", 1)[1]
        kw = TABLE_SHAPES[name]
        scale = grid_np.per_level_scale(kw["desired_resolution"], kw["base_resolution"], kw["num_levels"])
        assert scale == float(ref_cpu[f"scale.{name}"])
        offs = grid_np.level_offsets(3, kw["num_levels"], scale, kw["base_resolution"], kw["log2_hashmap_size"])
        np.testing.assert_array_equal(offs, ref_cpu[k])
        # the product's own host code
        from gridencoder.grid import level_offsets
        np.testing.assert_array_equal(level_offsets(3, kw["num_levels"], scale, kw["base_resolution"],
                                                    kw["log2_hashmap_size"]), ref_cpu[k])


@pytest.mark.parametrize("cfg", [(3, 6, 2, 4, 10, 64, 0, False, 0), (3, 5, 4, 4, 9, 48, 1, False, 1),
                                 (2, 5, 2, 4, 8, 64, 0, True, 0)])
def test_numpy_and_torch_restatements_agree(cfg):
    D, L, C, base, log2T, desired, gridtype, align, interp = cfg
    rng = np.random.default_rng(0)
    scale = grid_np.per_level_scale(desired, base, L)
    S = float(np.log2(scale))
    offs = grid_np.level_offsets(D, L, scale, base, log2T)
    table = rng.uniform(-1, 1, size=(int(offs[-1]), C)).astype(np.float32)
    x = rng.uniform(0, 1, size=(500, D)).astype(np.float32)
    x[0] = -0.1
    a = grid_np.grid_encode_forward(x, table, offs, S, base, gridtype, align, interp)
    tt = torch.from_numpy(table).requires_grad_(True)
    b = R.grid_encode(torch.from_numpy(x), tt, offs.tolist(), S, base, gridtype, align, interp)
    np.testing.assert_allclose(a, b.detach().numpy(), rtol=1e-5, atol=5e-6)  # torch restatement has no FMA
    g = rng.uniform(-1, 1, size=a.shape).astype(np.float32)
    b.backward(torch.from_numpy(g))
    gt = grid_np.grid_encode_backward(g, x, offs, table.shape[0], C, S, base, gridtype, align, interp)
    np.testing.assert_allclose(gt, tt.grad.numpy(), rtol=1e-4, atol=1e-5)
    # dy_dx against central differences of the (piecewise trilinear) forward, away from cell borders
    if interp == 0 and not align:
        _, jac = grid_np.grid_encode_forward(x, table, offs, S, base, gridtype, align, interp, want_dy_dx=True)
        gi = grid_np.grid_input_backward(g, jac, 500, D, C, L)
        eps = 1e-4
        num = np.zeros_like(gi)
        for d in range(D):
            dx = np.zeros_like(x); dx[:, d] = eps
            hi = grid_np.grid_encode_forward(x + dx, table.astype(np.float64).astype(np.float32), offs, S, base)
            lo = grid_np.grid_encode_forward(x - dx, table, offs, S, base)
            num[:, d] = ((hi.astype(np.float64) - lo) * g).sum(1) / (2 * eps)
        close = np.abs(num - gi) < 0.02 * np.abs(num).max() + 0.05
        assert close[1:].mean() > 0.8


def test_sh_closed_form_vs_reference_polynomials(ref_cpu):
    """Golden = the 64 polynomials + 192 derivatives evaluated from the text of shencoder.cu:50-349."""
    out, jac = encoders_np.sh_encode(ref_cpu["sh_dirs"], 8, want_jacobian=True)
    np.testing.assert_allclose(out, ref_cpu["sh_out"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(jac, ref_cpu["sh_jac"], rtol=0, atol=1e-11)
    # constants quoted in SURVEY §8 c4
    assert abs(encoders_np.sh_norm(0, 0) - 0.28209479177387814) < 1e-16
    assert abs(-encoders_np.sh_norm(1, 1) - 0.48860251190291987) < 1e-15
    # torch fp32 restatement used by the CPU baseline
    d = torch.from_numpy(ref_cpu["sh_dirs"]).float()
    np.testing.assert_allclose(R.sh_encode(d, 4).numpy(), ref_cpu["sh_out"][:, :16], rtol=1e-4, atol=1e-5)


def test_product_sh_norm_table_matches_formula():
    """The constants compiled into csrc/sh_common.cuh are the closed-form N_l^m."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "segment-anything-nerf_b200", "csrc",
                            "sh_common.cuh")).read()
    body = src[src.index("kShNorm[8][8] = {"):]
    body = body[:body.index("};")]
    vals = [float(v) for v in re.findall(r"(-?\d+\.\d+(?:e[-+]?\d+)?)f", body)]
    assert len(vals) == 64
    for l in range(8):
        for m in range(8):
            exp = encoders_np.sh_norm(l, m) if m <= l else 0.0
            assert math.isclose(vals[l * 8 + m], exp, rel_tol=1e-15, abs_tol=1e-18)


def test_freq_oracle_vs_reference_torch_encoder(ref_cpu):
    out = encoders_np.freq_encode(ref_cpu["freq_x"], 6)
    np.testing.assert_allclose(out, ref_cpu["freq_out"], rtol=1e-6, atol=1e-6)
    g = np.random.default_rng(0).normal(size=out.shape)
    x = torch.from_numpy(ref_cpu["freq_x"]).double().requires_grad_(True)
    R.freq_encode(x, 6).backward(torch.from_numpy(g))
    np.testing.assert_allclose(encoders_np.freq_backward(g, R.freq_encode(x, 6).detach().numpy(), 3, 6),
                               x.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_grid_oracle_vs_reference_extension_goldens(ref_gpu):
    """numpy oracle vs what the unmodified reference CUDA extension produced on a B200."""
    names = sorted({k.split(".")[1] for k in ref_gpu.files if k.startswith("grid.")})
    assert names
    for name in names:
        D, L, C, base, log2T, desired, gridtype, align, interp, is_half = ref_gpu[f"grid.{name}.meta"].tolist()
        S = float(ref_gpu[f"grid.{name}.S"])
        offs, x, table = ref_gpu[f"grid.{name}.offsets"], ref_gpu[f"grid.{name}.x"], ref_gpu[f"grid.{name}.table"]
        B = x.shape[0]
        out, jac = grid_np.grid_encode_forward(x, table.astype(np.float32), offs, S, base, gridtype, bool(align),
                                               interp, want_dy_dx=True)
        exp = ref_gpu[f"grid.{name}.out_LBC"].astype(np.float32).transpose(1, 0, 2).reshape(B, L * C)
        tol = 1e-2 if is_half else 1e-5
        np.testing.assert_allclose(out, exp, rtol=tol, atol=tol if is_half else 2e-6, err_msg=name)
        if not is_half:
            ej = ref_gpu[f"grid.{name}.dy_dx"]
            np.testing.assert_allclose(jac, ej, rtol=1e-3, atol=1e-5 * max(1.0, np.abs(ej).max()), err_msg=name)
            grad = ref_gpu[f"grid.{name}.grad_LBC"].transpose(1, 0, 2).reshape(B, L * C)
            gt = grid_np.grid_encode_backward(grad, x, offs, table.shape[0], C, S, base, gridtype, bool(align), interp)
            eg = ref_gpu[f"grid.{name}.grad_table"]
            np.testing.assert_allclose(gt, eg, rtol=1e-4, atol=1e-4 * np.abs(eg).max(), err_msg=name)
            gi = grid_np.grid_input_backward(grad, ej, B, D, C, L)
            egi = ref_gpu[f"grid.{name}.grad_inputs"]
            np.testing.assert_allclose(gi, egi, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(egi).max()), err_msg=name)
        if f"grid.{name}.tv" in ref_gpu.files:
            tv = grid_np.grad_total_variation(x, table, offs, 0.37, S, base, gridtype, bool(align))
            etv = ref_gpu[f"grid.{name}.tv"]
            np.testing.assert_allclose(tv, etv, rtol=1e-3, atol=1e-4 * np.abs(etv).max(), err_msg=name)
        if f"grid.{name}.wd" in ref_gpu.files:
            np.testing.assert_allclose(grid_np.grad_weight_decay(table, offs, 0.1), ref_gpu[f"grid.{name}.wd"],
                                       rtol=1e-5, atol=1e-9, err_msg=name)


def test_level_resolutions_pinned_by_reference_probe(ref_gpu):
    """SURVEY §8 c7 probe tables run through the REFERENCE kernel on the GPU: the checksum the reference
    produced must equal sum_k w_k * probe[row_k] with the oracle's rows — pins resolution (device exp2f),
    dense-vs-hash, hash and modulo for every table shape the reference builds."""
    names = sorted({k.split(".")[1] for k in ref_gpu.files if k.startswith("probe.")})
    assert names
    for name in names:
        x, offs = ref_gpu[f"probe.{name}.x"], ref_gpu[f"probe.{name}.offsets"]
        S = float(ref_gpu[f"probe.{name}.S"])
        exp = ref_gpu[f"probe.{name}.out_LBC"].astype(np.float64)
        rows, geom = grid_np.dump_indices(x, offs, S, 16, 0, False)
        rows = rows.astype(np.int64)
        for l in range(len(offs) - 1):
            _, frac, _ = grid_np.locate(x, int(geom[l, 0]), False, 0)
            w = grid_np.corner_weights(frac).astype(np.float64)
            for ch, p in enumerate((251, 241)[:exp.shape[2]]):
                err = np.abs((w * (rows[:, l] % p)).sum(1) - exp[l, :, ch])
                assert err.max() < 2e-3, (name, l, err.max())
