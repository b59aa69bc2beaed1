"""C-ABI surface checks that need no GPU: the library loads, exports exactly what
include/sanerf_b200.h declares, the ctypes table matches the header's arity, and the reference-
compatible shims reject CPU tensors the way the reference's CHECK_CUDA does."""
import ctypes
import os
import re

import pytest
import torch

from sanerf_b200 import _lib


def _header_decls():
    text = open(_lib._HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"SANERF_API\s+[\w\s\*]+?\b(sanerf_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_library_exists_and_loads():
    assert os.path.exists(_lib.lib_path()), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.sanerf_abi_version() == _lib.ABI_VERSION


def test_every_header_symbol_is_exported_and_typed():
    decls = _header_decls()
    assert len(decls) >= 16
    raw = ctypes.CDLL(_lib.lib_path())
    for name, n_args in decls.items():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib._SIGNATURES[name]) == n_args, f"{name}: ctypes arity != header arity"
    assert set(_lib._SIGNATURES) == set(decls)
    assert set(_lib.header_symbols()) == set(decls)


def test_status_strings():
    lib = _lib.load()
    assert lib.sanerf_status_string(0) == b"ok"
    assert lib.sanerf_status_string(1) == b"invalid argument"


def test_invalid_arguments_are_reported_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    lib = _lib.load()
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    rc = lib.sanerf_grid_encode_forward(one, one, one, one, 4, 7, 2, 1, 1, 0.0, 16, None, 0, 0, 0, 0, 0, 0, None)
    assert rc == 1 and b"D must be 2, 3, 4 or 5" in lib.sanerf_last_error()
    rc = lib.sanerf_grid_encode_forward(one, one, one, one, 4, 3, 3, 1, 1, 0.0, 16, None, 0, 0, 0, 0, 0, 0, None)
    assert rc == 1 and b"C must be 1, 2, 4, 8, 16 or 32" in lib.sanerf_last_error()
    rc = lib.sanerf_grid_encode_forward(None, one, one, one, 4, 3, 2, 1, 1, 0.0, 16, None, 0, 0, 0, 0, 0, 0, None)
    assert rc == 2
    rc = lib.sanerf_sh_encode_forward(one, one, 4, 3, 9, None, 0, None)
    assert rc == 1 and b"degree in [1, 8]" in lib.sanerf_last_error()
    rc = lib.sanerf_freq_encode_forward(one, 4, 3, 6, 38, one, None)
    assert rc == 1
    with pytest.raises(RuntimeError, match="invalid argument"):
        _lib.check(rc, "freq_encode_forward")


def test_shims_reject_cpu_tensors_like_check_cuda():
    import _freqencoder
    import _gridencoder
    import _shencoder

    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _gridencoder.grid_encode_forward(x, x, torch.zeros(2, dtype=torch.int32), x, 4, 3, 2, 1, 1, 0.0, 16, None,
                                         0, False, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _shencoder.sh_encode_forward(x, x, 4, 3, 4, None)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _freqencoder.freq_encode_forward(x, 4, 3, 6, 39, x)
    assert {"grid_encode_forward", "grid_encode_backward", "grad_total_variation",
            "grad_weight_decay"} <= set(dir(_gridencoder))
    assert {"sh_encode_forward", "sh_encode_backward"} <= set(dir(_shencoder))
    assert {"freq_encode_forward", "freq_encode_backward"} <= set(dir(_freqencoder))


def test_operator_surface_matches_reference_names():
    """Constructor arguments / attributes the reference's network.py relies on."""
    from encoding import get_encoder
    from gridencoder import GridEncoder

    enc = GridEncoder(input_dim=3, num_levels=4, level_dim=2, base_resolution=4, log2_hashmap_size=8,
                      desired_resolution=32)
    assert enc.output_dim == 8 and enc.embeddings.shape[1] == 2 and enc.offsets.dtype == torch.int32
    assert set(enc.state_dict().keys()) == {"embeddings", "offsets"}
    assert float(enc.embeddings.abs().max()) <= 1e-4
    for attr in ("per_level_scale", "base_resolution", "n_params", "gridtype_id", "interp_id", "align_corners"):
        assert hasattr(enc, attr)
    sh, n = get_encoder("sh", degree=4)
    assert n == 16
    fr, n = get_encoder("frequency", multires=6)
    assert n == 39
    ft, n = get_encoder("frequency_torch", multires=6)
    assert n == 39
    hg, n = get_encoder("hashgrid", num_levels=4, level_dim=2, base_resolution=4, log2_hashmap_size=8,
                        desired_resolution=32)
    assert n == 8
    with pytest.raises(NotImplementedError):
        get_encoder("nope")
    with pytest.raises(ValueError, match="grad is None"):
        enc.grad_weight_decay(0.1)


def test_ops_fail_loudly_without_cuda():
    """No CPU fallback anywhere on the product path."""
    from activation import trunc_exp
    from gridencoder import GridEncoder
    from sanerf_b200.ops import composite

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    enc = GridEncoder(num_levels=2, level_dim=2, base_resolution=4, log2_hashmap_size=6, desired_resolution=8)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(3, 3))
    with pytest.raises(RuntimeError):
        trunc_exp(torch.zeros(3))
    with pytest.raises(RuntimeError):
        composite(torch.ones(2, 4), torch.ones(2, 4), torch.ones(2, 4))


def test_stage2_ops_fail_loudly_without_cuda():
    """The stage-2 operators (ray features, tensor-core GEMM, LayerNorm + MSE) have no CPU path either, and the
    SAM head falls back to nothing: SkipConnMLP on CPU tensors runs the plain torch layers of the reference."""
    from gridencoder import GridEncoder
    from nerf.network import SkipConnMLP
    from sanerf_b200 import fused

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    enc = GridEncoder(num_levels=2, level_dim=8, base_resolution=4, log2_hashmap_size=6, desired_resolution=8)
    with pytest.raises(RuntimeError):
        fused.ray_features(torch.zeros(2, 4, 3), torch.ones(2, 4), enc)
    a, b, c = torch.zeros(8, 8), torch.zeros(8, 8), torch.zeros(8, 8)
    with pytest.raises(RuntimeError):
        fused.gemm_tc(a, b, c, 8, 8, 8)
    mlp = SkipConnMLP(6, 4, 8, 3, skip_layers=[1])
    mlp.tc = True
    x = torch.randn(5, 6)
    assert not fused.skip_mlp_supported(mlp, x)            # CPU tensor: the torch layers run (module semantics of the reference)
    y = mlp(x)
    h = torch.nn.functional.leaky_relu(mlp.net[0](x))
    h = torch.nn.functional.leaky_relu(mlp.net[1](torch.cat([h, x], dim=-1)))
    torch.testing.assert_close(y, mlp.net[2](h))


def test_update_stream_priority_policy(monkeypatch):
    from sanerf_b200.step import _update_priority

    monkeypatch.delenv("SANERF_UPD_PRIO", raising=False)
    assert _update_priority(1) == 0 and _update_priority(2) == -1 and _update_priority(8) == -1
    monkeypatch.setenv("SANERF_UPD_PRIO", "0")
    assert _update_priority(8) == 0
