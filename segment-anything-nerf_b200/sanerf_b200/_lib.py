"""ctypes binding of ``libsanerf_b200.so`` (C ABI declared in ``include/sanerf_b200.h``).

There is deliberately NO fallback: if the shared library is missing, or a call returns a
non-zero status, a ``RuntimeError`` is raised.  The product path never routes through
``oracle/`` or any CPU implementation.
"""
from __future__ import annotations

import ctypes
import os
import re

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# SANERF_LIB_PATH: diagnostic builds of the same library (e.g. -DSANERF_HEAD_TRACE); never a different implementation
_LIB_PATH = os.environ.get("SANERF_LIB_PATH") or os.path.join(_PKG_DIR, "lib", "libsanerf_b200.so")
_HEADER = os.path.join(os.path.dirname(_PKG_DIR), "include", "sanerf_b200.h")

SANERF_F32, SANERF_F16 = 0, 1
LAYOUT_LBC, LAYOUT_BLC = 0, 1
ABI_VERSION = 27

c_void_p, c_int, c_u32, c_u64, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                                          ctypes.c_uint64, ctypes.c_float)

# name -> argtypes; restype is int unless listed in _RESTYPES
_SIGNATURES = {
    "sanerf_abi_version": [],
    "sanerf_last_error": [],
    "sanerf_status_string": [c_int],
    "sanerf_grid_encode_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32,
                                   c_u32, c_float, c_u32, c_void_p, c_u32, c_int, c_u32, c_int, c_int,
                                   c_int, c_void_p],
    "sanerf_grid_encode_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32,
                                    c_u32, c_u32, c_float, c_u32, c_void_p, c_void_p, c_u32, c_int, c_u32,
                                    c_int, c_int, c_void_p],
    "sanerf_grad_total_variation": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_u32, c_u32, c_u32,
                                    c_u32, c_float, c_u32, c_u32, c_int, c_int, c_void_p],
    "sanerf_grad_weight_decay": [c_void_p, c_void_p, c_void_p, c_float, c_u32, c_u32, c_u32, c_int, c_void_p],
    "sanerf_grid_dump_indices": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_float, c_u32,
                                 c_u32, c_int, c_void_p],
    "sanerf_ray_features_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32, c_float, c_u32,
                                    c_void_p, c_u32, c_void_p],
    "sanerf_ray_features_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32, c_float, c_u32,
                                     c_void_p, c_u32, c_u32, c_u32, c_void_p],
    "sanerf_sam_pack": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_int, c_void_p, c_u32, c_void_p],
    "sanerf_sh_encode_forward": [c_void_p, c_void_p, c_u32, c_u32, c_u32, c_void_p, c_u32, c_void_p],
    "sanerf_sh_encode_backward": [c_void_p, c_void_p, c_u32, c_u32, c_u32, c_void_p, c_void_p, c_void_p],
    "sanerf_freq_encode_forward": [c_void_p, c_u32, c_u32, c_u32, c_u32, c_void_p, c_void_p],
    "sanerf_freq_encode_backward": [c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32, c_void_p, c_void_p],
    "sanerf_trunc_exp_forward": [c_void_p, c_void_p, c_u64, c_u32, c_u32, c_void_p],
    "sanerf_trunc_exp_backward": [c_void_p, c_void_p, c_void_p, c_u64, c_u32, c_u32, c_void_p],
    "sanerf_composite_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_void_p, c_u32, c_u32, c_u32,
                                 c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_composite_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_void_p, c_u32, c_u32, c_u32,
                                  c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_u32, c_void_p],
    "sanerf_head_composite_forward": [c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_int, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_head_composite_backward": [c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_int, c_float, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_generate_rays": [c_void_p, c_u32, c_void_p, c_u32, c_void_p, c_u32, c_u32, c_void_p, c_void_p, c_void_p],
    "sanerf_sample_uniform": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_u32, c_void_p, c_u32, c_u32, c_int,
                              c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_sample_pdf": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_u32, c_void_p, c_void_p, c_u32, c_void_p,
                          c_u32, c_u32, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_void_p, c_void_p],
    "sanerf_prop_density_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_float, c_u32,
                                    c_void_p, c_void_p, c_void_p],
    "sanerf_prop_density_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_float, c_u32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_proposal_loss": [c_void_p, c_void_p, c_u32, c_void_p, c_void_p, c_u32, c_u32, c_float, c_void_p, c_void_p,
                             c_void_p],
    "sanerf_distortion_loss": [c_void_p, c_void_p, c_u32, c_u32, c_float, c_void_p, c_void_p, c_void_p],
    "sanerf_view_head": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_u32,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_gemm_tc": [c_void_p, c_u32, c_int, c_void_p, c_u32, c_int, c_void_p, c_u32, c_u32, c_u32, c_u32, c_u32, c_int,
                       c_void_p, c_int, c_float, c_void_p, c_u32, c_u32, c_void_p, c_int, c_void_p],
    "sanerf_copy_rows": [c_void_p, c_u32, c_void_p, c_u32, c_u32, c_u32, c_void_p],
    "sanerf_colsum_add": [c_void_p, c_u32, c_u32, c_u32, c_void_p, c_void_p],
    "sanerf_layernorm_mse": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_u64, c_u64, c_u32, c_u32, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_adam_schedule": [c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_float, c_void_p],
    "sanerf_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_u64, c_void_p, c_float, c_float, c_float,
                         c_float, c_int, c_void_p, c_void_p, c_void_p],
    "sanerf_ema_update": [c_void_p, c_void_p, c_u64, c_float, c_void_p],
    "sanerf_adam_step_half": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u64, c_void_p, c_float, c_float, c_float,
                              c_float, c_int, c_void_p],
    "sanerf_symm_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_u64, c_u64, c_u32, c_u32, c_void_p, c_float, c_float, c_float, c_float,
                              c_void_p, c_u32, c_u32, c_u32, c_void_p],
    "sanerf_uniform_fill": [c_void_p, c_u64, c_u64, c_void_p, c_void_p, c_u32, c_void_p],
    "sanerf_field_head_forward": [c_void_p, c_void_p, c_void_p, c_float, c_u32, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_u32, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "sanerf_field_head_forward_chunk": [c_void_p, c_void_p, c_void_p, c_float, c_u32, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32, c_void_p],
    "sanerf_head_composite_chunk": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32, c_u32, c_u32, c_u32, c_int,
                                    c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "sanerf_field_head_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_u32,
                                   c_void_p, c_void_p, c_void_p, c_float, c_u32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p],
    "sanerf_umma_selftest": [c_int, c_u32, c_u32, c_u32, c_void_p, c_void_p, c_void_p, c_void_p],
}
_RESTYPES = {"sanerf_last_error": ctypes.c_char_p, "sanerf_status_string": ctypes.c_char_p}

_lib = None


def header_symbols(header: str = _HEADER):
    """Names of every function ``include/sanerf_b200.h`` declares (used by the CPU tests)."""
    with open(header) as f:
        return sorted(set(re.findall(r"SANERF_API\s+[\w\s\*]+?\b(sanerf_\w+)\s*\(", f.read())))


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and type every entry point.  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C segment-anything-nerf_b200/csrc`. There is no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library is stale
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    got = lib.sanerf_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libsanerf_b200.so has ABI {got}, Python binding expects {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    """Turn a non-zero C status into the exception the reference's TORCH_CHECK would raise."""
    if status != 0:
        lib = load()
        msg = lib.sanerf_last_error().decode(errors="replace")
        kind = lib.sanerf_status_string(status).decode()
        raise RuntimeError(f"{what}: {kind}: {msg}")


class LaunchStats:
    """Counts C-ABI kernel launches and, when asked, brackets the launches of ONE named entry point with CUDA
    events on the launching stream (bench.py uses this for the roofline of the dominant kernel)."""

    def __init__(self):
        self.count = 0
        self.watch = None          # (name, predicate) or None
        self.spans = []            # [(start_event, end_event, info)]

    def reset(self, watch=None, predicate=None):
        self.count = 0
        self.watch = (watch, predicate) if watch else None
        self.spans = []

    def span(self, name, **info):
        return _Span(self, name, info)

    def durations_ms(self):
        return [(s.elapsed_time(e), info) for s, e, info in self.spans]


class _Span:
    def __init__(self, stats, name, info):
        self.stats, self.name, self.info, self.ev = stats, name, info, None

    def __enter__(self):
        st = self.stats
        st.count += 1
        if st.watch and st.watch[0] in (self.name, "*") and (st.watch[1] is None or st.watch[1](self.info)):
            import torch

            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            self.stats.spans.append((self.ev[0], self.ev[1], dict(self.info, name=self.name)))
        return False


stats = LaunchStats()


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else t.data_ptr()


def current_stream(device=None):
    import torch

    return torch.cuda.current_stream(device).cuda_stream
