"""Drop-in for the reference's compiled ``_gridencoder`` pybind module.

Same four function names, positional signatures, argument meaning and error behaviour as
``gridencoder/src/bindings.cpp:5-10`` / ``gridencoder.h:12-16`` of the reference, so that the
reference's own ``gridencoder/grid.py`` (``import _gridencoder as _backend``, grid.py:9-12)
runs unmodified on top of the sm_100a kernels.  Tensors are caller-allocated; nothing is
returned.  Implemented as a thin ctypes hop into ``libsanerf_b200.so`` (no CPU fallback).
"""
import torch

from sanerf_b200 import _lib

_FLOATING = (torch.float32, torch.float16, torch.float64)


def _chk(t, name, *, integer=False):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if integer:
        if t.dtype != torch.int32:
            raise RuntimeError(f"{name} must be an int tensor")
    elif t.dtype not in _FLOATING:
        raise RuntimeError(f"{name} must be a floating tensor")


def _dtype_id(t, what):
    if t.dtype == torch.float32:
        return _lib.SANERF_F32
    if t.dtype == torch.float16:
        return _lib.SANERF_F16
    # the reference also instantiates double; it is never used (fp64 tables) and not carried over
    raise RuntimeError(f"{what}: only float32 and float16 tables are supported on sm_100a")


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx,
                        gridtype, align_corners, interp):
    """outputs[L,B,C] <- encode(inputs[B,D]); optional dy_dx[B, L*D*C] (gridencoder.cu:467-490)."""
    _chk(inputs, "inputs"); _chk(embeddings, "embeddings"); _chk(offsets, "offsets", integer=True)
    _chk(outputs, "outputs")
    if inputs.dtype != torch.float32:
        raise RuntimeError("inputs must be float32")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_grid_encode_forward(
            inputs.data_ptr(), embeddings.data_ptr(), offsets.data_ptr(), outputs.data_ptr(),
            int(B), int(D), int(C), int(L), int(max_level), float(S), int(H), _lib.ptr(dy_dx),
            int(gridtype), int(bool(align_corners)), int(interp), _dtype_id(embeddings, "grid_encode_forward"),
            _lib.LAYOUT_LBC, 0, _lib.current_stream(inputs.device))
    _lib.check(rc, "grid_encode_forward")


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H,
                         dy_dx, grad_inputs, gridtype, align_corners, interp):
    """grad_embeddings += scatter(grad[L,B,C]); optional grad_inputs (gridencoder.cu:492-522)."""
    _chk(grad, "grad"); _chk(inputs, "inputs"); _chk(embeddings, "embeddings")
    _chk(offsets, "offsets", integer=True); _chk(grad_embeddings, "grad_embeddings")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_grid_encode_backward(
            grad.data_ptr(), inputs.data_ptr(), embeddings.data_ptr(), offsets.data_ptr(),
            grad_embeddings.data_ptr(), int(B), int(D), int(C), int(L), int(max_level), float(S), int(H),
            _lib.ptr(dy_dx), _lib.ptr(grad_inputs) if dy_dx is not None else None, int(gridtype),
            int(bool(align_corners)), int(interp), _dtype_id(grad, "grid_encode_backward"),
            _lib.LAYOUT_LBC, _lib.current_stream(inputs.device))
    _lib.check(rc, "grid_encode_backward")


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    """In-place TV gradient on ``grad`` (gridencoder.cu:662-668)."""
    lib = _lib.load()
    with torch.cuda.device(embeddings.device):
        rc = lib.sanerf_grad_total_variation(
            inputs.data_ptr(), embeddings.data_ptr(), grad.data_ptr(), offsets.data_ptr(), float(weight),
            int(B), int(D), int(C), int(L), float(S), int(H), int(gridtype), int(bool(align_corners)),
            _dtype_id(embeddings, "grad_total_variation"), _lib.current_stream(embeddings.device))
    _lib.check(rc, "grad_total_variation")


def grad_weight_decay(embeddings, grad, offsets, weight, B, C, L):
    """In-place level-mean weight decay on ``grad`` (gridencoder.cu:705-713)."""
    lib = _lib.load()
    with torch.cuda.device(embeddings.device):
        rc = lib.sanerf_grad_weight_decay(
            embeddings.data_ptr(), grad.data_ptr(), offsets.data_ptr(), float(weight), int(B), int(C), int(L),
            _dtype_id(embeddings, "grad_weight_decay"), _lib.current_stream(embeddings.device))
    _lib.check(rc, "grad_weight_decay")
