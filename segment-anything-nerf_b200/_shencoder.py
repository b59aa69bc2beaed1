"""Drop-in for the reference's compiled ``_shencoder`` pybind module
(``shencoder/src/bindings.cpp:5-8``, ``shencoder.h:9-10``): same names and positional
signatures, implemented on ``libsanerf_b200.so``."""
import torch

from sanerf_b200 import _lib


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        # the Python wrapper forces fp32 (sphere_harmonics.py:16); half/double are not carried over
        raise RuntimeError(f"{name} must be a float32 tensor")


def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
    """outputs[B, C*C] <- SH basis of unit vectors inputs[B,3]; C is the degree (shencoder.cu:400-417)."""
    _chk(inputs, "inputs"); _chk(outputs, "outputs")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_sh_encode_forward(inputs.data_ptr(), outputs.data_ptr(), int(B), int(D), int(C),
                                          _lib.ptr(dy_dx), 0, _lib.current_stream(inputs.device))
    _lib.check(rc, "sh_encode_forward")


def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    """grad_inputs += grad . dy_dx (shencoder.cu:419-439)."""
    _chk(grad, "grad"); _chk(inputs, "inputs"); _chk(dy_dx, "dy_dx"); _chk(grad_inputs, "grad_inputs")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_sh_encode_backward(grad.data_ptr(), inputs.data_ptr(), int(B), int(D), int(C),
                                           dy_dx.data_ptr(), grad_inputs.data_ptr(),
                                           _lib.current_stream(inputs.device))
    _lib.check(rc, "sh_encode_backward")
