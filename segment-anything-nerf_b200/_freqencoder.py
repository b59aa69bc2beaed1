"""Drop-in for the reference's compiled ``_freqencoder`` pybind module
(``freqencoder/src/bindings.cpp:5-8``, ``freqencoder.h:7,10``): same names and positional
signatures, implemented on ``libsanerf_b200.so``."""
import torch

from sanerf_b200 import _lib


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float32 tensor")  # reference reads data_ptr<float>() (freqencoder.cu:109)


def freq_encode_forward(inputs, B, D, deg, C, outputs):
    _chk(inputs, "inputs"); _chk(outputs, "outputs")
    lib = _lib.load()
    with torch.cuda.device(inputs.device):
        rc = lib.sanerf_freq_encode_forward(inputs.data_ptr(), int(B), int(D), int(deg), int(C),
                                            outputs.data_ptr(), _lib.current_stream(inputs.device))
    _lib.check(rc, "freq_encode_forward")


def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
    _chk(grad, "grad"); _chk(outputs, "outputs"); _chk(grad_inputs, "grad_inputs")
    lib = _lib.load()
    with torch.cuda.device(grad.device):
        rc = lib.sanerf_freq_encode_backward(grad.data_ptr(), outputs.data_ptr(), int(B), int(D), int(deg),
                                             int(C), grad_inputs.data_ptr(), _lib.current_stream(grad.device))
    _lib.check(rc, "freq_encode_backward")
