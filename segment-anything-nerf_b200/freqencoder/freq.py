"""``FreqEncoder`` — NeRF positional encoding on sm_100a.

Mirrors the reference operator (``freqencoder/freq.py``): ``FreqEncoder(input_dim=3, degree=4)``,
``output_dim = input_dim * (1 + 2*degree)``, column order ``[x, sin 2^0 x, cos 2^0 x, ...]``,
fp32 forced, backward from the saved outputs.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from sanerf_b200 import _lib


class _FreqEncode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        if not inputs.is_cuda:
            inputs = inputs.cuda()  # freq.py:22
        inputs = inputs.contiguous()
        B, D = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        lib = _lib.load()
        with torch.cuda.device(inputs.device):
            rc = lib.sanerf_freq_encode_forward(inputs.data_ptr(), B, D, int(degree), int(output_dim),
                                                outputs.data_ptr(), _lib.current_stream(inputs.device))
        _lib.check(rc, "freq_encode_forward")
        ctx.save_for_backward(outputs)
        ctx.meta = (B, D, int(degree), int(output_dim))
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        (outputs,) = ctx.saved_tensors
        B, D, degree, output_dim = ctx.meta
        grad = grad.contiguous()
        grad_inputs = torch.empty(B, D, dtype=outputs.dtype, device=outputs.device)
        lib = _lib.load()
        with torch.cuda.device(outputs.device):
            rc = lib.sanerf_freq_encode_backward(grad.data_ptr(), outputs.data_ptr(), B, D, degree, output_dim,
                                                 grad_inputs.data_ptr(), _lib.current_stream(outputs.device))
        _lib.check(rc, "freq_encode_backward")
        return grad_inputs, None, None


freq_encode = _FreqEncode.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        lead = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        out = freq_encode(flat, self.degree, self.output_dim)
        return out.reshape(lead + [self.output_dim])
