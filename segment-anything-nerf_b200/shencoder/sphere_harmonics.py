"""``SHEncoder`` — real spherical-harmonics direction encoding (degree 1..8) on sm_100a.

Mirrors the reference operator (``shencoder/sphere_harmonics.py``): ``SHEncoder(input_dim=3,
degree=4)``, ``forward(inputs, size=1)``, ``output_dim = degree**2``, fp32 forced, gradient
w.r.t. the direction only when it requires grad.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from sanerf_b200 import _lib


class _SHEncode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        if not inputs.is_cuda:
            raise RuntimeError("inputs must be a CUDA tensor")
        inputs = inputs.contiguous()
        B, D = inputs.shape
        n_out = degree * degree
        outputs = torch.empty(B, n_out, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, D * n_out, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
        lib = _lib.load()
        with torch.cuda.device(inputs.device), _lib.stats.span("sh_encode_forward", B=B):
            rc = lib.sanerf_sh_encode_forward(inputs.data_ptr(), outputs.data_ptr(), B, D, int(degree),
                                              _lib.ptr(dy_dx), 0, _lib.current_stream(inputs.device))
        _lib.check(rc, "sh_encode_forward")
        ctx.save_for_backward(inputs, dy_dx)
        ctx.meta = (B, D, int(degree))
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, dy_dx = ctx.saved_tensors
        if dy_dx is None:
            return None, None, None
        B, D, degree = ctx.meta
        grad = grad.contiguous()
        grad_inputs = torch.zeros_like(inputs)
        lib = _lib.load()
        with torch.cuda.device(inputs.device):
            rc = lib.sanerf_sh_encode_backward(grad.data_ptr(), inputs.data_ptr(), B, D, degree, dy_dx.data_ptr(),
                                               grad_inputs.data_ptr(), _lib.current_stream(inputs.device))
        _lib.check(rc, "sh_encode_backward")
        return grad_inputs, None, None


sh_encode = _SHEncode.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        """inputs [..., 3] in [-size, size] -> [..., degree**2]; direction is normalised first."""
        d = inputs / size
        d = d / torch.norm(d, dim=-1, keepdim=True)  # sphere_harmonics.py:79-82
        lead = list(d.shape[:-1])
        flat = d.reshape(-1, self.input_dim)
        out = sh_encode(flat, self.degree, flat.requires_grad)
        return out.reshape(lead + [self.output_dim])
