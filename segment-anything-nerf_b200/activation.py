"""``trunc_exp`` — density activation with a clamped-exponent backward (reference: activation.py:5-18).

forward  y = exp(x) in fp32;  backward  dx = g * exp(clamp(x, -15, 15)).
One sm_100a kernel each way (the reference runs exp / clamp / exp / mul torch kernels).
"""
import torch
from torch.autograd import Function

from sanerf_b200 import _lib


class _TruncExp(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        if not x.is_cuda:
            raise RuntimeError("trunc_exp: x must be a CUDA tensor (no CPU fallback)")
        xc = x.contiguous()
        y = torch.empty_like(xc)
        lib = _lib.load()
        with torch.cuda.device(xc.device), _lib.stats.span("trunc_exp_forward", n=xc.numel()):
            rc = lib.sanerf_trunc_exp_forward(xc.data_ptr(), y.data_ptr(), xc.numel(), 1, 0,
                                              _lib.current_stream(xc.device))
        _lib.check(rc, "trunc_exp_forward")
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        lib = _lib.load()
        with torch.cuda.device(x.device), _lib.stats.span("trunc_exp_backward", n=x.numel()):
            rc = lib.sanerf_trunc_exp_backward(g.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), 1, 0,
                                               _lib.current_stream(x.device))
        _lib.check(rc, "trunc_exp_backward")
        return dx


trunc_exp = _TruncExp.apply
