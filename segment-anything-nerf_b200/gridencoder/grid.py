"""``GridEncoder`` — multiresolution hash / tiled grid encoding on sm_100a.

Host-side mirror of the reference operator (``gridencoder/grid.py`` of
lyclyc52/Segment-Anything-NeRF): same constructor arguments, attributes, ``forward`` keyword
arguments, ``state_dict`` keys (``embeddings`` [rows, C] fp32, ``offsets`` int32 [L+1]) and
autograd/AMP behaviour, so reference checkpoints load and ``nerf/network.py`` runs unchanged.

What differs is below the surface: the kernels are called through the C ABI
(``include/sanerf_b200.h``) and read/write the ``[B, L*C]`` layout directly, so the
``[L,B,C]`` staging tensor and the two permute copies of the reference (grid.py:63, :80) do
not exist here.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from sanerf_b200 import _lib

GRIDTYPE_IDS = {"hash": 0, "tiled": 1}
INTERP_IDS = {"linear": 0, "smoothstep": 1}


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size):
    """Row offsets of each level's table slice (reference: grid.py:124-134, host fp64 math)."""
    cap = 2 ** log2_hashmap_size
    edges = [0]
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        rows = min(cap, res ** input_dim)
        rows = int(np.ceil(rows / 8) * 8)
        edges.append(edges[-1] + rows)
    return np.asarray(edges, dtype=np.int32)


def _dtype_id(t):
    if t.dtype == torch.float32:
        return _lib.SANERF_F32
    if t.dtype == torch.float16:
        return _lib.SANERF_F16
    raise RuntimeError(f"GridEncoder: unsupported table dtype {t.dtype}")


class _GridEncode(Function):
    """Autograd node; argument order follows the reference's ``_grid_encode`` (grid.py:24-95)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0, max_level=None):
        if not inputs.is_cuda:
            raise RuntimeError("inputs must be a CUDA tensor")
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()  # coordinates always fp32 (gridencoder.cu:135)
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        max_level = L if max_level is None else min(int(max_level), L)

        table = embeddings
        if torch.is_autocast_enabled() and C % 2 == 0:  # same rule as grid.py:43-46
            table = embeddings.to(torch.half)
        table = table.contiguous()

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=table.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None
        lib = _lib.load()
        with torch.cuda.device(inputs.device), _lib.stats.span("grid_encode_forward", B=B, L=L, C=C, D=D,
                                                               half=table.dtype == torch.float16):
            rc = lib.sanerf_grid_encode_forward(
                inputs.data_ptr(), table.data_ptr(), offsets.data_ptr(), outputs.data_ptr(), B, D, C, L,
                max_level, S, H, _lib.ptr(dy_dx), int(gridtype), int(bool(align_corners)), int(interpolation),
                _dtype_id(table), _lib.LAYOUT_BLC, 1, _lib.current_stream(inputs.device))
        _lib.check(rc, "grid_encode_forward")

        ctx.save_for_backward(inputs, table, offsets, dy_dx)
        ctx.meta = (B, D, C, L, S, H, int(gridtype), int(interpolation), max_level, bool(align_corners))
        ctx.param_dtype = embeddings.dtype
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, table, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, max_level, align_corners = ctx.meta
        if not ctx.needs_input_grad[1] and dy_dx is None:
            return (None,) * 10                     # frozen table (stage 2 freezes the stage-1 grids)
        grad = grad.contiguous()
        if grad.dtype != table.dtype:
            grad = grad.to(table.dtype)
        grad_table = torch.zeros_like(table)
        grad_inputs = torch.empty(B, D, device=inputs.device, dtype=table.dtype) if dy_dx is not None else None
        lib = _lib.load()
        with torch.cuda.device(inputs.device), _lib.stats.span("grid_encode_backward", B=B, L=L, C=C, D=D,
                                                               half=table.dtype == torch.float16):
            rc = lib.sanerf_grid_encode_backward(
                grad.data_ptr(), inputs.data_ptr(), table.data_ptr(), offsets.data_ptr(), grad_table.data_ptr(),
                B, D, C, L, max_level, S, H, _lib.ptr(dy_dx), _lib.ptr(grad_inputs), gridtype, int(align_corners),
                interpolation, _dtype_id(table), _lib.LAYOUT_BLC, _lib.current_stream(inputs.device))
        _lib.check(rc, "grid_encode_backward")
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        if grad_table.dtype != ctx.param_dtype:
            grad_table = grad_table.to(ctx.param_dtype)
        return grad_inputs, grad_table, None, None, None, None, None, None, None, None


grid_encode = _GridEncode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:
            # finest level hits `desired_resolution` (grid.py:107-108; fp64 on the host)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = GRIDTYPE_IDS[gridtype]
        self.interpolation = interpolation
        self.interp_id = INTERP_IDS[interpolation]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size

        edges = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer("offsets", torch.from_numpy(edges))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(edges[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)  # grid.py:144-146

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1, max_level=None):
        """inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]."""
        unit = (inputs + bound) / (2 * bound)  # same fp32 expression as grid.py:156
        lead = list(unit.shape[:-1])
        flat = unit.view(-1, self.input_dim)
        out = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          flat.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, max_level)
        return out.view(lead + [self.output_dim])

    def _geometry(self):
        C = self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        return self.input_dim, C, L, float(np.log2(self.per_level_scale)), int(self.base_resolution)

    @torch.amp.autocast("cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """Adds the TV gradient to ``embeddings.grad`` in place (grid.py:170-192)."""
        D, C, L, S, H = self._geometry()
        if inputs is None:
            inputs = torch.rand(B, D, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, D)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        inputs = inputs.contiguous().to(self.embeddings.dtype)
        lib = _lib.load()
        with torch.cuda.device(self.embeddings.device):
            rc = lib.sanerf_grad_total_variation(
                inputs.data_ptr(), self.embeddings.data_ptr(), self.embeddings.grad.data_ptr(),
                self.offsets.data_ptr(), float(weight), int(B), D, C, L, S, H, self.gridtype_id,
                int(self.align_corners), _dtype_id(self.embeddings), _lib.current_stream(self.embeddings.device))
        _lib.check(rc, "grad_total_variation")

    @torch.amp.autocast("cuda", enabled=False)
    def grad_weight_decay(self, weight=0.1):
        """Adds the level-mean weight-decay gradient to ``embeddings.grad`` in place (grid.py:194-205)."""
        rows, C = self.embeddings.shape
        L = self.offsets.shape[0] - 1
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        lib = _lib.load()
        with torch.cuda.device(self.embeddings.device):
            rc = lib.sanerf_grad_weight_decay(
                self.embeddings.data_ptr(), self.embeddings.grad.data_ptr(), self.offsets.data_ptr(),
                float(weight), int(rows), int(C), int(L), _dtype_id(self.embeddings),
                _lib.current_stream(self.embeddings.device))
        _lib.check(rc, "grad_weight_decay")
