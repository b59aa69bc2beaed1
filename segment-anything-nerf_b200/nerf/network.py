"""``NeRFNetwork`` — the field queried by the renderer (reference: ``nerf/network.py:94-308``).

Same sub-module names, shapes and ``state_dict`` keys as the reference (``grid``, ``grid_mlp.net.{0,1,2}``,
``view_mlp.net.*``, ``prop_encoders.{0,1}``, ``prop_mlp.{0,1}.net.*``, ``s_grid``, ``samvit_mlp.0.net.*``,
``samvit_mlp.1``) so reference checkpoints load.  Mask heads (stage 3) are out of scope.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from activation import trunc_exp
from encoding import get_encoder
from gridencoder.grid import grid_encode
from sanerf_b200 import fused

from .renderer import NeRFRenderer


class MLP(nn.Module):
    """ReLU MLP (network.py:9-34).  ``save_intermedian_results`` keeps the reference's (misspelt) hook."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        widths = [dim_in] + [dim_hidden] * (num_layers - 1) + [dim_out]
        self.net = nn.ModuleList(nn.Linear(widths[i], widths[i + 1], bias=bias) for i in range(num_layers))

    def forward(self, x, save_intermedian_results=False):
        keep = [] if save_intermedian_results else None
        for i, layer in enumerate(self.net):
            x = layer(x)
            if i + 1 < self.num_layers:
                x = F.relu(x)
            if keep is not None:
                keep.append(x.detach())
        if keep is not None:
            self.intermedian_reuslts = keep
        return x


class SkipConnMLP(nn.Module):
    """Leaky-ReLU MLP that re-concatenates its input at ``skip_layers`` (network.py:36-75)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, skip_layers=(), bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.skip_layers = list(skip_layers)
        layers = []
        for i in range(num_layers):
            fan_in = dim_in if i == 0 else dim_hidden + (dim_in if i in self.skip_layers else 0)
            layers.append(nn.Linear(fan_in, dim_out if i == num_layers - 1 else dim_hidden, bias=bias))
        self.net = nn.ModuleList(layers)

    tc = False               # set by NeRFNetwork: run on the tensor cores (sanerf_b200.fused.skip_mlp) when possible
    precision = "fp32"

    def forward(self, x, save_intermedian_results=False):
        if self.tc and not save_intermedian_results and x.dim() >= 2:
            flat = x.reshape(-1, x.shape[-1])
            if fused.skip_mlp_supported(self, flat):
                return fused.skip_mlp(flat, self, self.precision).view(*x.shape[:-1], -1)
        x_in = x
        keep = [] if save_intermedian_results else None
        for i, layer in enumerate(self.net):
            if i in self.skip_layers:
                x = torch.cat([x, x_in], dim=-1)
            x = layer(x)
            if i + 1 < self.num_layers:
                x = F.leaky_relu(x)
            if keep is not None:
                keep.append(x.detach())
        if keep is not None:
            self.intermedian_reuslts = keep
        return x


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt):
        super().__init__(opt)
        if getattr(opt, "with_mask", False):
            raise NotImplementedError("mask heads (stage 3) are outside the B200 render path")
        self.geom_feat_dim = 15
        # tensor-core field head (sanerf_b200/fused.py): "fp32" = 3xTF32 split products (fp32 parity), "tf32" = one pass
        self.tc_head = bool(getattr(opt, "tc_head", True))
        self.mlp_precision = getattr(opt, "mlp_precision", "fp32")

        self.grid, self.grid_in_dim = get_encoder("hashgrid", input_dim=3, level_dim=2, num_levels=16,
                                                  log2_hashmap_size=19, desired_resolution=2048 * self.bound)
        self.grid_mlp = MLP(self.grid_in_dim, 1 + self.geom_feat_dim, 64, 3, bias=False)
        self.view_encoder, self.view_in_dim = get_encoder("sh", input_dim=3, degree=4)
        self.view_mlp = MLP(self.geom_feat_dim + self.view_in_dim, 3, 32, 3, bias=False)

        if self.opt.with_sam:
            self.s_grid, self.s_dim = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=8,
                                                  base_resolution=16, log2_hashmap_size=19, desired_resolution=512)
            # fan-in is fixed at s_dim + 15 + 16 + 4 = 163 whatever the flags say (network.py:121)
            self.samvit_mlp = nn.Sequential(
                SkipConnMLP(self.s_dim + self.geom_feat_dim + self.view_in_dim + 4, 256, 256, 5, skip_layers=[2],
                            bias=True),
                nn.LayerNorm(256))
            self.samvit_mlp[0].tc, self.samvit_mlp[0].precision = self.tc_head, self.mlp_precision

        self.prop_encoders = nn.ModuleList()
        self.prop_mlp = nn.ModuleList()
        for finest in (128, 256):  # network.py:211-219
            enc, width = get_encoder("hashgrid", input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17,
                                     desired_resolution=finest)
            self.prop_encoders.append(enc)
            self.prop_mlp.append(MLP(width, 1, 16, 2, bias=False))

    def common_forward(self, x, save_intermedian_results=False):
        grid_output = self.grid(x, bound=self.bound)
        f = self.grid_mlp(grid_output, save_intermedian_results)
        return trunc_exp(f[..., 0]), f[..., 1:], grid_output

    def forward(self, x, d, save_intermedian_results=False, **kwargs):
        """Per-sample interface of the reference (network.py:231-246): d is [..., 3] per sample."""
        sigma, feat, grid_output = self.common_forward(x, save_intermedian_results)
        return {"sigma": sigma, "geo_feat": feat, "color": torch.cat([feat, self.view_encoder(d)], dim=-1),
                "grid_output": grid_output}

    def field(self, xyz, rays_d):
        """Renderer fast path: xyz [N,T,3], ONE direction per ray [N,3]; the SH basis is evaluated N times and
        broadcast over T (identical values to network.py:237, which evaluates it N*T times)."""
        sigma, feat, grid_output = self.common_forward(xyz)
        sh = self.view_encoder(rays_d)                                   # normalises internally
        color = torch.cat([feat, sh.unsqueeze(1).expand(-1, xyz.shape[1], -1)], dim=-1)
        return {"sigma": sigma, "geo_feat": feat, "color": color, "grid_output": grid_output}

    # ---- fast-path hooks used by NeRFRenderer._run_fused: positions already mapped to [0,1]^3 -------------
    @staticmethod
    def _encode_unit(enc, x01):
        flat = x01.reshape(-1, enc.input_dim)
        out = grid_encode(flat, enc.embeddings, enc.offsets, enc.per_level_scale, enc.base_resolution, False,
                          enc.gridtype_id, enc.align_corners, enc.interp_id, None)
        return out.view(*x01.shape[:-1], enc.output_dim)

    def density_unit(self, x01, proposal):
        enc, mlp = self.prop_encoders[proposal], self.prop_mlp[proposal]
        if fused.prop_density_supported(enc, mlp):
            return fused.prop_density(x01, enc, mlp)
        return trunc_exp(mlp(self._encode_unit(enc, x01)).squeeze(-1))

    def head_unit(self, x01):
        """[N,T,3] -> [N,T,16]: grid_mlp(grid(x)); column 0 is the density logit (network.py:223-227)."""
        if self.tc_head and fused.field_head_supported(self.grid, self.grid_mlp):
            return fused.field_head(x01, self.grid, self.grid_mlp, self.mlp_precision)
        return self.grid_mlp(self._encode_unit(self.grid, x01))

    def features_unit(self, x01):
        return self._encode_unit(self.s_grid, x01)

    def density(self, x, proposal=-1):
        if 0 <= proposal < len(self.prop_encoders):
            h = self.prop_encoders[proposal](x, bound=self.bound)
            return {"sigma": trunc_exp(self.prop_mlp[proposal](h).squeeze(-1))}
        return {"sigma": self.common_forward(x)[0]}

    def apply_total_variation(self, w):
        (self.s_grid if self.opt.with_sam else self.grid).grad_total_variation(w)

    def apply_weight_decay(self, w):
        (self.s_grid if self.opt.with_sam else self.grid).grad_weight_decay(w)

    def get_params(self, lr):
        groups = [self.grid, self.grid_mlp, self.view_mlp, self.prop_encoders, self.prop_mlp]
        if self.opt.with_sam:
            groups += [self.s_grid, self.samvit_mlp]
        return [{"params": m.parameters(), "lr": lr} for m in groups]
