"""Generates the golden fixtures under ``tests/golden/`` from the REAL reference.

TEST INFRASTRUCTURE ONLY.  Two modes:

``python -m oracle.make_golden --cpu``   (build container; needs /root/reference)
    Imports the reference's pure-torch modules (``nerf/renderer.py``, ``nerf/network.py``,
    ``activation.py``, ``encoding.py``) on CPU and records seeded input/output vectors of
    ``contract``, ``sample_pdf``, ``near_far_from_aabb``, ``MLP``, ``SkipConnMLP``,
    ``FreqEncoder_torch``, ``trunc_exp`` and of ``NeRFRenderer.run`` driven by an analytic stub
    field (so sampling + sigma->weights + compositing + proposal loss are pinned end to end without
    CUDA).  Also evaluates the 64 SH polynomials and their 192 partial derivatives straight from the
    text of ``shencoder/src/shencoder.cu`` and records ``GridEncoder.__init__``'s offsets for every
    table shape the reference builds.  -> ``tests/golden/ref_cpu.npz``

``python -m oracle.make_golden --gpu``   (B200 box; needs oracle/_ref/*.so, no /root/reference)
    Runs the unmodified reference CUDA extensions (rebuilt for sm_100 by ``oracle/build_ref.py``)
    on seeded inputs: grid forward/backward (fp32 + fp16, hash + tiled, linear + smoothstep,
    dy_dx), SH and freq forward/backward, TV / weight-decay gradients; and the level resolutions of
    the five reference table shapes recovered with probe tables.
    -> ``gpurun_out/ref_gpu.npz`` (copied into ``tests/golden/`` and committed).

Stubs: ``mcubes``, ``trimesh`` and ``torch_efficient_distloss`` are not installed; the first two are
unused by the functions called here, the third is replaced by ``oracle.render_torch.eff_distloss``
(so the distortion-loss golden only pins the oracle against itself — stated in DESIGN.md).
"""
from __future__ import annotations

import argparse
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("SANERF_REFERENCE", "/root/reference")

# table shapes the reference instantiates (nerf/network.py:102,111,211,216) + BASELINE cfg5
TABLE_SHAPES = {
    "main": dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096),
    "s_grid": dict(num_levels=16, level_dim=8, base_resolution=16, log2_hashmap_size=19, desired_resolution=512),
    "prop0": dict(num_levels=5, level_dim=2, base_resolution=16, log2_hashmap_size=17, desired_resolution=128),
    "prop1": dict(num_levels=5, level_dim=2, base_resolution=16, log2_hashmap_size=17, desired_resolution=256),
    "cfg5": dict(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=22, desired_resolution=4096),
}


# ------------------------------------------------------------------------------------ CPU
def _import_reference():
    from . import render_torch

    for name in ("mcubes", "trimesh"):
        sys.modules.setdefault(name, types.ModuleType(name))
    dl = types.ModuleType("torch_efficient_distloss")
    dl.eff_distloss = render_torch.eff_distloss
    sys.modules.setdefault("torch_efficient_distloss", dl)
    # the reference's encoders import their compiled backends at module import; give them the rebuilt ones
    from . import build_ref
    for ext in ("gridencoder", "shencoder", "freqencoder"):
        try:
            sys.modules.setdefault(f"_{ext}", build_ref.load(ext))
        except Exception as e:  # noqa: BLE001
            print(f"[make_golden] oracle/_ref/_ref_{ext}.so not loadable ({e}); offsets golden skipped")
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import warnings
    warnings.filterwarnings("ignore")
    import activation as ref_activation  # noqa: E402
    import encoding as ref_encoding  # noqa: E402
    from nerf import network as ref_network  # noqa: E402
    from nerf import renderer as ref_renderer  # noqa: E402
    return ref_activation, ref_encoding, ref_renderer, ref_network


def parse_sh_polynomials(path):
    """Python expressions of outputs[i], dx[i], dy[i], dz[i] from shencoder.cu:50-349."""
    src = open(path).read()

    def table(prefix):
        d = {}
        for m in re.finditer(r"\b%s\[(\d+)\]\s*=\s*([^;]+);" % prefix, src):
            expr = re.sub(r"(\d+\.\d+(?:e[-+]?\d+)?|\d+\.)f", r"\1", m.group(2))
            expr = re.sub(r"pow\(z,\s*3\)", "z**3", expr)
            d[int(m.group(1))] = expr.strip()
        return [d[i] for i in range(64)]

    return table("outputs"), table("dx"), table("dy"), table("dz")


def eval_sh_polynomials(tables, v):
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    env = dict(x=x, y=y, z=z, xy=x * y, xz=x * z, yz=y * z, x2=x * x, y2=y * y, z2=z * z, xyz=x * y * z)
    env.update(x4=env["x2"] ** 2, y4=env["y2"] ** 2, z4=env["z2"] ** 2)
    env.update(x6=env["x4"] * env["x2"], y6=env["y4"] * env["y2"], z6=env["z4"] * env["z2"])
    ev = lambda exprs: np.stack([np.broadcast_to(eval(e, {}, env), x.shape) for e in exprs], 1)  # noqa: E731
    out, dx, dy, dz = (ev(t) for t in tables)
    return out, np.stack([dx, dy, dz], 1)


def stub_sigma(x, k):
    """Analytic density used by the stub field (shared with tests/test_oracle_render.py)."""
    return torch.exp(1.5 * torch.sin(3.0 * x[..., 0] + k) + 1.0 * torch.cos(2.0 * x[..., 1] - x[..., 2]) + 0.5)


def stub_color(x, d, C=31):
    ph = torch.arange(C, dtype=x.dtype, device=x.device) * 0.37
    return torch.sin(x[..., :1] * 2.0 + x[..., 1:2] - 0.5 * x[..., 2:3] + d[..., :1] + ph)


def make_cpu(out_path):
    ref_activation, ref_encoding, ref_renderer, ref_network = _import_reference()
    g = {}
    gen = torch.Generator().manual_seed(20240601)
    rnd = lambda *s: torch.rand(*s, generator=gen)  # noqa: E731
    rndn = lambda *s: torch.randn(*s, generator=gen)  # noqa: E731

    # --- contract (renderer.py:60-69)
    x = rndn(257, 3) * 3.0
    g["contract_in"], g["contract_out"] = x.numpy(), ref_renderer.contract(x.clone()).numpy()
    # --- near/far (renderer.py:122-139)
    o, d = rnd(120, 3) - 0.5, torch.nn.functional.normalize(rndn(120, 3), dim=-1)
    o[:20] = o[:20] * 400  # some rays start outside the box / miss it
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3)
    near, far = ref_renderer.near_far_from_aabb(o, d, aabb, 0.2)
    g.update(nf_o=o.numpy(), nf_d=d.numpy(), nf_near=near.numpy(), nf_far=far.numpy())
    # --- sample_pdf, deterministic (renderer.py:84-119)
    bins = torch.sort(rnd(64, 33), dim=-1).values
    w = rnd(64, 32) ** 3
    w[:4] = 0
    g.update(pdf_bins=bins.numpy(), pdf_w=w.numpy(), pdf_out=ref_renderer.sample_pdf(bins, w, 17, False).numpy())
    # --- trunc_exp fwd/bwd (activation.py)
    xe = (rndn(513) * 8).requires_grad_(True)
    ye = ref_activation.trunc_exp(xe)
    ge = rndn(513)
    ye.backward(ge)
    g.update(texp_x=xe.detach().numpy(), texp_y=ye.detach().numpy(), texp_g=ge.numpy(), texp_dx=xe.grad.numpy())
    # --- FreqEncoder_torch (encoding.py:6-44)
    fe = ref_encoding.FreqEncoder_torch(input_dim=3, max_freq_log2=5, N_freqs=6, log_sampling=True)
    xf = rnd(100, 3) * 2 - 1
    g.update(freq_x=xf.numpy(), freq_out=fe(xf).numpy())
    # --- MLP / SkipConnMLP (network.py:9-75)
    torch.manual_seed(7)
    mlp = ref_network.MLP(32, 16, 64, 3, bias=False)
    skip = ref_network.SkipConnMLP(19, 8, 24, 5, skip_layers=[2], bias=True)
    xm, xs = rndn(50, 32), rndn(20, 19)
    g.update(mlp_x=xm.numpy(), mlp_y=mlp(xm, save_intermedian_results=False).detach().numpy(),
             skip_x=xs.numpy(), skip_y=skip(xs).detach().numpy())
    for k, v in mlp.state_dict().items():
        g["mlp_sd." + k] = v.numpy()
    for k, v in skip.state_dict().items():
        g["skip_sd." + k] = v.numpy()

    # --- NeRFRenderer.run with an analytic stub field (renderer.py:221-390)
    opt = types.SimpleNamespace(bound=128, contract=True, min_near=0.2, density_thresh=10, num_steps=[128, 64, 32],
                                background="last_sample", with_sam=False, with_mask=False, sum_after_mlp=False,
                                sam_use_view_direction=True, mask_mlp_type="default", lambda_proposal=1.0,
                                lambda_distort=0.02, max_ray_batch=16384)

    class StubField(ref_renderer.NeRFRenderer):
        def __init__(self, opt):
            super().__init__(opt)
            torch.manual_seed(11)
            self.view_mlp = ref_network.MLP(31, 3, 32, 3, bias=False)

        def density(self, x, proposal=-1):
            return {"sigma": stub_sigma(x, float(proposal))}

        def forward(self, x, d, **kw):
            color = stub_color(x, d)
            return {"sigma": stub_sigma(x, 2.0), "color": color, "geo_feat": color[..., :15], "grid_output": None}

    model = StubField(opt).train()
    ro = rnd(96, 3) - 0.5
    rd = torch.nn.functional.normalize(rndn(96, 3), dim=-1)
    res = model.run(ro, rd, perturb=False, update_proposal=True)
    g.update(run_o=ro.numpy(), run_d=rd.numpy(), run_image=res["image"].detach().numpy(),
             run_depth=res["depth"].detach().numpy(), run_wsum=res["weights_sum"].detach().numpy(),
             run_weights=res["weights"].detach().numpy(), run_prop_loss=res["proposal_loss"].detach().numpy(),
             run_dist_loss=res["distort_loss"].detach().numpy())
    for k, v in model.view_mlp.state_dict().items():
        g["run_view_sd." + k] = v.numpy()

    # --- SH polynomials from the reference source text (shencoder.cu:50-349)
    tables = parse_sh_polynomials(os.path.join(REFERENCE, "shencoder", "src", "shencoder.cu"))
    v = torch.nn.functional.normalize(rndn(400, 3).double(), dim=-1).numpy()
    sh_out, sh_jac = eval_sh_polynomials(tables, v)
    g.update(sh_dirs=v, sh_out=sh_out, sh_jac=sh_jac)

    # --- GridEncoder.__init__ offsets (grid.py:103-146)
    if "_gridencoder" in sys.modules:
        from gridencoder.grid import GridEncoder as RefGrid
        for name, kw in TABLE_SHAPES.items():
            if name == "cfg5":
                continue  # 42.6 M rows x 2 fp32 = 340 MB just to read the offsets: use the formula test instead
            enc = RefGrid(input_dim=3, **kw)
            g[f"offsets.{name}"] = enc.offsets.numpy()
            g[f"scale.{name}"] = np.float64(enc.per_level_scale)

    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **g)
    print(f"[make_golden] wrote {out_path}: {len(g)} arrays, {os.path.getsize(out_path) / 1024:.0f} KiB")


# ------------------------------------------------------------------------------------ GPU
def make_gpu(out_path):
    from . import build_ref, grid_np

    dev = torch.device("cuda:0")
    ref_grid, ref_sh, ref_freq = build_ref.load("gridencoder"), build_ref.load("shencoder"), build_ref.load("freqencoder")
    g = {}
    gen = torch.Generator(device="cpu").manual_seed(777)

    def small_table(D, L, C, base, log2T, desired, dtype):
        scale = grid_np.per_level_scale(desired, base, L)
        offs = grid_np.level_offsets(D, L, scale, base, log2T)
        tab = (torch.rand(int(offs[-1]), C, generator=gen) * 2 - 1).to(dtype)
        return float(np.log2(scale)), offs, tab

    cases = {
        # name: (D, L, C, base, log2T, desired, dtype, gridtype, align, interp)
        "hash_f32": (3, 8, 2, 4, 10, 96, torch.float32, 0, False, 0),
        "hash_c8": (3, 6, 8, 4, 9, 64, torch.float32, 0, False, 0),
        "tiled_smooth": (3, 6, 2, 4, 10, 64, torch.float32, 1, False, 1),
        "align_d2": (2, 6, 4, 4, 10, 128, torch.float32, 0, True, 0),
        "hash_f16": (3, 8, 2, 4, 10, 96, torch.float16, 0, False, 0),
        "c1_f32": (3, 5, 1, 4, 10, 40, torch.float32, 0, False, 0),
    }
    for name, (D, L, C, base, log2T, desired, dtype, gridtype, align, interp) in cases.items():
        S, offs, tab = small_table(D, L, C, base, log2T, desired, dtype)
        B = 384
        x = torch.rand(B, D, generator=gen)
        x[:8] = torch.tensor([0.0, 1.0, 0.5, 1.0 - 1e-7, 1e-7, 0.25, 0.75, 0.999])[:, None]
        x[8] = -0.01  # out of range sample
        x[9] = 1.01
        xd, td, od = x.to(dev), tab.to(dev), torch.from_numpy(offs).to(dev)
        out = torch.empty(L, B, C, device=dev, dtype=dtype)
        dydx = torch.empty(B, L * D * C, device=dev, dtype=dtype)
        ref_grid.grid_encode_forward(xd, td, od, out, B, D, C, L, L, S, base, dydx, gridtype, align, interp)
        grad = (torch.rand(L, B, C, generator=gen) * 2 - 1).to(dtype)
        gd = grad.to(dev)
        gtab = torch.zeros_like(td)
        gin = torch.zeros(B, D, device=dev, dtype=dtype)
        ref_grid.grid_encode_backward(gd, xd, td, od, gtab, B, D, C, L, L, S, base, dydx, gin, gridtype, align, interp)
        torch.cuda.synchronize()
        g.update({f"grid.{name}.meta": np.array([D, L, C, base, log2T, desired, gridtype, int(align), interp,
                                                 int(dtype == torch.float16)], dtype=np.int64),
                  f"grid.{name}.S": np.float32(S), f"grid.{name}.offsets": offs,
                  f"grid.{name}.x": x.numpy(), f"grid.{name}.table": tab.numpy(),
                  f"grid.{name}.out_LBC": out.cpu().numpy(), f"grid.{name}.dy_dx": dydx.cpu().numpy(),
                  f"grid.{name}.grad_LBC": grad.numpy(), f"grid.{name}.grad_table": gtab.cpu().numpy(),
                  f"grid.{name}.grad_inputs": gin.cpu().numpy()})
        if dtype == torch.float32 and D == 3:
            gtv = torch.zeros_like(td)
            ref_grid.grad_total_variation(xd, td, gtv, od, 0.37, B, D, C, L, S, base, gridtype, align)
            gwd = torch.zeros_like(td)
            ref_grid.grad_weight_decay(td, gwd, od, 0.1, td.shape[0], C, L)
            torch.cuda.synchronize()
            g[f"grid.{name}.tv"] = gtv.cpu().numpy()
            if name == "tiled_smooth":
                g[f"grid.{name}.wd"] = gwd.cpu().numpy()

    # level resolutions of the real table shapes, recovered from the reference kernel with a
    # probe: a table whose row r holds (r - offset_l), sampled at cell centres of axis x
    # (SURVEY §8 c7).  out = row index of the (unique, weight-1) corner -> resolution via the
    # largest x-row seen on dense levels; on hashed levels we record the raw outputs instead.
    for name, kw in TABLE_SHAPES.items():
        D, L, C, base = 3, kw["num_levels"], kw["level_dim"], kw["base_resolution"]
        scale = grid_np.per_level_scale(kw["desired_resolution"], base, L)
        S = float(np.log2(scale))
        offs = grid_np.level_offsets(D, L, scale, base, kw["log2_hashmap_size"])
        rows_total = int(offs[-1])
        tab = torch.zeros(rows_total, C, device=dev)
        for l in range(L):
            n = int(offs[l + 1] - offs[l])
            r = torch.arange(n, device=dev, dtype=torch.float32)
            tab[int(offs[l]):int(offs[l + 1]), 0] = torch.remainder(r, 251.0)
            if C > 1:
                tab[int(offs[l]):int(offs[l + 1]), 1] = torch.remainder(r, 241.0)
        od = torch.from_numpy(offs).to(dev)
        B = 768
        x = torch.rand(B, 3, generator=gen)
        xd = x.to(dev)
        out = torch.empty(L, B, C, device=dev)
        ref_grid.grid_encode_forward(xd, tab, od, out, B, D, C, L, L, S, base, None, 0, False, 0)
        torch.cuda.synchronize()
        g[f"probe.{name}.x"] = x.numpy()
        g[f"probe.{name}.out_LBC"] = out[:, :, :min(C, 2)].cpu().numpy()
        g[f"probe.{name}.offsets"] = offs
        g[f"probe.{name}.S"] = np.float32(S)
        del tab, out

    # SH / freq extensions
    dirs = torch.nn.functional.normalize(torch.randn(120, 3, generator=gen), dim=-1)
    for deg in (1, 4, 8):
        o = torch.empty(120, deg * deg, device=dev)
        j = torch.empty(120, 3 * deg * deg, device=dev)
        ref_sh.sh_encode_forward(dirs.to(dev), o, 120, 3, deg, j)
        gr = torch.randn(120, deg * deg, generator=gen)
        gi = torch.zeros(120, 3, device=dev)
        ref_sh.sh_encode_backward(gr.to(dev), dirs.to(dev), 120, 3, deg, j, gi)
        torch.cuda.synchronize()
        g.update({f"sh.{deg}.out": o.cpu().numpy(), f"sh.{deg}.dy_dx": j.cpu().numpy(),
                  f"sh.{deg}.grad": gr.numpy(), f"sh.{deg}.grad_inputs": gi.cpu().numpy()})
    g["sh.dirs"] = dirs.numpy()
    xf = torch.rand(200, 3, generator=gen) * 2 - 1
    deg, Cf = 6, 3 + 3 * 2 * 6
    of = torch.empty(200, Cf, device=dev)
    ref_freq.freq_encode_forward(xf.to(dev), 200, 3, deg, Cf, of)
    grf = torch.randn(200, Cf, generator=gen)
    gif = torch.zeros(200, 3, device=dev)
    ref_freq.freq_encode_backward(grf.to(dev), of, 200, 3, deg, Cf, gif)
    torch.cuda.synchronize()
    g.update(freq_x=xf.numpy(), freq_out=of.cpu().numpy(), freq_grad=grf.numpy(), freq_grad_inputs=gif.cpu().numpy())

    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **g)
    print(f"[make_golden] wrote {out_path}: {len(g)} arrays, {os.path.getsize(out_path) / 1024:.0f} KiB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.cpu:
        make_cpu(args.out or os.path.join(GOLDEN, "ref_cpu.npz"))
    if args.gpu:
        make_gpu(args.out or os.path.join(ROOT, "gpurun_out", "ref_gpu.npz"))
