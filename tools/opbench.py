"""Per-operator roofline table on one B200 (SURVEY §8 d): CUDA events, L2 flushed between iterations, median of 10.

  python tools/opbench.py > profiles/opbench.md

Algorithmic bytes follow SURVEY §8 d4; peak = MEASURED_PEAKS.json (HBM copy GB/s)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib, fused
from sanerf_b200.ops import composite
from gridencoder import GridEncoder
from gridencoder.grid import grid_encode
from shencoder import SHEncoder

peak = 6650.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
lib = _lib.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timeit(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

rows = []
def report(name, cfg, us, nbytes):
    gbs = nbytes / us / 1e3
    rows.append(f"| {name} | {cfg} | {us:.1f} | {nbytes / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} |")

def ray_samples(N, T):
    g = torch.Generator().manual_seed(1)
    o = (torch.rand(N, 3, generator=g) - 0.5).to(dev)
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(dev)
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3, device=dev)
    noise = torch.rand(N, T + 1, device=dev)
    return fused.sample_uniform(o, d, aabb, 0.2, T, noise)[3].reshape(-1, 3).contiguous()

# ---- hash-grid encode forward / backward
for name, (L, C, Tl, fin, dtype, N, T) in {
    "main grid L16 F2 T2^19 fp32 (cfg1: 4096 rays x 64)": (16, 2, 19, 4096, torch.float32, 4096, 64),
    "main grid L16 F2 T2^19 fp32 (cfg2 final level: 8192 x 32)": (16, 2, 19, 4096, torch.float32, 8192, 32),
    "SAM grid L16 F8 T2^19 fp32 (cfg3: 4096 x 32)": (16, 8, 19, 512, torch.float32, 4096, 32),
    "large grid L16 F2 T2^22 fp16 (cfg5: 2^20 samples)": (16, 2, 22, 4096, torch.float16, 8192, 128),
}.items():
    enc = GridEncoder(input_dim=3, num_levels=L, level_dim=C, base_resolution=16, log2_hashmap_size=Tl, desired_resolution=fin).to(dev)
    table = enc.embeddings.detach().to(dtype).contiguous()
    sz = 2 if dtype == torch.float16 else 4
    for order in ("ray-ordered", "uniform random"):
        x = ray_samples(N, T) if order == "ray-ordered" else torch.rand(N * T, 3, device=dev)
        B = x.shape[0]
        S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
        out = torch.empty(B, L * C, device=dev, dtype=dtype)
        grad = torch.randn(B, L * C, device=dev).to(dtype)
        gt = torch.zeros_like(table)
        dt = _lib.SANERF_F16 if dtype == torch.float16 else _lib.SANERF_F32
        st = _lib.current_stream(dev)
        def fwd():
            _lib.check(lib.sanerf_grid_encode_forward(x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3, C, L, L, S, H,
                                                      None, 0, 0, 0, dt, _lib.LAYOUT_BLC, 0, st), "fwd")
        def bwd():
            _lib.check(lib.sanerf_grid_encode_backward(grad.data_ptr(), x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), gt.data_ptr(),
                                                       B, 3, C, L, L, S, H, None, None, 0, 0, 0, dt, _lib.LAYOUT_BLC, st), "bwd")
        nbytes = B * (12 + L * 8 * C * sz + L * C * sz)
        report("grid encode forward", f"{name}, {order}", timeit(fwd), nbytes)
        report("grid encode backward (scatter)", f"{name}, {order}", timeit(bwd), nbytes)
    del enc, table

# ---- compositing (C ABI with caller-allocated outputs: no torch allocator / autograd time inside the events)
for name, (N, T, C) in {"cfg1: 4096 rays, T=64, C=3": (4096, 64, 3), "RGB training: 8192 rays, T=32, C=31": (8192, 32, 31),
                        "SAM: 4096 rays, T=32, C=159": (4096, 32, 159), "frame: 65536 rays, T=32, C=128": (65536, 32, 128),
                        "frame: 262144 rays, T=32, C=31": (262144, 32, 31)}.items():
    sig = torch.rand(N, T, device=dev).exp()
    bins = torch.sort(torch.rand(N, T + 1, device=dev), -1).values * 4 + 0.2
    deltas, ts = (bins[:, 1:] - bins[:, :-1]).contiguous(), ((bins[:, 1:] + bins[:, :-1]) / 2).contiguous()
    feats = torch.randn(N, T, C, device=dev)
    w, ws, dp, out = torch.empty(N, T, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, C, device=dev)
    alive = torch.empty(N, device=dev, dtype=torch.int32)
    go, gws, gdp = torch.randn(N, C, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)
    gs, gf = torch.empty(N, T, device=dev), torch.empty(N, T, C, device=dev)
    st = _lib.current_stream(dev)
    def cf():
        _lib.check(lib.sanerf_composite_forward(sig.data_ptr(), deltas.data_ptr(), ts.data_ptr(), feats.data_ptr(), 0, None, N, T, C, 1, 0.0,
                                                w.data_ptr(), ws.data_ptr(), dp.data_ptr(), out.data_ptr(), alive.data_ptr(), st), "cf")
    def cb():
        _lib.check(lib.sanerf_composite_backward(sig.data_ptr(), deltas.data_ptr(), ts.data_ptr(), feats.data_ptr(), 0, None, N, T, C, 1, 0.0,
                                                 w.data_ptr(), None, gws.data_ptr(), gdp.data_ptr(), go.data_ptr(), gs.data_ptr(), gf.data_ptr(),
                                                 0, st), "cb")
    report("composite forward", name, timeit(cf), N * (T * (12 + 4 * C) + 4 * (C + 2)))
    report("composite backward", name, timeit(cb), N * (T * (12 + 4 * C) + 4 * (C + 2) + T * (4 + 4 * C)))
    del feats, gf

# ---- fused field head (final level of the RGB step)
for order in ("ray-ordered", "uniform random"):
    B = 262144
    x = ray_samples(8192, 32) if order == "ray-ordered" else torch.rand(B, 3, device=dev)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096).to(dev)
    w1 = (torch.randn(64, 32) / 32 ** 0.5).to(dev); w2 = (torch.randn(64, 64) / 8).to(dev); w3 = (torch.randn(16, 64) / 8).to(dev)
    out = torch.empty(B, 16, device=dev); e = torch.empty(B, 32, device=dev); h1 = torch.empty(B, 64, device=dev); h2 = torch.empty(B, 64, device=dev)
    g_out = torch.randn(B, 16, device=dev); g_enc = torch.empty(B, 32, device=dev)
    gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
    S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
    st = _lib.current_stream(dev)
    def hf(train):
        _lib.check(lib.sanerf_field_head_forward(x.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), S, H, None, w1.data_ptr(),
                                                 w2.data_ptr(), w3.data_ptr(), B, e.data_ptr() if train else None, h1.data_ptr() if train else None,
                                                 h2.data_ptr() if train else None, out.data_ptr(), 0, st), "hf")
    def hb():
        _lib.check(lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                                  w3.data_ptr(), B, g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(),
                                                  gw[2].data_ptr(), 0, st), "hb")
    report("field head forward, inference (gather + MLP, 3xTF32)", f"B=262144, {order}", timeit(lambda: hf(False)), B * (12 + 1024 + 64))
    report("field head forward, training (+ saved activations)", f"B=262144, {order}", timeit(lambda: hf(True)), B * (12 + 1024 + 64 + 640))
    hf(True)
    report("field head backward (MLP, 3xTF32)", f"B=262144, {order}", timeit(hb), B * (640 + 64 + 128))

# ---- stage-2 kernels: ray-composited grid features, tensor-core GEMM of the samvit head, LayerNorm + MSE
from sanerf_b200 import fused as F2
enc = GridEncoder(input_dim=3, num_levels=16, level_dim=8, base_resolution=16, log2_hashmap_size=19, desired_resolution=512).to(dev)
Nr, Tr = 4096, 32
xr = ray_samples(Nr, Tr).view(Nr, Tr, 3).contiguous()
wr = torch.rand(Nr, Tr, device=dev)
f_out = torch.empty(Nr, 128, device=dev); g_ray = torch.randn(Nr, 128, device=dev); g_tab = torch.zeros_like(enc.embeddings)
S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
st = _lib.current_stream(dev)
rf_bytes = Nr * Tr * (16 + 16 * 8 * 8 * 4) + Nr * 128 * 4
report("ray features forward (encode + weighted ray sum)", "SAM grid L16 F8 T2^19, 4096 rays x 32, ray-ordered",
       timeit(lambda: _lib.check(lib.sanerf_ray_features_forward(xr.data_ptr(), wr.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(),
                                                                 Nr, Tr, 8, 16, S, H, f_out.data_ptr(), 0, st), "rf")), rf_bytes)
report("ray features backward (factorised-gradient scatter)", "SAM grid L16 F8 T2^19, 4096 rays x 32, ray-ordered",
       timeit(lambda: _lib.check(lib.sanerf_ray_features_backward(xr.data_ptr(), wr.data_ptr(), g_ray.data_ptr(), enc.offsets.data_ptr(),
                                                                  Nr, Tr, 8, 16, S, H, g_tab.data_ptr(), 0, 16, 0, st), "rb")), rf_bytes)
del enc, g_tab
Mg = 4096
for name, (N_, K_, kw) in {"forward layer 4096x256x256 (bias + leaky ReLU)": (256, 256, dict(act=True)),
                           "forward layer 4096x256x419 (skip layer, unaligned rows)": (256, 419, dict(act=True)),
                           "data gradient 4096x256x256 (B transposed, activation mask)": (256, 256, dict(b_trans=True, epilogue=1))}.items():
    A = torch.randn(Mg, K_, device=dev); Bm = torch.randn((K_, N_) if kw.get("b_trans") else (N_, K_), device=dev)
    Cm = torch.empty(Mg, N_, device=dev); bias = torch.randn(N_, device=dev); mask = torch.randn(Mg, N_, device=dev)
    us = timeit(lambda: F2.gemm_tc(A, Bm, Cm, Mg, N_, K_, bias=None if kw.get("epilogue") else bias, mask=mask if kw.get("epilogue") else None,
                                   mask_cols=N_ if kw.get("epilogue") else 0, **kw))
    rows.append(f"| gemm_tc (tcgen05 3xTF32) | {name} | {us:.1f} | {(Mg * K_ + N_ * K_ + Mg * N_) * 4 / 1e6:.1f} | "
                f"{2 * Mg * N_ * K_ / us / 1e6:.1f} TFLOP/s fp32-equivalent ({6 * Mg * N_ * K_ / us / 1e6:.1f} tf32 issued) | — |")
Ad = torch.randn(Mg, 256, device=dev); Xd = torch.randn(Mg, 256, device=dev); gW = torch.zeros(256, 256, device=dev)
us = timeit(lambda: F2.gemm_tc(Ad, Xd, gW, 256, 256, Mg, a_trans=True, b_trans=True, k_splits=16, epilogue=2))
rows.append(f"| gemm_tc (tcgen05 3xTF32) | weight gradient 256x256x4096 (both operands transposed, split-K 16, red.add) | {us:.1f} | "
            f"{(2 * Mg * 256 + 256 * 256) * 4 / 1e6:.1f} | {2 * 256 * 256 * Mg / us / 1e6:.1f} TFLOP/s fp32-equivalent | — |")
ln = torch.nn.LayerNorm(256).to(dev); ln.weight.grad = torch.zeros_like(ln.weight); ln.bias.grad = torch.zeros_like(ln.bias)
xo = torch.randn(Mg, 256, device=dev); tgt = torch.randn(1, 256, 64, 64, device=dev); lossb = torch.zeros(1, device=dev); yb = torch.empty_like(xo)
report("LayerNorm(256) + MSE, forward + backward in one kernel", "4096 rays, [1,256,64,64] target read in place",
       timeit(lambda: F2.layernorm_mse(xo, ln, tgt, lossb, yb)), Mg * 256 * 4 * 4)

# ---- SH
d = torch.nn.functional.normalize(torch.randn(262144, 3, device=dev), dim=-1)
sh_out = torch.empty(262144, 16, device=dev)
report("SH degree 4", "262144 directions",
       timeit(lambda: _lib.check(lib.sanerf_sh_encode_forward(d.data_ptr(), sh_out.data_ptr(), 262144, 3, 4, None, 0, _lib.current_stream(dev)), "sh")),
       262144 * (12 + 64))

print(f"# Per-operator roofline (one B200, fp32 unless noted; peak = {peak:.0f} GB/s measured HBM copy)\n")
print("L2 flushed (256 MiB write) before every timed launch; median of 10; `frac` = algorithmic GB/s / peak.  Tables that fit the")
print("126 MB L2 are re-read from L2 within a launch, so gather/scatter rows are bounded by L1/L2 request rates, not DRAM.\n")
print("| operator | configuration | µs | algorithmic MB | GB/s | frac |\n|---|---|---:|---:|---:|---:|")
print("\n".join(rows))
