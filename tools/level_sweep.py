"""Where the time of the main-grid gather / scatter goes, level by level: the kernels are launched with max_level = 1..16
(levels 0..max_level-1) and the differences are the per-level costs.  CUDA events, L2 flushed, median of 10.

  python tools/level_sweep.py [fp32|fp16] > gpurun_out/level_sweep.md"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib, fused
from gridencoder import GridEncoder

which = sys.argv[1] if len(sys.argv) > 1 else "fp32"
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

def ray_samples(N, T):
    g = torch.Generator().manual_seed(1)
    o = (torch.rand(N, 3, generator=g) - 0.5).to(dev)
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(dev)
    aabb = torch.tensor([-128.0] * 3 + [128.0] * 3, device=dev)
    noise = torch.rand(N, T + 1, device=dev)
    return fused.sample_uniform(o, d, aabb, 0.2, T, noise)[3].reshape(-1, 3).contiguous()

if which == "fp32":
    L, C, Tl, fin, dtype, N, T = 16, 2, 19, 4096, torch.float32, 8192, 32
else:
    L, C, Tl, fin, dtype, N, T = 16, 2, 22, 4096, torch.float16, 8192, 128
enc = GridEncoder(input_dim=3, num_levels=L, level_dim=C, base_resolution=16, log2_hashmap_size=Tl, desired_resolution=fin).to(dev)
table = enc.embeddings.detach().to(dtype).contiguous()
offs = enc.offsets.tolist()
S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
dt = _lib.SANERF_F16 if dtype == torch.float16 else _lib.SANERF_F32
print(f"# level sweep, {which}: L16 F2 T2^{Tl}, B = {N * T}")
for order in ("ray-ordered", "uniform random"):
    x = ray_samples(N, T) if order == "ray-ordered" else torch.rand(N * T, 3, device=dev)
    B = x.shape[0]
    out = torch.empty(B, L * C, device=dev, dtype=dtype)
    grad = torch.randn(B, L * C, device=dev).to(dtype)
    gt = torch.zeros_like(table)
    st = _lib.current_stream(dev)
    print(f"\n## {order}\n\n| levels 0..k | rows of level k | res | fwd us | d fwd | bwd us | d bwd |\n|---|---:|---:|---:|---:|---:|---:|")
    pf = pb = 0.0
    for ml in range(1, L + 1):
        def fwd():
            _lib.check(lib.sanerf_grid_encode_forward(x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3, C, L, ml, S, H,
                                                      None, 0, 0, 0, dt, _lib.LAYOUT_BLC, 0, st), "fwd")
        def bwd():
            _lib.check(lib.sanerf_grid_encode_backward(grad.data_ptr(), x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), gt.data_ptr(),
                                                       B, 3, C, L, ml, S, H, None, None, 0, 0, 0, dt, _lib.LAYOUT_BLC, st), "bwd")
        tf, tb = timeit(fwd), timeit(bwd)
        res = int(np.ceil(2.0 ** ((ml - 1) * S) * H))
        print(f"| {ml - 1} | {offs[ml] - offs[ml - 1]} | {res} | {tf:.1f} | {tf - pf:+.1f} | {tb:.1f} | {tb - pb:+.1f} |")
        pf, pb = tf, tb
