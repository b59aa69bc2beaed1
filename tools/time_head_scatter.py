"""Head backward: separate scatter kernel vs fused scatter (dedicated warps), ray-ordered samples, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib, fused
from gridencoder import GridEncoder
lib = _lib.load()
dev = torch.device("cuda", 0)
N, T = 8192, 32
g = torch.Generator().manual_seed(1)
o = (torch.rand(N, 3, generator=g) - 0.5).to(dev); d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(dev)
aabb = torch.tensor([-128.0] * 3 + [128.0] * 3, device=dev)
x = fused.sample_uniform(o, d, aabb, 0.2, T, torch.rand(N, T + 1, device=dev))[3].reshape(-1, 3).contiguous()
B = x.shape[0]
enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096).to(dev)
w1 = (torch.randn(64, 32) / 32 ** 0.5).to(dev); w2 = (torch.randn(64, 64) / 8).to(dev); w3 = (torch.randn(16, 64) / 8).to(dev)
out = torch.empty(B, 16, device=dev); e = torch.empty(B, 32, device=dev); h1 = torch.empty(B, 64, device=dev); h2 = torch.empty(B, 64, device=dev)
g_out = torch.randn(B, 16, device=dev); g_enc = torch.empty(B, 32, device=dev); gt = torch.zeros_like(enc.embeddings)
gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
st = _lib.current_stream(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
_lib.check(lib.sanerf_field_head_forward(x.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), S, H, None, w1.data_ptr(), w2.data_ptr(),
                                         w3.data_ptr(), B, e.data_ptr(), h1.data_ptr(), h2.data_ptr(), out.data_ptr(), 0, st), "fwd")
def separate():
    _lib.check(lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                              g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), 0, st), "hb")
    _lib.check(lib.sanerf_grid_encode_backward(g_enc.data_ptr(), x.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), gt.data_ptr(), B, 3, 2, 16, 16,
                                               S, H, None, None, 0, 0, 0, _lib.SANERF_F32, _lib.LAYOUT_BLC, st), "gb")
def fused_scatter():
    _lib.check(lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                              None, x.data_ptr(), enc.offsets.data_ptr(), S, H, gt.data_ptr(), gw[0].data_ptr(), gw[1].data_ptr(),
                                              gw[2].data_ptr(), 0, st), "hb")
def timeit(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
print(f"head backward + separate scatter kernel: {timeit(separate):.1f} us;  fused scatter (dedicated warps): {timeit(fused_scatter):.1f} us")
