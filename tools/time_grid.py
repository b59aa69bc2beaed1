"""GPU timing of hash-grid backward / proposal kernels per level group on ray-ordered samples (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib, fused
from gridencoder import GridEncoder

lib = _lib.load()
N = 8192
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
o = (torch.rand(N, 3, generator=g) - 0.5).cuda()
d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).cuda()
aabb = torch.tensor([-128.0] * 3 + [128.0] * 3).cuda()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
st = _lib.current_stream(o.device)

def timeit(fn, n=8, pre=None):
    ts = []
    for i in range(n + 2):
        if pre: pre()
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

for T, (L, Tl, fin, C) in {32: (16, 19, 4096, 2), 128: (5, 17, 128, 2)}.items():
    enc = GridEncoder(input_dim=3, num_levels=L, level_dim=C, base_resolution=16, log2_hashmap_size=Tl, desired_resolution=fin).cuda()
    noise = torch.rand(N, T + 1, device="cuda")
    _, _, _, x01 = fused.sample_uniform(o, d, aabb, 0.2, T, noise)
    x = x01.reshape(-1, 3).contiguous()
    B = x.shape[0]
    S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
    grad = torch.randn(B, L * C, device="cuda")
    gt = torch.zeros_like(enc.embeddings)
    out = torch.empty(B, L * C, device="cuda")
    for rand in (False, True):
        xx = torch.rand_like(x) if rand else x
        res = []
        for ml in sorted(set([1, 2, 3, 4, 5, 8, 12, 16]) & set(range(1, L + 1))):
            def bwd():
                rc = lib.sanerf_grid_encode_backward(grad.data_ptr(), xx.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), gt.data_ptr(),
                                                     B, 3, C, L, ml, S, H, None, None, 0, 0, 0, _lib.SANERF_F32, _lib.LAYOUT_BLC, st)
                _lib.check(rc, "bwd")
            def fwd():
                rc = lib.sanerf_grid_encode_forward(xx.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3, C, L, ml, S, H,
                                                    None, 0, 0, 0, _lib.SANERF_F32, _lib.LAYOUT_BLC, 1, st)
                _lib.check(rc, "fwd")
            res.append((ml, round(timeit(fwd), 1), round(timeit(bwd), 1)))
        print(f"L={L} T=2^{Tl} B={B} {'random' if rand else 'ray-ordered'} x: (max_level, fwd us, bwd us): {res}", flush=True)
