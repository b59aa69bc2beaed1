"""GPU timing of the fused field head (forward / backward) at the bench shape, CUDA events, L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
from gridencoder import GridEncoder

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096).cuda()
w1 = (torch.randn(64, 32) / 32 ** 0.5).cuda(); w2 = (torch.randn(64, 64) / 8).cuda(); w3 = (torch.randn(16, 64) / 8).cuda()
x01 = torch.rand(B, 3, device="cuda")
out = torch.empty(B, 16, device="cuda"); Bp = (B + 127) // 128 * 128
e = torch.empty(Bp, 32, device="cuda"); h1 = torch.empty(Bp, 64, device="cuda"); h2 = torch.empty(Bp, 64, device="cuda")
g_out = torch.randn(B, 16, device="cuda"); g_enc = torch.empty(B, 32, device="cuda")
gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
st = _lib.current_stream(x01.device)
S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)

def fwd(prec, train):
    rc = lib.sanerf_field_head_forward(x01.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), S, H, None, w1.data_ptr(), w2.data_ptr(),
                                       w3.data_ptr(), B, e.data_ptr() if train else None, h1.data_ptr() if train else None,
                                       h2.data_ptr() if train else None, out.data_ptr(), prec, st)
    _lib.check(rc, "fwd")
def bwd(prec):
    rc = lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                        g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), prec, st)
    _lib.check(rc, "bwd")
def timeit(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(min(ts))
for prec in (0, 1):
    print(f"B={B} precision={prec}: fwd(train) us {timeit(lambda: fwd(prec, True))}  fwd(infer) us {timeit(lambda: fwd(prec, False))}  "
          f"bwd us {timeit(lambda: bwd(prec))}", flush=True)

# warm (no flush between launches): what the cold caches (instructions, weights, first tile) cost
def warm(fn, n=10):
    ts = []
    for i in range(n + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(min(ts))
print(f"B={B} warm: fwd(train) us {warm(lambda: fwd(0, True))}  bwd us {warm(lambda: bwd(0))}", flush=True)
