"""Field-head forward with the encoding supplied (no gather): isolates the MLP pipeline of the kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B = 262144
w1 = (torch.randn(64, 32) / 32 ** 0.5).to(dev); w2 = (torch.randn(64, 64) / 8).to(dev); w3 = (torch.randn(16, 64) / 8).to(dev)
enc_in = torch.randn(B, 32, device=dev)
out = torch.empty(B, 16, device=dev); e = torch.empty(B, 32, device=dev); h1 = torch.empty(B, 64, device=dev); h2 = torch.empty(B, 64, device=dev)
st = _lib.current_stream(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def run(prec, train):
    _lib.check(lib.sanerf_field_head_forward(None, None, None, 0.0, 0, enc_in.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                             None, h1.data_ptr() if train else None, h2.data_ptr() if train else None, out.data_ptr(), prec, st), "f")
def timeit(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
for prec in (0, 1):
    print(f"precision={prec}: MLP-only forward, inference {timeit(lambda: run(prec, False)):.1f} us, training (saves h1/h2) {timeit(lambda: run(prec, True)):.1f} us")
