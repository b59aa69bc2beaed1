"""Adam pass over a large flat range, alone on the GPU (CUDA events, L2 flushed): register-staged kernel vs TMA-fed kernel
(SANERF_ADAM_TMA=0/1 in the environment selects)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import fused
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5258512 * 8
p = torch.randn(n, device="cuda").requires_grad_(True)
opt = fused.FusedAdam([p], lr=1e-2, eps=1e-15)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
opt.schedule()
ts = []
for i in range(8):
    p.grad.normal_()
    flush.fill_(float(i))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); opt.apply(0, n); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
us = float(np.median(ts[2:]))
print(f"SANERF_ADAM_TMA={os.environ.get('SANERF_ADAM_TMA', '1')} n={n}: {us:.1f} us = {32.0 * n / us / 1e6:.2f} TB/s (32 B per parameter)")
