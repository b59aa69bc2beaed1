"""cfg3: stage-2 SAM feature-field training step (64x64 rays, 256-d target map), one B200, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from nerf.network import NeRFNetwork
from sanerf_b200 import _lib
from sanerf_b200.train import SAMTrainer, default_opt
import bench

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = NeRFNetwork(default_opt(with_sam=True)).to(dev)
trainer = SAMTrainer(model)
N = 4096
o, d, _ = bench.synthetic_rays(N, dev, 1234)
target = torch.randn(1, 256, 64, 64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for _ in range(5):
    trainer.step(o, d, target, 64, 64)
torch.cuda.synchronize()
ts = []
for i in range(int(os.environ.get('SAM_STEPS', 20))):
    flush.fill_(float(i))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loss = trainer.step(o, d, target, 64, 64); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
print(f"SAM feature-field training step (stage 2, 4096 rays = 64x64, s_grid L16 F8 T2^19, samvit_mlp, fwd+bwd+Adam): "
      f"{ms:.3f} ms/step = {N / ms * 1e3 / 1e6:.2f} M rays/s, loss {float(loss):.4f}")
_lib.stats.reset("*", None)
trainer.step(o, d, target, 64, 64); torch.cuda.synchronize()
agg = {}
for s, e, info in _lib.stats.spans:
    agg.setdefault(info["name"], [0.0, 0]); agg[info["name"]][0] += s.elapsed_time(e) * 1e3; agg[info["name"]][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {v[0]:8.1f} us x{v[1]}  {k}")
