"""Warm per-kernel durations and start times of one eager training step (CUDA events on the launching streams)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import torch
from nerf.network import NeRFNetwork
from sanerf_b200 import _lib
from sanerf_b200.train import RGBTrainer, default_opt
import bench

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = NeRFNetwork(default_opt()).to(dev)
trainer = RGBTrainer(model, use_graph=False)
o, d, rgb = bench.synthetic_rays(bench.N_RAYS, dev, 1234)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):      # warm-up steps (training progress changes the timings)
    trainer.step(o, d, rgb)
torch.cuda.synchronize()
flush.fill_(1.0)
_lib.stats.reset("*", None)
t0 = torch.cuda.Event(enable_timing=True); t0.record()
trainer.step(o, d, rgb)
t1 = torch.cuda.Event(enable_timing=True); t1.record()
torch.cuda.synchronize()
rows = [(t0.elapsed_time(s) * 1e3, s.elapsed_time(e) * 1e3, info) for s, e, info in _lib.stats.spans]
tot = 0.0
for start, dur, info in sorted(rows, key=lambda r: r[0]):
    tot += dur
    print(f"{start:8.1f} +{dur:7.1f} us  {info['name']} { {k: v for k, v in info.items() if k != 'name'} }")
print(f"sum of kernel durations {tot:.1f} us; eager step wall {t0.elapsed_time(t1) * 1e3:.1f} us")
