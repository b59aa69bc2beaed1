"""Phase timing of the hand-scheduled stage-2 step: each phase captured as its own CUDA graph, CUDA events around replays."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from nerf.network import NeRFNetwork
from sanerf_b200.train import SAMTrainer, default_opt
import bench

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = NeRFNetwork(default_opt(with_sam=True)).to(dev)
trainer = SAMTrainer(model)
N = 4096
o, d, _ = bench.synthetic_rays(N, dev, 1234)
target = torch.randn(1, 256, 64, 64, device=dev)
for _ in range(3):
    trainer.step(o, d, target, 64, 64)
trainer.flush()
plan = trainer.plan(N, 64, 64, tuple(target.shape))
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def graph_of(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g

from sanerf_b200.step import _critical
def front_with_adam():
    def body():
        main, upd = plan._deferred_update()
        plan._launch_front()
        main.wait_stream(upd)
    _critical(plan, body)

phases = {
    "front (frozen stage-1 forward) alone": graph_of(plan._launch_front),
    "s_grid Adam alone": graph_of(plan._update_main),
    "front || s_grid Adam": graph_of(front_with_adam),
    "back (ray features fwd, samvit head fwd+bwd, LN+MSE, scatter)": graph_of(plan._launch_back),
    "small Adam": graph_of(plan._update_rest),
    "whole step": graph_of(plan._whole_step),
}
for name, g in phases.items():
    ts = []
    for i in range(8):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    print(f"{np.median(ts):8.1f} us  {name}")
# back phase without the L2 flush (activations hot)
from sanerf_b200 import _lib
_lib.stats.reset("*", None)
plan._launch_front(); plan._launch_back(); torch.cuda.synchronize()
agg = {}
for s, e, info in _lib.stats.spans:
    k = info["name"] + (f" e{info['epilogue']} {info['M']}x{info['N']}x{info['K']}" if info["name"] == "gemm_tc" else "")
    agg.setdefault(k, [0.0, 0]); agg[k][0] += s.elapsed_time(e) * 1e3; agg[k][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {v[0]:8.1f} us x{v[1]}  {k}  (eager, events around the launch)")
