"""cfg4: interactive frame, 512x512 RGB + depth and a 64x64x256 SAM feature map, one B200 (CUDA events, 10 frames)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from nerf.network import NeRFNetwork
from sanerf_b200.train import default_opt, render_frame

dev = torch.device("cuda", 0)
torch.manual_seed(0)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = NeRFNetwork(default_opt(with_sam=True, max_ray_batch=chunk)).to(dev).eval()

def rays(h, w):
    fy = 0.5 * h / np.tan(np.radians(30)); fx = fy
    j, i = torch.meshgrid(torch.arange(h, device=dev) + 0.5, torch.arange(w, device=dev) + 0.5, indexing="ij")
    d = torch.stack([(i - w / 2) / fx, (j - h / 2) / fy, torch.ones_like(i)], -1).reshape(-1, 3)      # un-normalised pinhole dirs
    o = torch.tensor([0.1, 0.0, -0.4], device=dev).expand_as(d).contiguous()
    return o, d.contiguous()

o, d = rays(512, 512)
fo, fd = rays(64, 64)
fused = "--staged" not in sys.argv
def frame(with_feats=True):
    return render_frame(model, o, d, fo if with_feats else None, fd if with_feats else None, 64, 64, fused=fused)
print("RGB pass:", "hand-scheduled forward, one CUDA graph" if fused else f"staged renderer, chunks of {chunk}")
for with_feats in (False, True):
    for _ in range(3): frame(with_feats)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = frame(with_feats); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(f"chunk={chunk} 512x512 RGB+depth{' + 64x64x256 SAM features' if with_feats else ''}: {ms:.2f} ms/frame = {1e3 / ms:.1f} FPS "
          f"(image {tuple(out['image'].shape)}{', samvit ' + str(tuple(out['samvit'].shape)) if with_feats else ''})", flush=True)
