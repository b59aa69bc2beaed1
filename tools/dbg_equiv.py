import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import torch
from nerf.network import NeRFNetwork
from sanerf_b200.train import RGBTrainer, default_opt
from sanerf_b200.step import FusedRGBStep
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = NeRFNetwork(default_opt()).to(dev)
tr = RGBTrainer(model, ema_decay=0.95)
sets = [bench.synthetic_rays(8192, dev, 1234 + i * 100) for i in range(2)]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
for i in range(steps):
    tr.step(*sets[i % 2])
tr.flush()
opt = tr.optimizer
def grads(plan, o, d, rgb):
    opt.zero_grad()
    plan.gradients_only(o, d, rgb, update_proposal=True)
    g = opt.flat_grad.clone(); opt.zero_grad(); return g
plan = tr.plan(8192); plan.perturb = False
g0a, g0b = grads(plan, *sets[0]), grads(plan, *sets[0])
g1 = grads(plan, *sets[1])
multi = (g0a + g1) / 2
O, D, RGB = (torch.cat([a, b], 0) for a, b in zip(sets[0], sets[1]))
big = FusedRGBStep(model, opt, 16384, use_graph=False, perturb=False)
single = grads(big, O, D, RGB)
single2 = grads(big, O, D, RGB)
# autograd path on the concatenated batch
tr2 = RGBTrainer.__new__(RGBTrainer); 
opt.zero_grad()
loss, _ = tr.loss(O, D, RGB, update_proposal=True, perturb=False); loss.backward()
auto = opt.flat_grad.clone(); opt.zero_grad()
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
for n_, p in model.named_parameters():
    a, _ = opt.ranges[id(p)]; k = p.numel(); s = slice(a, a + k)
    print(f"{n_:32s} |g|={single[s].norm().item():.3e} multi-vs-single {rel(multi[s], single[s]):.2e}  single-vs-auto {rel(single[s], auto[s]):.2e} "
          f"repeat8192 {rel(g0a[s], g0b[s]):.2e} repeat16384 {rel(single[s], single2[s]):.2e} multi-vs-auto {rel(multi[s], auto[s]):.2e}")
