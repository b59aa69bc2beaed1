"""Multi-GPU check (torchrun, one rank per GPU).  Run by tests/test_gpu_multi.py when the box has >= 2 GPUs.

  1. one update from identical gradients and optimizer state: the fused symmetric-memory kernel (reduce + Adam + broadcast over
     NVLink, csrc/symm_adam.cu) == NCCL all-reduce + full Adam on every rank, for the deferred main-table range and the tail
     range; all ranks bit-identical; gradients cleared everywhere;
  2. N-rank gradients == one rank on the concatenated batch (SURVEY §4 tier iv), max relative L2 <= 1e-4;
  3. a few real steps (one CUDA graph per step and rank, no NCCL inside): ranks stay bit-identical, the two forms stay close;
  4. the NCCL reduce-scatter / all-gather form (SANERF_SYMM=0) against the same all-reduce checker;
  5. EMA / optimizer state gathered from the rank-sharded form == the unsharded one.
"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import torch, torch.distributed as dist
from nerf.network import NeRFNetwork
from sanerf_b200.train import RGBTrainer, default_opt
from sanerf_b200.step import FusedRGBStep
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model_a = NeRFNetwork(default_opt()).to(dev)
with torch.no_grad():
    for enc in [model_a.grid, *model_a.prop_encoders]:
        enc.embeddings.uniform_(-0.5, 0.5)
model_b, model_c = copy.deepcopy(model_a), copy.deepcopy(model_a)


def trainer(model, symm, sharded=True):
    os.environ["SANERF_SYMM"] = "1" if symm else "0"
    t = RGBTrainer(model, world_size=world, ema_decay=0.95)
    os.environ.pop("SANERF_SYMM")
    return t


ta, tb, tc = trainer(model_a, True), trainer(model_b, False), trainer(model_c, False)
assert ta.optimizer.symm is not None and tb.optimizer.symm is None
N = 2048
o, d, rgb = bench.synthetic_rays(N, dev, 100 + rank)
pa, pb, pc = ta.plan(N), tb.plan(N), tc.plan(N)
pa.perturb = pb.perturb = pc.perturb = False
pb.sharded_update = False                      # checker: all-reduce + full-size Adam
msg = [f"multicast={ta.optimizer.symm.multicast()}"]

# ---- 1. one update of both ranges from identical gradients / optimizer state
for t in (ta, tb, tc):
    t.optimizer.exp_avg.normal_(0, 1e-3, generator=torch.Generator(device=dev).manual_seed(7))
    t.optimizer.exp_avg_sq.uniform_(0, 1e-6, generator=torch.Generator(device=dev).manual_seed(8))
pa.gradients_only(o, d, rgb)
for t in (tb, tc):
    t.optimizer.flat_grad.copy_(ta.optimizer.flat_grad)
for t, p in ((ta, pa), (tb, pb), (tc, pc)):
    t.optimizer.schedule()                     # sets the gate and the step-1 terms
    p._update_main(); p._update_rest(True)
torch.cuda.synchronize()
ta.optimizer.symm.check()
for name, t in (("symm", ta), ("nccl-sharded", tc)):
    diff = (t.optimizer.flat_param - tb.optimizer.flat_param).abs().max().item()
    scale = tb.optimizer.flat_param.abs().max().item()
    assert diff <= 2e-6 * max(scale, 1.0), (name, diff)       # summation order of <= 8 terms differs (switch vs ring)
    assert float(t.optimizer.flat_grad.abs().max()) == 0.0, name
    ref = t.optimizer.flat_param.clone(); dist.broadcast(ref, 0)
    assert torch.equal(ref, t.optimizer.flat_param), name     # every rank holds bit-identical parameters
    msg.append(f"{name}: one update max abs diff {diff:.1e}")
# ---- 5. sharded state gathered == unsharded
ta.optimizer.gather_sharded_state(); tc.optimizer.gather_sharded_state()
for buf in ("exp_avg", "exp_avg_sq", "ema"):
    for name, t in (("symm", ta), ("nccl-sharded", tc)):
        x, y = getattr(t.optimizer, buf), getattr(tb.optimizer, buf)
        assert (x - y).abs().max().item() <= 1e-6 * max(1.0, y.abs().max().item()), (name, buf)

# ---- 2. N-rank gradient == one rank on the concatenated batch
ta.optimizer.zero_grad()
pa.gradients_only(o, d, rgb)
multi = ta.optimizer.flat_grad.clone(); dist.all_reduce(multi); multi /= world
ta.optimizer.zero_grad()
parts = [[torch.empty_like(t) for _ in range(world)] for t in (o, d, rgb)]
for lst, t in zip(parts, (o, d, rgb)):
    dist.all_gather(lst, t.contiguous())
O, D, RGB = (torch.cat(lst, 0) for lst in parts)
big = FusedRGBStep(model_a, ta.optimizer, world * N, world_size=1, use_graph=False, perturb=False)
big.gradients_only(O, D, RGB)
single = ta.optimizer.flat_grad.clone(); ta.optimizer.zero_grad()
worst = 0.0
for n_, p in model_a.named_parameters():
    a, _ = ta.optimizer.ranges[id(p)]; k = p.numel()
    rel = ((multi[a:a + k].double() - single[a:a + k].double()).norm() / single[a:a + k].double().norm()).item()
    worst = max(worst, rel)
    assert rel <= 1e-4, (n_, rel)
msg.append(f"N-rank vs concatenated-batch gradient: max rel L2 {worst:.1e}")

# ---- 3. real steps: one graph per step with the fused update vs the eager NCCL all-reduce form
for i in range(6):
    la, lb = float(ta.step(o, d, rgb)), float(tb.step(o, d, rgb))
ta.flush(); tb.flush(); torch.cuda.synchronize()
ta.optimizer.symm.check()
assert (True, True) in pa.graphs and len(pa.graphs[(True, True)]) == 1       # ONE graph, no NCCL in the step
worst = 0.0
for (n_, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
    rel = ((p - q).norm() / q.norm().clamp_min(1e-12)).item()
    ref = p.detach().clone(); dist.broadcast(ref, 0)
    assert (p - ref).abs().max().item() == 0.0, n_             # ranks bit-identical
    worst = max(worst, rel)
    assert rel < 5e-3, (n_, rel)                               # atomic-order noise through Adam's sign-like first updates
msg.append(f"6 steps: ranks bit-identical, fused vs NCCL forms differ by {worst:.1e} relative; losses {la:.5f} {lb:.5f}")
if rank == 0:
    print(f"OK world={world}: " + "; ".join(msg))
dist.destroy_process_group()
