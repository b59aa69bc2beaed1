"""Multi-GPU check (torchrun, one rank per GPU): the sharded main-table update (reduce-scatter + Adam shard + all-gather)
gives every rank the same parameters as plain all-reduce + full Adam, and all ranks agree."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import torch, torch.distributed as dist
from nerf.network import NeRFNetwork
from sanerf_b200.train import RGBTrainer, default_opt
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model_a = NeRFNetwork(default_opt()).to(dev)
with torch.no_grad():
    for enc in [model_a.grid, *model_a.prop_encoders]:
        enc.embeddings.uniform_(-0.5, 0.5)
model_b = copy.deepcopy(model_a)
ta, tb = RGBTrainer(model_a, world_size=world), RGBTrainer(model_b, world_size=world)
N = 2048
o, d, rgb = bench.synthetic_rays(N, dev, 100 + rank)
pa, pb = ta.plan(N), tb.plan(N)
pa.perturb = pb.perturb = False
pb.sharded_update = False
# ---- 1. one update from identical gradients and optimizer state: the two forms must agree (bit-exact at world = 2,
#         where a two-term sum has one order only)
for t in (ta, tb):
    t.optimizer.exp_avg.normal_(0, 1e-3, generator=torch.Generator(device=dev).manual_seed(7))
    t.optimizer.exp_avg_sq.uniform_(0, 1e-6, generator=torch.Generator(device=dev).manual_seed(8))
pa.gradients_only(o, d, rgb)
tb.optimizer.flat_grad.copy_(ta.optimizer.flat_grad)
pa._update_main(); pb._update_main(); torch.cuda.synchronize()
a0, b0 = pa._main_range()
pa_, pb_ = ta.optimizer.flat_param[a0:b0], tb.optimizer.flat_param[a0:b0]
diff = (pa_ - pb_).abs().max().item()
assert (diff == 0.0) if world == 2 else (diff < 1e-6), diff
assert float(ta.optimizer.flat_grad[a0:b0].abs().max()) == 0.0 and float(tb.optimizer.flat_grad[a0:b0].abs().max()) == 0.0
ta.optimizer.flat_grad.zero_(); tb.optimizer.flat_grad.zero_()
# ---- 2. a few real steps: ranks stay bit-identical; the two forms stay close (atomic-order noise through Adam's
#         sign-like first updates bounds how close two runs of even the SAME form can be)
for i in range(6):
    la, lb = float(ta.step(o, d, rgb)), float(tb.step(o, d, rgb))
ta.flush(); tb.flush(); torch.cuda.synchronize()
worst = 0.0
for (n, p), (_, q) in zip(model_a.named_parameters(), model_b.named_parameters()):
    rel = ((p - q).norm() / q.norm().clamp_min(1e-12)).item()
    ref = p.detach().clone(); dist.broadcast(ref, 0)
    across = (p - ref).abs().max().item()
    worst = max(worst, rel)
    assert across == 0.0, (n, across)          # every rank holds bit-identical parameters
    assert rel < 5e-3, (n, rel)
if rank == 0:
    print(f"OK world={world}: one sharded update == all-reduce update (max abs diff {diff:.1e}); after 6 steps ranks are "
          f"bit-identical and the two forms differ by {worst:.1e} relative; losses {la:.5f} {lb:.5f}")
dist.destroy_process_group()
