"""TMA-fed GEMM of the samvit chain (4096 x 256 x 256, bias + leaky ReLU) against the register-staged kernel: graph of 50
back-to-back launches, CUDA events, hot L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import fused
dev = torch.device("cuda", 0)
def bench(M, N, K, precision, reps=50):
    A, B, C, bias = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.zeros(M, N, device=dev), torch.randn(N, device=dev)
    one = lambda: fused.gemm_tc(A, B, C, M, N, K, bias=bias, act=True, precision=precision)
    one(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): one()
    ts = []
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return float(np.median(ts))
for K in (64, 128, 256, 512):
    print(f"4096 x 256 x {K}: TMA-fed {bench(4096, 256, K, 0):.2f} us   register-staged {bench(4096, 256, K, 0 | 128):.2f} us", flush=True)
