"""gemm_tc micro-benchmark: fixed cost vs per-K-chunk cost (graph of 50 back-to-back launches, CUDA events, hot L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import fused
dev = torch.device("cuda", 0)
M = 4096
def bench(N, K, a_trans=False, b_trans=False, epilogue=0, k_splits=1, reps=50, MM=M, precision=0, colsum=True):
    A = torch.randn((K, MM) if a_trans else (MM, K), device=dev)
    B = torch.randn((K, N) if b_trans else (N, K), device=dev)
    C = torch.zeros(MM, N, device=dev)
    mask = torch.randn(MM, N, device=dev) if epilogue == 1 else None
    bias = torch.randn(N, device=dev)
    cs = torch.zeros(N, device=dev) if (epilogue == 1 and colsum) else None
    def one():
        fused.gemm_tc(A, B, C, MM, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=epilogue, k_splits=k_splits, bias=bias if epilogue == 0 else None,
                      act=epilogue == 0, mask=mask, mask_cols=N if epilogue == 1 else 0, colsum=cs, precision=precision)
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
for name, dbg in (("no MMA", 1), ("no shared stores", 2), ("no global loads", 4), ("no MMA, no stores", 3), ("nothing but barriers", 7), ("1xTF32", None)):
    pr = 1 if dbg is None else dbg << 4
    print(f"diagnostic [{name:22s}] K=256: {bench(256, 256, precision=pr):6.2f}  K=1024: {bench(256, 1024, precision=pr):6.2f} us/launch")
for K in (32, 256, 1024):
    print(f"forward-like  4096 x 256 x K={K:4d}: {bench(256, K):6.2f} us/launch")
print(f"e0 b_trans 4096x256x256           : {bench(256, 256, b_trans=True):6.2f}")
print(f"e1 (mask) no colsum, B normal     : {bench(256, 256, epilogue=1, colsum=False):6.2f}")
print(f"e1 (mask) + colsum, B normal      : {bench(256, 256, epilogue=1):6.2f}")
print(f"K=163 (unaligned rows)            : {bench(256, 163):6.2f}")
print(f"K=419 (unaligned rows)            : {bench(256, 419):6.2f}")
print(f"dX  e1 4096x256x256 b_trans       : {bench(256, 256, b_trans=True, epilogue=1):6.2f}")
print(f"dX  e1 4096x419x256 b_trans       : {bench(419, 256, b_trans=True, epilogue=1):6.2f}")
print(f"dW  e2 256x256x4096 trans/trans x16: {bench(256, 4096, a_trans=True, b_trans=True, epilogue=2, k_splits=16, MM=256):6.2f}")
print(f"dW  e2 256x419x4096 trans/trans x16: {bench(419, 4096, a_trans=True, b_trans=True, epilogue=2, k_splits=16, MM=256):6.2f}")
x = torch.randn(4096, 256, device=dev); w = torch.randn(256, 256, device=dev)
g = torch.cuda.CUDAGraph(); y = x @ w.t(); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(50): y = torch.nn.functional.linear(x, w)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
print(f"cuBLAS fp32 F.linear 4096x256x256  : {a.elapsed_time(b) * 1e3 / 50:6.2f} us/launch")
