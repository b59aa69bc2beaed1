"""GPU probe: which descriptor encodings of the MN-major tf32 operand reproduce At^T @ Bt exactly (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import torch
from sanerf_b200 import _lib

lib = _lib.load()
def run(mode, M, N, K, A, B):
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.sanerf_umma_selftest(mode, M, N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(), _lib.current_stream(A.device))
    _lib.check(rc, "selftest"); torch.cuda.synchronize(); return D
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(128, 64, 64), (64, 32, 128), (64, 64, 128), (64, 16, 128), (64, 16, 8), (128, 32, 8), (128, 64, 128)]:
    At = torch.randint(-8, 9, (K, M), generator=g).float().cuda(); Bt = torch.randint(-8, 9, (K, N), generator=g).float().cuda()
    ref = At.t() @ Bt
    for variant in (0, 1):
        D = run(1 | (variant << 8), M, N, K, At, Bt)
        ok = torch.equal(D, ref)
        print(f"MN-major M={M} N={N} K={K} variant={variant}: exact={ok} nonzero={(D != 0).float().mean().item():.3f} "
              f"match_frac={(D == ref).float().mean().item():.3f}", flush=True)
