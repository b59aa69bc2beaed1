import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
lib = _lib.load()
B = 262144
w1 = (torch.randn(64, 32) / 32 ** 0.5).cuda(); w2 = (torch.randn(64, 64) / 8).cuda(); w3 = (torch.randn(16, 64) / 8).cuda()
e = torch.randn(B, 32, device="cuda"); h1 = torch.randn(B, 64, device="cuda").relu(); h2 = torch.randn(B, 64, device="cuda").relu()
g_out = torch.randn(B, 16, device="cuda"); g_enc = torch.empty(B, 32, device="cuda")
gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
for prec in (0, 1):
    for _ in range(3):
        rc = lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                            g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), prec, _lib.current_stream(e.device))
        _lib.check(rc, "bwd"); torch.cuda.synchronize()
    out = (ctypes.c_longlong * 88)()
    lib.sanerf_debug_head_trace.argtypes = [ctypes.c_void_p]
    assert lib.sanerf_debug_head_trace(out) == 0
    t = np.array(out[:]).reshape(2, 4, 11)
    names = ["top", "staged", "pub1", "mma3", "E3", "pub2", "mma4", "E4", "pub3", "mma5", "E5"]
    for it in range(1, 3):
        print("abs owner ", [int(x - t[0, it, 0]) for x in t[0, it]])
        print("abs loader", [int(x - t[0, it, 0]) for x in t[1, it]])
    for who, nm in ((0, "t0"), (1, "t128")):
        for it in range(1, 3):
            d = np.diff(t[who, it]); base = t[who, it, 0] - t[0, it, 0]
            print(f"prec={prec} {nm} tile#{it} start+{base}: " + " ".join(f"{n}:{int(x)}" for n, x in zip(names[1:], d)) + f" total={int(t[who, it, -1] - t[who, it, 0])}")
