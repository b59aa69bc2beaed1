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
import time
DBG = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
for prec, dbg in [(pp, d) for d in DBG for pp in (0, 1)]:
    lib.sanerf_debug_head_flags(dbg)
    evs = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                            g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), prec, _lib.current_stream(e.device))
        b.record(); _lib.check(rc, "bwd"); torch.cuda.synchronize(); evs.append(a.elapsed_time(b) * 1e3)
    print(f"dbg={dbg} prec={prec} kernel us {evs[1:]}")
    out = (ctypes.c_longlong * 128)()
    lib.sanerf_debug_head_trace.argtypes = [ctypes.c_void_p]
    assert lib.sanerf_debug_head_trace(out) == 0
    mk = (ctypes.c_longlong * 8)()
    lib.sanerf_debug_head_marks.argtypes = [ctypes.c_void_p]
    if lib.sanerf_debug_head_marks(mk) == 0:
        m = list(mk)
        print(f"marks (clk): pdl wait {m[1]-m[0]}  prologue {m[2]-m[1]}  tiles {m[3]-m[2]}  weight-gradient write-out {m[4]-m[3]}")
    lock = os.environ.get("SANERF_HEAD_BWD_LOCKSTEP", "0") != "0"
    names = (["top", "staged", "pub1", "mma3", "E3", "pub2", "mma4", "E4", "pub3", "mma5", "E5"] if lock else
             ["top", "G3>DG2", "w(dW1)", "H2>dW3", "w(DG2)", "E3a>DG1", "w(dW3)", "E3b>dW2", "w(DG1)", "E4a>DGE", "w(dW2)", "E4b>dW1", "w(DGE)", "E5"])
    raw = np.array(out[:]).reshape(2, 4, 16)
    if not lock:
        for it in range(1, 3):
            print(f"tile#{it}: E4b stores {int(raw[0, it, 14] - raw[0, it, 10])}  publish {int(raw[0, it, 15] - raw[0, it, 14])}  load issue {int(raw[0, it, 11] - raw[0, it, 15])} clk")
    t = raw[:, :, :len(names)]
    for it in range(1, 3):
        print("abs owner ", [int(x - t[0, it, 0]) for x in t[0, it]])
        print("abs loader", [int(x - t[0, it, 0]) for x in t[1, it]])
    for who, nm in ((0, "t0"), (1, "t128")):
        for it in range(1, 3):
            d = np.diff(t[who, it]); base = t[who, it, 0] - t[0, it, 0]
            print(f"prec={prec} {nm} tile#{it} start+{base}: " + " ".join(f"{n}:{int(x)}" for n, x in zip(names[1:], d)) + f" total={int(t[who, it, -1] - t[who, it, 0])}")
