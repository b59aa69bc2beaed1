"""Marks of CTA 0 / thread 0 of head_forward_kernel (trace build, SANERF_LIB_PATH): where the fixed cost of a launch goes."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
from gridencoder import GridEncoder
lib = _lib.load()
lib.sanerf_debug_head_marks.argtypes = [ctypes.c_void_p]
enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096).cuda()
w1 = (torch.randn(64, 32) / 32 ** 0.5).cuda(); w2 = (torch.randn(64, 64) / 8).cuda(); w3 = (torch.randn(16, 64) / 8).cuda()
S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for B in (18944, 262144):
    x01 = torch.rand(B, 3, device="cuda"); out = torch.empty(B, 16, device="cuda")
    for cold in (True, False):
        for i in range(3):
            if cold: flush.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.sanerf_field_head_forward(x01.data_ptr(), enc.embeddings.data_ptr(), enc.offsets.data_ptr(), S, H, None, w1.data_ptr(), w2.data_ptr(),
                                               w3.data_ptr(), B, None, None, None, out.data_ptr(), 0, _lib.current_stream(x01.device))
            b.record(); _lib.check(rc, "fwd"); torch.cuda.synchronize()
        mk = (ctypes.c_longlong * 8)(); assert lib.sanerf_debug_head_marks(mk) == 0
        m = list(mk)
        print(f"B={B} {'cold' if cold else 'warm'}: kernel {a.elapsed_time(b) * 1e3:.1f} us; clocks: pdl wait {m[1]-m[0]}, prologue {m[2]-m[1]}, first encoding ready +{m[3]-m[2]}, "
              f"first tile done +{m[4]-m[3]}, loop end +{m[5]-m[4]}, exit +{m[6]-m[5]}; total {m[6]-m[0]}")
