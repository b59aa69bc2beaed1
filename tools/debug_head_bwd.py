import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2 + 5
torch.manual_seed(1)
w1 = (torch.randn(64, 32) / 32 ** 0.5).cuda(); w2 = (torch.randn(64, 64) / 8).cuda(); w3 = (torch.randn(16, 64) / 8).cuda()
g = torch.Generator(device="cuda").manual_seed(3)
e = torch.randn(B, 32, device="cuda", generator=g); g_out = torch.randn(B, 16, device="cuda", generator=g)
e64 = e.double().requires_grad_(True)
ws = [w.double().requires_grad_(True) for w in (w1, w2, w3)]
h1r = torch.relu(e64 @ ws[0].t()); h2r = torch.relu(h1r @ ws[1].t()); o = h2r @ ws[2].t()
(o * g_out.double()).sum().backward()
from sanerf_b200.fused import rows_to_tcm
h1 = rows_to_tcm(h1r.detach().float().contiguous()); h2 = rows_to_tcm(h2r.detach().float().contiguous()); e_rows = e; e = rows_to_tcm(e)
for prec in (0, 1):
    g_enc = torch.full((B, 32), float("nan"), device="cuda")
    gw = [torch.zeros_like(w) for w in (w1, w2, w3)]
    rc = lib.sanerf_field_head_backward(e.data_ptr(), h1.data_ptr(), h2.data_ptr(), g_out.data_ptr(), w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), B,
                                        g_enc.data_ptr(), None, None, 0.0, 0, None, gw[0].data_ptr(), gw[1].data_ptr(), gw[2].data_ptr(), prec, _lib.current_stream(e.device))
    _lib.check(rc, "bwd"); torch.cuda.synchronize()
    err = (g_enc.double() - e64.grad).abs().max(dim=1).values / e64.grad.abs().max()
    tiles = (B + 127) // 128
    pad = torch.zeros(tiles * 128, device="cuda", dtype=torch.float64); pad[:B] = err
    per_tile = pad.view(tiles, 128).max(dim=1).values.cpu().numpy()
    bad = np.nonzero(per_tile > (1e-5 if prec == 0 else 1e-2))[0]
    print(f"prec={prec} B={B} tiles={tiles} max_err={err.max().item():.3e} bad tiles: {len(bad)} first: {bad[:20]} (mod 148: {bad[:20] % 148}) it: {bad[:20] // 148}")
    if len(bad):
        t = bad[0]; rows = np.nonzero(pad.view(tiles, 128)[t].cpu().numpy() > 1e-5)[0]
        print("  bad rows in first bad tile:", rows[:40], len(rows))
    for got, ref, name in zip(gw, ws, ("w1", "w2", "w3")):
        print(f"   d{name} rel err {((got.double() - ref.grad).abs().max() / ref.grad.abs().max()).item():.3e}")
