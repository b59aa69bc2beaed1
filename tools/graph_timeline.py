"""Timeline of ONE replay of the captured RGB (or SAM) training step: when every kernel's first block really started
(%globaltimer stamped behind griddepcontrol.wait; needs the -DSANERF_HEAD_TRACE build: SANERF_LIB_PATH=.../lib_trace/...).
usage: python tools/graph_timeline.py [rgb|sam] [warm-up steps]"""
import ctypes, glob, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segment-anything-nerf_b200")]
import numpy as np, torch
from sanerf_b200 import _lib
import bench

def file_tag(name):
    h = 0
    for ch in name.encode():
        h = (h * 31 + ch) & 0xffffffff
    return h & 63

def kernel_ids():
    ids = {}
    for path in glob.glob(os.path.join(ROOT, "segment-anything-nerf_b200", "csrc", "*.cu")):
        kern = "?"
        for ln, line in enumerate(open(path), 1):
            m = re.search(r"__global__.*?(\w+)\s*\(", line)
            if m: kern = m.group(1)
            elif "__global__" in line: kern = "?"
            if kern == "?" and re.match(r"\s*(\w+)\(", line) and "__global__" not in line:
                pass
            if "pdl_begin();" in line:
                ids[ln * 64 + file_tag(os.path.basename(path))] = kern
            m2 = re.match(r"\s*(?:void\s+)?(\w+)\(const", line)
    return ids

def kernel_ids2():
    """second pass for kernels whose name sits on the line after `__global__ void __launch_bounds__(..)`"""
    ids = {}
    for path in glob.glob(os.path.join(ROOT, "segment-anything-nerf_b200", "csrc", "*.cu")):
        text = open(path).read().split("\n")
        last = "?"
        for i, line in enumerate(text):
            if "__global__" in line:
                joined = " ".join(text[i:i + 3])
                m = re.search(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*|__maxnreg__\([^)]*\)\s*)*(\w+)\s*\(", joined)
                if m: last = m.group(1)
            if "pdl_begin();" in line:
                ids[(i + 1) * 64 + file_tag(os.path.basename(path))] = last
    return ids

what = sys.argv[1] if len(sys.argv) > 1 else "rgb"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lib = _lib.load()
lib.sanerf_debug_stamps.argtypes = [ctypes.c_int, ctypes.c_void_p]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if what == "rgb":
    from nerf.network import NeRFNetwork
    from sanerf_b200.train import RGBTrainer, default_opt
    model = NeRFNetwork(default_opt()).to(dev)
    trainer = RGBTrainer(model)
    o, d, rgb = bench.synthetic_rays(bench.N_RAYS, dev, 1234)
    step = lambda: trainer.step(o, d, rgb)
else:
    raise SystemExit("only rgb for now")
for _ in range(warm): step()
torch.cuda.synchronize()
assert lib.sanerf_debug_stamps(0, None) == 0
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
names = kernel_ids2()
class Buf(ctypes.Structure):
    _fields_ = [("count", ctypes.c_uint), ("pad", ctypes.c_uint), ("s", ctypes.c_ulonglong * (2 * 4096))]
for rep in range(3):
    flush.fill_(float(rep)); torch.cuda.synchronize()
    assert lib.sanerf_debug_stamps(1, None) == 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); torch.cuda.synchronize()
    buf = Buf(); assert lib.sanerf_debug_stamps(2, ctypes.byref(buf)) == 0
    rows = sorted((buf.s[2 * i], buf.s[2 * i + 1] & 0xffffffff) for i in range(min(buf.count, 4096)))
    t0 = rows[0][0]
    print(f"--- replay {rep}: {a.elapsed_time(b) * 1e3:.1f} us by events, {buf.count} launches")
    prev = t0
    for t, kid in rows:
        print(f"{(t - t0) / 1e3:8.1f} us  (+{(t - prev) / 1e3:6.1f})  {names.get(kid, hex(kid))}")
        prev = t
