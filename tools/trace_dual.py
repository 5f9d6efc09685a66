"""Cycle trace of the dual-slot forward kernel (library built with CSPN_TRACE=1): python tools/trace_dual.py [nyu|kitti1] [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cspn_monodepth_b200 import _lib
from tools.time_paths import CFGS, make
lib = _lib.load()
name = sys.argv[1] if len(sys.argv) > 1 else "nyu"
cfg = dict(CFGS[name])
if len(sys.argv) > 2:
    cfg["B"] = int(sys.argv[2])
dev = torch.device("cuda:0")
b, h, w, it = cfg["B"], cfg["H"], cfg["W"], cfg["iters"]
plan = _lib.forward_plan(b, 1, h, w, it, 3, cfg["mode"])
print(name, cfg["B"], plan)
assert plan["kernel"] == _lib.KERNEL_DUAL
sets = [make(cfg, i, dev) for i in range(3)]
sfx = "f32" if cfg["dtype"] == torch.float32 else "f16"
out = torch.empty_like(sets[0][1])
n = lib.cspn_fwd_workspace_bytes(b, 1, h, w, it, 3, cfg["mode"])
ws = torch.empty(n, dtype=torch.uint8, device=dev)
fn = getattr(lib, "cspn_fwd_" + sfx)
def step(i):
    g, d, s, _ = sets[i % 3]
    _lib.check(fn(g.data_ptr(), 8 * h * w, d.data_ptr(), s.data_ptr(), 1, out.data_ptr(), b, 1, h, w, it, 3, cfg["mode"], ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
nctas, nw, slots = plan["ctas"], 12, 128
trace = torch.zeros(nctas * nw * slots, dtype=torch.int64, device=dev)
step(0); step(1); torch.cuda.synchronize()
lib.cspn_debug_set_trace_dual.argtypes = [ctypes.c_void_p]
assert lib.cspn_debug_set_trace_dual(trace.data_ptr()) == 0
step(2); torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(nctas, nw, slots)[:, :8]      # compute warps
t0 = tr[:, :, 0].min(axis=1, keepdims=True)[:, :, None]
rel = (tr - t0).astype(np.float64)
rel[tr == 0] = np.nan
names = {0: "start", 1: "TMA issued", 2: "depth/sparse in registers", 3: "raw weights in registers", 4: "normalised", 5: "prologue done", 6: "loop done", 7: "stores issued"}
for k in sorted(names):
    print(f"  {names[k]:26s} med {np.nanmedian(rel[:, :, k]):9.0f}   max {np.nanmax(rel[:, :, k]):9.0f}")
ne = min(24, (it + 1) // 2)
P = rel[:, :, 8:8 + 4 * ne].reshape(nctas, 8, ne, 4)
prev_end = np.concatenate([rel[:, :, 5:6], P[:, :, :-1, 3]], axis=2)          # end of the slot's previous period
def med(x): return np.nanmedian(x)
sl = slice(1, ne - 1)
for name_, ws in (("slot A (warps 0-3)", slice(0, 4)), ("slot B (warps 4-7)", slice(4, 8))):
    print("  %s, per period (medians over CTAs, warps, periods 1..%d):" % (name_, ne - 2))
    d = (P[:, ws, :, 0] - prev_end[:, ws])[:, :, sl]
    print("    wait for + take halo %6.0f   (p10 %6.0f  p90 %6.0f)" % (med(d), np.nanpercentile(d, 10), np.nanpercentile(d, 90)))
    print("    even step            %6.0f" % med((P[:, ws, :, 1] - P[:, ws, :, 0])[:, :, sl]))
    print("    odd step             %6.0f" % med((P[:, ws, :, 2] - P[:, ws, :, 1])[:, :, sl]))
    print("    stage rim            %6.0f" % med((P[:, ws, :, 3] - P[:, ws, :, 2])[:, :, sl]))
    print("    whole period         %6.0f" % med((P[:, ws, :, 3] - prev_end[:, ws])[:, :, sl]))
print("  first period %.0f, CTA total median %.0f max %.0f" % (med(P[:, :, 0, 3] - rel[:, :, 5]), med(np.nanmax(rel[:, :, 7], axis=1)), np.nanmax(rel[:, :, 7])))
