"""Cycle trace of the fused kernel (needs a library built with CSPN_TRACE=1): python tools/trace_nyu.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cspn_monodepth_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(bench.NYU, B=B)
dev = torch.device("cuda:0")
sets = [[t.to(dev) for t in bench.synth(cfg, i)] for i in range(2)]
step = bench.run_module(cfg, sets)
nctas, nw, slots = B * 15, 8, 96
trace = torch.zeros(nctas * nw * slots, dtype=torch.int64, device=dev)
with torch.no_grad():
    step(0); step(1)
    torch.cuda.synchronize()
    lib.cspn_debug_set_trace.argtypes = [ctypes.c_void_p]
    assert lib.cspn_debug_set_trace(trace.data_ptr()) == 0
    step(0)
    torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(nctas, nw, slots)
t0 = tr[:, :, 0].min(axis=1, keepdims=True)[:, :, None]          # per-CTA start (clock64 is per SM)
rel = tr - t0
names = {0: "start", 1: "bar init+TMA issued", 2: "depth/sparse loaded", 10: "last TMA box consumed", 11: "raw weights in regs", 12: "normalised", 13: "cluster ready", 14: "loop done", 15: "stores issued"}
print(f"B={B}: {nctas} CTAs; medians over CTAs of warp-0 stamps (cycles since CTA start), and max over warps")
for k in sorted(names):
    print(f"  {names[k]:26s} med {np.median(rel[:, 0, k]):9.0f}   max {rel[:, :, k].max():9.0f}")
ex = rel[:, :, 17:17 + 72:3] - rel[:, :, 16:16 + 72:3]      # exchange_rows (STS+barrier+LDS) per step
cp = np.diff(rel[:, :, 16:16 + 72:3], axis=2)               # full step time
print("  per-step: exchange_rows (sts+bar+lds) median %.0f cyc; full step median %.0f cyc (even %.0f / odd %.0f)" % (
    np.median(ex), np.median(cp), np.median(cp[:, :, 1::2]), np.median(cp[:, :, 0::2])))
wait = rel[:, :, 18:18 + 72:3] - rel[:, :, 17:17 + 72:3]
print("  refresh wait+apply (even steps t>=2) median %.0f cyc, p90 %.0f" % (np.median(wait[:, :, 2::2]), np.percentile(wait[:, :, 2::2], 90)))
comp = rel[:, :, 19:16 + 72:3] - rel[:, :, 18:18 + 69:3]
print("  compute phase median %.0f cyc (even %.0f, odd/push %.0f)" % (np.median(comp), np.median(comp[:, :, 0::2]), np.median(comp[:, :, 1::2])))
print("  CTA total (max over warps) median %.0f, max %.0f cycles" % (np.median(rel[:, :, 15].max(axis=1)), rel[:, :, 15].max()))

# last push step: stamps 89 (before compute), 90 (sweep done), 91 (fence done), 92 (bulk copies issued)
d = rel[:, :, 89:93]
print("  push step (last one): sweep %.0f, fence %.0f, issue bulk %.0f cycles (medians over warps)" % (
    np.median(d[:, :, 1] - d[:, :, 0]), np.median(d[:, :, 2] - d[:, :, 1]), np.median(d[:, :, 3] - d[:, :, 2])))
for w in range(8):
    print("    warp %d: sweep %.0f fence %.0f bulk %.0f" % (w, np.median(d[:, w, 1] - d[:, w, 0]), np.median(d[:, w, 2] - d[:, w, 1]), np.median(d[:, w, 3] - d[:, w, 2])))
print("  per warp: exchange_rows even/odd | refresh wait | compute even/odd (medians over CTAs and steps)")
for w in range(8):
    print("    warp %d: %5.0f %5.0f | %5.0f | %5.0f %5.0f" % (w, np.median(ex[:, w, 2::2]), np.median(ex[:, w, 1::2]), np.median(wait[:, w, 2::2]),
                                                          np.median(comp[:, w, 0::2]), np.median(comp[:, w, 1::2])))
ctas_by_row = {r: [c for c in range(nctas) if (c // 5) % 3 == r] for r in range(3)}
for r, cs in ctas_by_row.items():
    print("  tile row %d: refresh wait median %.0f (warp 0 %.0f, warp 7 %.0f)" % (r, np.median(wait[cs][:, :, 2::2]), np.median(wait[cs][:, 0, 2::2]), np.median(wait[cs][:, 7, 2::2])))
print("  last refresh (t = 22), per warp, medians over CTAs: after row exchange -> polls done | -> mbarrier done | -> applied")
for w in range(8):
    a, b, c, d = rel[:, w, 17 + 66], rel[:, w, 93], rel[:, w, 94], rel[:, w, 18 + 66]
    print("    warp %d: %5.0f | %5.0f | %5.0f     (row exchange itself %5.0f)" % (w, np.median(b - a), np.median(c - b), np.median(d - c), np.median(a - rel[:, w, 16 + 66])))
print("  re-polls per warp over the whole launch (median / max over CTAs): " + "  ".join("%d/%d" % (np.median(tr[:, w, 95]), tr[:, w, 95].max()) for w in range(8)))
