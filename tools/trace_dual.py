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
full = trace.cpu().numpy().reshape(nctas, 12, slots).astype(np.float64)
full[full == 0] = np.nan
ne = min(12, (it + 1) // 2)
C = full[:, :8, 8:8 + 4 * ne].reshape(nctas, 8, ne, 4)          # compute warps, SM clock: [halo taken, even done, odd done, shipped]
G = full[:, :8, 64:64 + 4 * ne].reshape(nctas, 8, ne, 4)        # same points, global ns
CC = full[:, 8:, 8:8 + 4 * ne].reshape(nctas, 4, ne, 4)         # communication warps (index = e1): [poll start, detected, arrived]
GC = full[:, 8:, 64:64 + 4 * ne].reshape(nctas, 4, ne, 4)
def med(x): return np.nanmedian(x)
sl = slice(1, ne - 1)
print("  compute warps, per period (SM cycles, medians over CTAs / warps / periods 1..%d):" % (ne - 2))
prev_ship = C[:, :, :-1, 3]
print("    ship -> halo taken (wait)   %6.0f  (p10 %6.0f p90 %6.0f)" % (med(C[:, :, 1:, 0] - prev_ship), np.nanpercentile(C[:, :, 1:, 0] - prev_ship, 10), np.nanpercentile(C[:, :, 1:, 0] - prev_ship, 90)))
print("    even step                   %6.0f" % med((C[:, :, :, 1] - C[:, :, :, 0])[:, :, sl]))
print("    odd step                    %6.0f" % med((C[:, :, :, 2] - C[:, :, :, 1])[:, :, sl]))
print("    ship                        %6.0f" % med((C[:, :, :, 3] - C[:, :, :, 2])[:, :, sl]))
print("    whole period                %6.0f" % med(np.diff(C[:, :, :, 3], axis=2)))
print("  communication warps (SM cycles): poll start -> detected %6.0f (p10 %6.0f p90 %6.0f), detected -> arrived %6.0f" % (
    med(CC[:, :, 1:, 1] - CC[:, :, 1:, 0]), np.nanpercentile(CC[:, :, 1:, 1] - CC[:, :, 1:, 0], 10), np.nanpercentile(CC[:, :, 1:, 1] - CC[:, :, 1:, 0], 90), med(CC[:, :, 1:, 2] - CC[:, :, 1:, 1])))
# cross-SM chain in ns (global timer): own ship of refresh e+1 (compute, k=3 at period e) -> own detection of refresh e+1 (comm, k=1 at e1=e+1)
ship_ns = np.nanmax(G[:, :4, :-1, 3], axis=1), np.nanmax(G[:, 4:, :-1, 3], axis=1)      # last warp of slot A / B to ship
for sidx, nm in ((0, "A"), (1, "B")):
    det = np.nanmax(GC[:, sidx::2, 1:, 1], axis=1)
    arr = np.nanmax(GC[:, sidx::2, 1:, 2], axis=1)
    taken = np.nanmin(G[:, 4 * sidx:4 * sidx + 4, 1:, 0], axis=1)
    print("  slot %s (global ns): own ship -> all neighbours' rims detected %6.0f (p10 %6.0f p90 %6.0f); detected -> arrived %5.0f; arrived -> first warp past the wait %5.0f" % (
        nm, med(det - ship_ns[sidx]), np.nanpercentile(det - ship_ns[sidx], 10), np.nanpercentile(det - ship_ns[sidx], 90), med(arr - det), med(taken - arr)))
# neighbour skew: spread of ship times (ns) over the tiles of one unit for the same refresh
per_unit = plan["cx"] * plan["cy"]
nunits = nctas // per_unit
sh = ship_ns[0][:nunits * per_unit].reshape(nunits, per_unit, -1)
print("  ship-time spread over the tiles of a unit (slot A, ns): median %.0f, p90 %.0f" % (med(np.nanmax(sh, axis=1) - np.nanmin(sh, axis=1)), np.nanpercentile(np.nanmax(sh, axis=1) - np.nanmin(sh, axis=1), 90)))
print("  CTA total median %.0f max %.0f cycles" % (med(np.nanmax(rel[:, :, 7], axis=1)), np.nanmax(rel[:, :, 7])))
# stragglers: per CTA (median over periods, max over the slot's warps) busy time = period - wait
busy = np.nanmedian(np.nanmax((C[:, :4, 1:, 3] - C[:, :4, 1:, 0]), axis=1), axis=1)        # slot A: halo taken -> shipped
wait = np.nanmedian(np.nanmin((C[:, :4, 1:, 0] - C[:, :4, :-1, 3]), axis=1), axis=1)
order = np.argsort(wait)
print("  slot A tiles with the shortest wait (the ones everybody waits for): cta, pos in unit, wait, busy")
for i in order[:12]:
    print("    cta %3d  tile (%d,%d)  wait %6.0f  busy %6.0f" % (i, (i % per_unit) % plan["cx"], (i % per_unit) // plan["cx"], wait[i], busy[i]))
print("  wait percentiles over tiles: p5 %.0f p25 %.0f p50 %.0f p75 %.0f p95 %.0f;  busy: p5 %.0f p50 %.0f p95 %.0f max %.0f" % (
    *np.nanpercentile(wait, [5, 25, 50, 75, 95]), *np.nanpercentile(busy, [5, 50, 95]), np.nanmax(busy)))
# per-warp even-step durations within a CTA: who is slow
ev = np.nanmedian((C[:, :, 1:, 1] - C[:, :, 1:, 0]), axis=(0, 2))
od = np.nanmedian((C[:, :, 1:, 2] - C[:, :, 1:, 1]), axis=(0, 2))
shp = np.nanmedian((C[:, :, 1:, 3] - C[:, :, 1:, 2]), axis=(0, 2))
print("  per compute warp medians: even", np.round(ev), "odd", np.round(od), "ship", np.round(shp))
