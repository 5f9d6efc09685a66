"""Where the time of the host-buffer entry point goes (H2D of each input, kernel, D2H): python tools/e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cspn_monodepth_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
cfg = bench.NYU
g, d, s = bench.synth(cfg, 1)
arena = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
def carve(t, off):
    n = t.numel() * t.element_size()
    v = arena[off:off + n].view(t.dtype).view(t.shape); v.copy_(t); return v
hg, hd, hs = carve(g, 0), carve(d, 32 << 20), carve(s, 40 << 20)
ho = carve(d, 48 << 20)
dg, dd, ds, do = (torch.empty_like(t, device=dev) for t in (g, d, s, d))
def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print("H2D guidance %.0f us (%.1f GB/s)" % ((t := timed(lambda: dg.copy_(hg, non_blocking=True))), g.numel() * 4 / t / 1e3))
print("H2D depth    %.0f us" % timed(lambda: dd.copy_(hd, non_blocking=True)))
print("H2D sparse   %.0f us" % timed(lambda: ds.copy_(hs, non_blocking=True)))
print("D2H out      %.0f us" % timed(lambda: ho.copy_(do, non_blocking=True)))
b, h, w = cfg["B"], cfg["H"], cfg["W"]
stream = torch.cuda.current_stream().cuda_stream
def call():
    _lib.check(lib.cspn_fwd_host_f32(hg.data_ptr(), 8 * h * w, hd.data_ptr(), hs.data_ptr(), 1, ho.data_ptr(), b, 1, h, w, 24, 3, 0, stream))
for _ in range(3): call()
t0 = time.perf_counter()
for _ in range(50): call()
print("cspn_fwd_host_f32 wall %.0f us per call" % ((time.perf_counter() - t0) / 50 * 1e6))
