"""Steady-state cost of one propagation step without cluster exchange: one 64x80 image per SM, long recurrences.
python tools/step_cost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cspn_monodepth_b200 import cspn_new
dev = torch.device("cuda:0")
B, H, W = 148, 80, 64
g = torch.randn(B, 8, H, W, device=dev); d = torch.rand(B, 1, H, W, device=dev) * 10
s = (torch.rand(B, 1, H, W, device=dev) < 0.01).float()
def t(iters, reps=20):
    m = cspn_new.AffinityPropagate(iters, 3)
    for _ in range(3): m(g, d, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): m(g, d, s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
a, b = t(120), t(60)
print(f"T=120: {a:.1f} us, T=60: {b:.1f} us -> {(a - b) / 60 * 1e3:.0f} ns/step = {(a - b) / 60 * 1.965e3:.0f} cycles/step at 1.965 GHz (sweep + intra-CTA row exchange)")
