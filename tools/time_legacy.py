"""Timing of the legacy max-of-8 CSPN kernel (csrc/cspn_legacy.cu), 16 steps on 8 x 304 x 228: python tools/time_legacy.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cspn_monodepth_b200 import cspn_legacy  # noqa: E402

dev = torch.device("cuda:0")
g, d, s = [t.to(dev) for t in bench.synth(bench.NYU, 0)]
with torch.no_grad():
    ms = bench._event_ms(lambda: cspn_legacy.legacy_propagate(g, d, s, 16), 100, dev)
print("legacy CSPN-16, 8 x 304 x 228 fp32: %.1f us (4 launches)" % (ms * 1e3))
