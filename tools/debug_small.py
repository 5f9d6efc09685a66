"""Tiny forward runs for debugging: python tools/debug_small.py H W [mode] [iters] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cspn_monodepth_b200 import cspn_new, cspn_ours, _lib
from oracle import c_oracle
from tests.util import make_inputs
h, w = int(sys.argv[1]), int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 24
dt = torch.float16 if len(sys.argv) > 5 and sys.argv[5] == "f16" else torch.float32
g, d, s = make_inputs(1, 2, 8, 1, h, w, density=0.05)
if dt == torch.float16:
    g, d, s = (a.astype(np.float16).astype(np.float32) for a in (g, d, s))
tg, td, ts = (torch.from_numpy(a).cuda().to(dt) for a in (g, d, s))
y = cspn_new.AffinityPropagate(iters, 3)(tg, td, ts) if mode == 0 else cspn_ours.AffinityPropagate(iters)(td, tg, ts)
torch.cuda.synchronize()
ref = c_oracle.forward(g, d, s, iters, 3, mode)
err = np.abs(y.float().cpu().numpy() - ref)
print(f"{h}x{w} mode {mode} T {iters} {dt}: path {_lib.load().cspn_last_path()} max-abs {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}")
