"""Run a few forwards of one configuration (for ncu captures): python tools/run_one.py nyu|kitti|pac5 [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "nyu"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = {"nyu": bench.NYU, "kitti": bench.KITTI, "pac5": bench.PAC5}[name]
dev = torch.device("cuda:0")
sets = [[t.to(dev) for t in bench.synth(cfg, i)] for i in range(2)]
step = bench.run_module(cfg, sets)
with torch.no_grad():
    for i in range(reps):
        y = step(i)
torch.cuda.synchronize()
print(name, float(y.float().mean()))
