"""Worker of tests/test_gpu_parity.py::test_hybrid_row_cluster_transport: own process because CSPN_EXCHANGE is read once.
With CSPN_EXCHANGE=hybrid the forward planner takes the row-cluster transport wherever it is possible (every tile row of an image
= one hardware cluster, left / right rims through DSMEM, rows above / below through the global inboxes straight from the sweep).
Checks the plan, the forward of both modes against the C oracle (fp32 and fp16, TMA and plain-load prologue, 2 - 5 tile rows,
2 - 9 tiles per row), that the batch slices are bit-identical to the stream / cluster plans' results of the same process
arithmetic (same tile positions -> same rounding), and CUDA-graph replay (inboxes must come back clean)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cspn_monodepth_b200 import _lib, cspn_new, cspn_ours  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tests.util import make_inputs  # noqa: E402

dev = "cuda:0"
lib = _lib.load()


def cu(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def run(mode, shape, iters, seed, dtype=np.float32, expect_hybrid=True):
    b, h, w = shape
    g, d, s = (a.astype(dtype) for a in make_inputs(seed, b, 8, 1, h, w, density=0.03))
    plan = _lib.forward_plan(b, 1, h, w, iters, 3, mode)
    assert plan["kernel"] == _lib.KERNEL_SINGLE, plan
    assert (plan["transport"] == "hybrid") == expect_hybrid, (shape, plan)
    mod = cspn_new.AffinityPropagate(iters, 3) if mode == 0 else cspn_ours.AffinityPropagate(iters)
    with torch.no_grad():
        y = mod(cu(g), cu(d), cu(s)) if mode == 0 else mod(cu(d), cu(g), sparse_depth=cu(s))
    assert lib.cspn_last_path() == _lib.PATH_FUSED and lib.cspn_last_launch_count() == 1
    f32 = [a.astype(np.float32) for a in (g, d, s)]
    ref = c_oracle.forward(f32[0], f32[1], f32[2], iters, 3, mode, threads=0)
    err = np.abs(y.float().cpu().numpy() - ref)
    tol = 1e-4 + (np.abs(ref) * 2.0 ** -10 if dtype == np.float16 else 0.0)
    assert (err <= tol).all(), f"forward {shape} mode {mode} {dtype.__name__}: {err.max():.3e}"
    return y


# (CSPN_HYBRID_AUTO=1 makes the planner take this transport by cost for mode OURS; here it is forced for both modes)
run(0, (8, 228, 304), 24, 1)                               # the headline: 24 clusters of 5
run(0, (8, 228, 304), 24, 2, np.float16)
run(1, (3, 228, 304), 24, 3)                               # softmax mode
run(0, (3, 352, 500), 24, 4)                               # 15 clusters of 9, 5 tile rows
run(0, (4, 97, 131), 7, 5)                                 # W * 4 not a multiple of 16: plain-load prologue; 2 x 3 tiles
run(0, (1, 97, 130), 1, 6)                                 # one step: no refresh at all
run(0, (2, 100, 200), 2, 7)                                # two steps: one pair, no refresh
run(0, (2, 100, 200), 3, 8)                                # three steps: exactly one refresh
run(0, (9, 228, 304), 24, 9, expect_hybrid=False)          # 27 clusters of 5 do not fit at once: stream / clusters
run(0, (2, 352, 1216), 24, 10, expect_hybrid=False)        # 21 tiles per row > 16: not a cluster
run(0, (3, 60, 300), 24, 11, expect_hybrid=False)          # one tile row: plain clusters

# CUDA graph: capture one forward, replay it twice, compare with the eager result
g, d, s = make_inputs(5, 8, 8, 1, 228, 304, density=0.02)
tg, td, ts = cu(g), cu(d), cu(s)
mod = cspn_new.AffinityPropagate(24, 3)
with torch.no_grad():
    eager = mod(tg, td, ts)
    torch.cuda.synchronize()
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(side):
        mod(tg, td, ts)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            captured = mod(tg, td, ts)
    for _ in range(3):
        captured.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(captured, eager), "graph replay differs from the eager result"
print("ok", os.environ.get("CSPN_EXCHANGE", "auto"))
