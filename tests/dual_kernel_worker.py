"""Worker of tests/test_gpu_parity.py::test_dual_slot_forward_kernel: runs in its own process with CSPN_FWD_KERNEL=dual
(the selection is read once per process).  The dual-slot forward kernel over what its planner produces: one unit (one slot
per CTA), odd unit counts (last CTAs with one slot), several rounds, units cut with margins out of images larger than the
GPU, mode OURS, fp16, several depth channels, no sparse, T = 1 / 2 (no exchange) - against the C oracle; then a captured CUDA
graph replayed twice (the halo inboxes must come back clean)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cspn_monodepth_b200 import _lib, cspn_new, cspn_ours  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tests.util import make_inputs  # noqa: E402

assert os.environ.get("CSPN_FWD_KERNEL") == "dual"
dev = "cuda:0"
lib = _lib.load()


def cu(a, dtype=torch.float32):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dtype)


CASES = [
    (0, 1, 1, 228, 304, 24, 0.0072), (0, 7, 1, 228, 304, 24, 0.0072), (0, 8, 1, 228, 304, 24, 0.0072), (0, 19, 1, 228, 304, 24, 0.0072),
    (1, 5, 1, 228, 304, 24, 0.02), (0, 3, 1, 352, 1216, 24, 0.05), (1, 1, 1, 480, 640, 24, 0.02), (0, 1, 1, 720, 1280, 24, 0.01),
    (0, 2, 3, 120, 200, 9, 0.05), (0, 2, 1, 100, 72, 1, 0.05), (0, 2, 1, 100, 72, 2, 0.05), (0, 3, 1, 100, 136, 7, None),
    (0, 40, 1, 96, 128, 5, 0.05), (1, 2, 1, 64, 64, 40, 0.05), (0, 1, 1, 3, 1000, 24, 0.05),
]
for mode, b, c, h, w, iters, density in CASES:
    assert _lib.forward_plan(b, c, h, w, iters, 3, mode)["kernel"] == _lib.KERNEL_DUAL, (b, c, h, w)
    g, d, s = make_inputs(b * h + w + iters, b, 8, c, h, w, density=density)
    tg, td, ts = cu(g), cu(d), cu(s)
    y = cspn_new.AffinityPropagate(iters, 3)(tg, td, ts) if mode == 0 else cspn_ours.AffinityPropagate(prop_time=iters)(td, tg, sparse_depth=ts)
    assert lib.cspn_last_path() == _lib.PATH_FUSED and lib.cspn_last_launch_count() == 1
    ref = c_oracle.forward(g, d, s, iters, 3, mode, threads=0)
    err = float(np.abs(y.cpu().numpy() - ref).max())
    assert err <= 1e-4 * max(1.0, float(np.abs(ref).max()) / 10.0), f"dual mode {mode} {b}x{c}x{h}x{w} T {iters}: {err:.3e}"

# fp16 storage at the KITTI shape
g, d, s = (a.astype(np.float16) for a in make_inputs(11, 2, 8, 1, 352, 1216, density=0.05))
y = cspn_new.AffinityPropagate(24, 3)(cu(g, torch.float16), cu(d, torch.float16), cu(s, torch.float16))
ref = c_oracle.forward(g.astype(np.float32), d.astype(np.float32), s.astype(np.float32), 24, 3, 0, threads=0)
err = np.abs(y.float().cpu().numpy() - ref)
assert (err <= np.abs(ref) * 2.0 ** -10 + 1e-4).all(), f"dual fp16: {err.max():.3e}"

# CUDA graph: capture one forward, replay it twice, compare with the eager result
g, d, s = make_inputs(5, 10, 8, 1, 228, 304, density=0.02)
tg, td, ts = cu(g), cu(d), cu(s)
mod = cspn_new.AffinityPropagate(24, 3)
eager = mod(tg, td, ts)
torch.cuda.synchronize()
graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
with torch.cuda.stream(side):
    mod(tg, td, ts)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        captured = mod(tg, td, ts)
for _ in range(2):
    captured.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager), "graph replay differs from the eager result"
print("ok dual")
