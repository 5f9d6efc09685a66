"""Worker of tests/test_gpu_parity.py::test_forced_exchange_modes: runs in its own process because the halo transport
override (CSPN_EXCHANGE=dsmem|global) is read once per process.  Checks forward and backward of both modes against the
C oracle on shapes that give several tiles per persistent CTA in stream mode (more tiles than SMs), several cluster
tiles per image with hardware clusters, and replays a captured CUDA graph twice (halo inboxes must come back clean)."""
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


def check16(mode, shape, iters, seed):
    """fp16 storage (BASELINE configs[2]'s precision) against the oracle on the fp16-rounded inputs: 1 fp16 ulp of the
    element + the fp32 tolerance."""
    b, h, w = shape
    g, d, s = (a.astype(np.float16) for a in make_inputs(seed, b, 8, 1, h, w, density=0.03))
    go = np.random.default_rng(seed + 7).standard_normal(d.shape).astype(np.float16)
    tg, td, ts = (cu(a) for a in (g, d, s))
    tg.requires_grad_(True); td.requires_grad_(True)
    mod = cspn_new.AffinityPropagate(iters, 3) if mode == 0 else cspn_ours.AffinityPropagate(iters)
    y = mod(tg, td, ts) if mode == 0 else mod(td, tg, sparse_depth=ts)
    assert lib.cspn_last_path() == _lib.PATH_FUSED
    f32 = [a.astype(np.float32) for a in (g, d, s, go)]
    ref = c_oracle.forward(f32[0], f32[1], f32[2], iters, 3, mode, threads=0)
    err = np.abs(y.detach().float().cpu().numpy() - ref)
    assert (err <= np.abs(ref) * 2.0 ** -10 + 1e-4).all(), f"fp16 forward {shape} mode {mode}: {err.max():.3e}"
    y.backward(cu(go))
    assert lib.cspn_last_path() == _lib.PATH_FUSED and lib.cspn_last_launch_count() == 1
    gg, gd = c_oracle.backward(f32[0], f32[1], f32[2], f32[3], iters, 3, mode, threads=0)
    for got, want, what in ((td.grad, gd, "grad_depth"), (tg.grad, gg, "grad_guidance")):
        e = np.abs(got.float().cpu().numpy() - want)
        assert (e <= np.abs(want) * 2.0 ** -10 + 1e-4 * max(1.0, np.abs(want).max())).all(), f"fp16 {what} {shape} mode {mode}: {e.max():.3e}"


def check(mode, shape, iters, seed, backward):
    b, h, w = shape
    g, d, s = make_inputs(seed, b, 8, 1, h, w, density=0.03)
    tg, td, ts = cu(g), cu(d), cu(s)
    if backward:
        tg.requires_grad_(True); td.requires_grad_(True)
    mod = cspn_new.AffinityPropagate(iters, 3) if mode == 0 else cspn_ours.AffinityPropagate(iters)
    y = mod(tg, td, ts) if mode == 0 else mod(td, tg, sparse_depth=ts)
    assert lib.cspn_last_path() == _lib.PATH_FUSED
    ref = c_oracle.forward(g, d, s, iters, 3, mode, threads=0)
    err = float(np.abs(y.detach().cpu().numpy() - ref).max())
    assert err <= 1e-4, f"forward {shape} mode {mode}: {err:.3e}"
    if backward:
        go = np.random.default_rng(seed + 7).standard_normal(d.shape).astype(np.float32)
        y.backward(cu(go))
        assert lib.cspn_last_path() == _lib.PATH_FUSED and lib.cspn_last_launch_count() == 1
        gg, gd = c_oracle.backward(g, d, s, go, iters, 3, mode, threads=0)
        e1 = float(np.abs(td.grad.cpu().numpy() - gd).max() / max(1.0, np.abs(gd).max()))
        e2 = float(np.abs(tg.grad.cpu().numpy() - gg).max() / max(1.0, np.abs(gg).max()))
        assert e1 <= 1e-4 and e2 <= 1e-4, f"backward {shape} mode {mode}: {e1:.3e} {e2:.3e}"
    return y.detach()


# 12 x 15 = 180 forward tiles / 12 x 20 = 240 backward tiles on 148 SMs; 30 KITTI-ish tiles per image; ragged small ones
for mode, shape, iters in ((0, (12, 228, 304), 24), (1, (3, 352, 500), 24), (0, (40, 97, 131), 7), (0, (2, 352, 1216), 24), (1, (5, 65, 257), 3)):
    check(mode, shape, iters, seed=shape[1] + iters, backward=True)

# BASELINE configs[1] / [2] shapes in fp16 through the forced transport (forward and backward)
for mode, shape in ((0, (8, 228, 304)), (0, (2, 352, 1216)), (1, (3, 228, 304))):
    check16(mode, shape, 24, seed=shape[0] + shape[2])

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
print("ok", os.environ.get("CSPN_EXCHANGE", "auto"))
