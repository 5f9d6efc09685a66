"""Parity of the CUDA path against the oracle / the reference's golden vectors.  Every call goes through the
C ABI (ctypes) - either via the drop-in nn.Modules or with raw device pointers.

Tolerances: forward max-abs <= 1e-4 at depth scale 10 (north_star), scaled with the depth range otherwise;
fp16 I/O: 1 fp16 ulp of the output (2^-10 relative) + the fp32 tolerance; gradients: 1e-4 relative to the
largest entry.
"""
import ctypes
import threading

import numpy as np
import pytest
import torch

import cspn_monodepth_b200 as pkg
from cspn_monodepth_b200 import _lib, cspn_new, cspn_ours
from oracle import c_oracle
from tests.util import assert_close_nan, case_config, make_inputs, nyu_golden_inputs

pytestmark = pytest.mark.gpu

FWD_ATOL = 1e-4
GRAD_RTOL = 1e-4
DEV = "cuda:0"
PATHS = [_lib.PATH_GENERIC, _lib.PATH_AUTO]


@pytest.fixture(autouse=True)
def _need_cuda_and_reset_path():
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    lib = _lib.load()
    yield
    lib.cspn_set_path(_lib.PATH_AUTO)


def _cu(a, dtype=torch.float32):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


def _run(mode, g, d, s, iters, requires_grad=False, dtype=torch.float32):
    tg, td, ts = _cu(g, dtype), _cu(d, dtype), _cu(s, dtype)
    if requires_grad:
        tg.requires_grad_(True); td.requires_grad_(True)
    if mode == 0:
        y = cspn_new.AffinityPropagate(iters, 3)(tg, td, ts)
    else:
        y = cspn_ours.AffinityPropagate(prop_time=iters)(td, tg, sparse_depth=ts)
    return y, tg, td


_ORACLE_CACHE = {}


def _oracle(key, g, d, s, iters, ksize, mode):
    """C oracle on all host cores, memoised so the two CUDA paths share one CPU evaluation."""
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = c_oracle.forward(g, d, s, iters, ksize, mode, threads=0)
    return _ORACLE_CACHE[key]


def _atol(ref):
    """1e-4 at the reference's depth scale of 10; grows with the value range (negative sparse entries make
    the recurrence expansive: (1-m) = 2)."""
    return FWD_ATOL * max(1.0, float(np.nanmax(np.abs(ref))) / 10.0)


def _golden_names(golden):
    return sorted(n for n in golden if "guidance" in golden[n])


@pytest.mark.parametrize("path", PATHS)
def test_forward_matches_reference_golden(golden, path):
    _lib.load().cspn_set_path(path)
    for name in _golden_names(golden):
        case = golden[name]
        mode, _, iters = case_config(name, case)
        y, _, _ = _run(mode, case["guidance"], case["depth"], case.get("sparse"), iters)
        scale = max(1.0, float(np.nanmax(np.abs(case["depth"]))) / 10.0)
        assert_close_nan(y.cpu().numpy(), case["out"], FWD_ATOL * scale, name)


@pytest.mark.parametrize("path", PATHS)
def test_backward_matches_reference_autograd(golden, path):
    _lib.load().cspn_set_path(path)
    checked = 0
    for name in _golden_names(golden):
        case = golden[name]
        if "grad_out" not in case:
            continue
        mode, _, iters = case_config(name, case)
        y, tg, td = _run(mode, case["guidance"], case["depth"], case.get("sparse"), iters, requires_grad=True)
        y.backward(_cu(case["grad_out"]))
        for got, key in ((tg.grad, "grad_guidance"), (td.grad, "grad_depth")):
            ref = case[key]
            assert_close_nan(got.cpu().numpy(), ref, GRAD_RTOL * max(1.0, np.abs(ref).max()), f"{name}:{key}")
        if mode == 0 and case["guidance"].shape[1] > 8:
            assert torch.count_nonzero(tg.grad[:, 8:]) == 0
        checked += 1
    assert checked >= 12


@pytest.mark.parametrize("path", PATHS)
def test_nyu_known_answer_and_oracle(golden, path):
    _lib.load().cspn_set_path(path)
    g, d, s = nyu_golden_inputs()
    y, _, _ = _run(0, g, d, s, 24)
    assert_close_nan(y.cpu().numpy(), golden["A_nyu_seed304228_T24"]["out"], FWD_ATOL, "nyu golden")
    assert_close_nan(y.cpu().numpy(), _oracle("nyu", g, d, s, 24, 3, 0), FWD_ATOL, "nyu oracle")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(8, 228, 304), (2, 352, 1216), (3, 97, 131), (1, 64, 64), (2, 65, 257), (1, 3, 1000), (1, 500, 5)])
def test_forward_vs_oracle_shape_grid(path, shape):
    _lib.load().cspn_set_path(path)
    b, h, w = shape
    for mode, cg in ((0, 8), (0, 12), (1, 8)):
        g, d, s = make_inputs(h * w + mode, b, cg, 1, h, w, density=0.02)
        y, _, _ = _run(mode, g, d, s, 24)
        assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s, 24, 3, mode), FWD_ATOL, f"{shape} mode {mode} cg {cg}")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("iters", [1, 2, 3, 7, 24, 40])
def test_iteration_counts(path, iters):
    _lib.load().cspn_set_path(path)
    g, d, s = make_inputs(iters, 2, 8, 1, 50, 77, density=0.05)
    y, _, _ = _run(0, g, d, s, iters)
    assert_close_nan(y.cpu().numpy(), _oracle(("it", iters), g, d, s, iters, 3, 0), FWD_ATOL, f"T={iters}")
    y, _, _ = _run(0, g, d, None, iters)
    assert_close_nan(y.cpu().numpy(), _oracle(("it-ns", iters), g, d, None, iters, 3, 0), FWD_ATOL, f"T={iters} no sparse")


@pytest.mark.parametrize("path", PATHS)
def test_five_by_five_pac_variant(path):
    _lib.load().cspn_set_path(path)
    g, d, s = make_inputs(55, 2, 24, 1, 60, 80, density=0.02)           # cfg4 shape family: K=5, T=12
    y, _, _ = _run(1, g, d, s, 12)
    assert _lib.load().cspn_last_path() == (_lib.PATH_GENERIC if path == _lib.PATH_GENERIC else _lib.PATH_BLOCKED)
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s, 12, 5, 1), FWD_ATOL, "5x5")
    g, d, s = make_inputs(77, 1, 48, 1, 20, 30, density=0.02)
    y, _, _ = _run(1, g, d, s, 4)
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s, 4, 7, 1), FWD_ATOL, "7x7")


# The temporally blocked 5x5 kernel (4 steps per launch, trapezoid tiles): ragged shapes, widths that are not a
# multiple of 4 (scalar loads), every launch count (T = 1..13), no sparse, negative sparse, fp16, several channels.
@pytest.mark.parametrize("shape", [(2, 480, 640), (1, 100, 52), (3, 33, 47), (1, 7, 203), (2, 130, 9)])
def test_blocked_5x5_shape_grid(shape):
    b, h, w = shape
    g, d, s = make_inputs(h * w, b, 24, 1, h, w, density=0.02)
    y, _, _ = _run(1, g, d, s, 12)
    assert _lib.load().cspn_last_path() == _lib.PATH_BLOCKED and _lib.load().cspn_last_launch_count() == 3
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s, 12, 5, 1, threads=0), FWD_ATOL, f"5x5 {shape}")


@pytest.mark.parametrize("iters", [1, 3, 4, 5, 8, 9, 13])
def test_blocked_5x5_iteration_counts_and_variants(iters):
    g, d, s = make_inputs(500 + iters, 2, 24, 2, 70, 96, density=0.05, neg=iters == 5, sparse_channels=2 if iters % 2 else 1)
    y, _, _ = _run(1, g, d, s, iters)
    assert _lib.load().cspn_last_launch_count() == (iters + 3) // 4
    ref = c_oracle.forward(g, d, s, iters, 5, 1)
    assert_close_nan(y.cpu().numpy(), ref, _atol(ref), f"5x5 T={iters}")
    y, _, _ = _run(1, g, d, None, iters)
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, None, iters, 5, 1), FWD_ATOL, f"5x5 T={iters} no sparse")
    g16, d16, s16 = (a.astype(np.float16) for a in (g, d, s))
    y, _, _ = _run(1, g16, d16, s16, iters, dtype=torch.float16)
    ref = c_oracle.forward(g16.astype(np.float32), d16.astype(np.float32), s16.astype(np.float32), iters, 5, 1)
    err = np.abs(y.float().cpu().numpy() - ref)
    assert (err <= np.abs(ref) * 2.0 ** -10 + _atol(ref)).all(), f"5x5 fp16 T={iters}: max err {err.max():.3e}"


@pytest.mark.parametrize("path", PATHS)
def test_multichannel_depth_and_per_channel_sparse(path):
    _lib.load().cspn_set_path(path)
    g, d, s = make_inputs(21, 2, 8, 3, 30, 41, density=0.05, sparse_channels=3)
    y, _, _ = _run(0, g, d, s, 9)
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s, 9, 3, 0), FWD_ATOL, "C=3, sparse C=3")
    y, _, _ = _run(0, g, d, s[:, :1], 9)
    assert_close_nan(y.cpu().numpy(), c_oracle.forward(g, d, s[:, :1], 9, 3, 0), FWD_ATOL, "C=3, sparse C=1")


@pytest.mark.parametrize("path", PATHS)
def test_fp16_io(path):
    _lib.load().cspn_set_path(path)
    g, d, s = make_inputs(16, 2, 8, 1, 120, 200, density=0.05)
    g16, d16, s16 = (a.astype(np.float16) for a in (g, d, s))
    y, _, _ = _run(0, g16, d16, s16, 24, dtype=torch.float16)
    ref = c_oracle.forward(g16.astype(np.float32), d16.astype(np.float32), s16.astype(np.float32), 24, 3, 0)
    err = np.abs(y.float().cpu().numpy() - ref)
    assert (err <= np.abs(ref) * 2.0 ** -10 + FWD_ATOL).all(), f"fp16 I/O: max err {err.max():.3e}"


@pytest.mark.parametrize("path", PATHS)
def test_fp16_backward_runs_and_is_close(path):
    _lib.load().cspn_set_path(path)
    g, d, s = make_inputs(17, 1, 8, 1, 40, 56, density=0.05)
    g16, d16, s16 = (a.astype(np.float16) for a in (g, d, s))
    go = np.random.default_rng(3).standard_normal(d.shape).astype(np.float16)
    y, tg, td = _run(0, g16, d16, s16, 12, requires_grad=True, dtype=torch.float16)
    y.backward(_cu(go, torch.float16))
    gg, gd = c_oracle.backward(g16.astype(np.float32), d16.astype(np.float32), s16.astype(np.float32), go.astype(np.float32), 12, 3, 0)
    assert np.abs(td.grad.float().cpu().numpy() - gd).max() <= 2e-3 * max(1.0, np.abs(gd).max())
    assert np.abs(tg.grad.float().cpu().numpy() - gg).max() <= 2e-3 * max(1.0, np.abs(gg).max())


@pytest.mark.parametrize("path", PATHS)
def test_backward_vs_oracle_larger(path):
    _lib.load().cspn_set_path(path)
    for mode, cg, k, iters in ((0, 12, 3, 24), (1, 8, 3, 24), (1, 24, 5, 12)):
        g, d, s = make_inputs(90 + mode + k, 2, cg, 1, 45, 61, density=0.03)
        go = np.random.default_rng(5).standard_normal(d.shape).astype(np.float32)
        y, tg, td = _run(mode, g, d, s, iters, requires_grad=True)
        y.backward(_cu(go))
        gg, gd = c_oracle.backward(g, d, s, go, iters, k, mode)
        assert_close_nan(td.grad.cpu().numpy(), gd, GRAD_RTOL * max(1.0, np.abs(gd).max()), f"gd mode {mode} k {k}")
        assert_close_nan(tg.grad.cpu().numpy(), gg, GRAD_RTOL * max(1.0, np.abs(gg).max()), f"gg mode {mode} k {k}")


@pytest.mark.parametrize("path", PATHS)
def test_unet_head_vectors(unet_golden, path):
    """The tensors the reference's own UNets hand to the module (12-channel guidance of magnitude 1e-2, smooth depth):
    forward and gradients against the reference module's results, tolerance relative to the value range."""
    _lib.load().cspn_set_path(path)
    for name, case in sorted(unet_golden.items()):
        mode, _, iters = case_config(name, case)
        y, tg, td = _run(mode, case["guidance"], case["depth"], case["sparse"], iters, requires_grad=True)
        y.backward(_cu(case["grad_out"]))
        assert_close_nan(y.detach().cpu().numpy(), case["out"], 1e-5 * np.abs(case["out"]).max(), name + ":out")
        for got, key in ((td.grad, "grad_depth"), (tg.grad, "grad_guidance")):
            assert_close_nan(got.cpu().numpy(), case[key], GRAD_RTOL * np.abs(case[key]).max(), f"{name}:{key}")
        if mode == 0:
            assert torch.count_nonzero(tg.grad[:, 8:]) == 0


def _check_backward(mode, cg, shape, iters, seed, density=0.03, expect_fused=True):
    b, h, w = shape
    g, d, s = make_inputs(seed, b, cg, 1, h, w, density=density)
    go = np.random.default_rng(seed + 1).standard_normal(d.shape).astype(np.float32)
    y, tg, td = _run(mode, g, d, s, iters, requires_grad=True)
    y.backward(_cu(go))
    if expect_fused:
        assert _lib.load().cspn_last_path() == _lib.PATH_FUSED and _lib.load().cspn_last_launch_count() == 1
    gg, gd = c_oracle.backward(g, d, s, go, iters, 3, mode, threads=0)
    what = f"{shape} mode {mode} cg {cg} T {iters}"
    assert_close_nan(td.grad.cpu().numpy(), gd, GRAD_RTOL * max(1.0, np.abs(gd).max()), "grad_depth " + what)
    assert_close_nan(tg.grad.cpu().numpy(), gg, GRAD_RTOL * max(1.0, np.abs(gg).max()), "grad_guidance " + what)


# The fused backward (one launch: recompute with history + reverse sweep + Jacobians) over the tilings it can take:
# single CTA, one cluster (DSMEM halo exchange), several cluster tiles with decaying margins, ragged shapes, and
# a batch large enough that the clusters do not fit in one wave (global-memory halo exchange).
@pytest.mark.parametrize("shape", [(2, 228, 304), (4, 228, 304), (1, 352, 1216), (3, 97, 131), (1, 64, 64), (2, 65, 257), (1, 3, 1000), (1, 500, 5)])
def test_fused_backward_shape_grid(shape):
    for mode, cg in ((0, 12), (1, 8)):
        _check_backward(mode, cg, shape, 24, seed=shape[1] * shape[2] + mode)


@pytest.mark.parametrize("iters", [1, 2, 3, 7, 40])
def test_fused_backward_iteration_counts(iters):
    _check_backward(0, 8, (2, 150, 140), iters, seed=iters)
    _check_backward(1, 8, (1, 70, 200), iters, seed=iters + 100, density=None if iters == 3 else 0.05)


@pytest.mark.parametrize("exchange", ["global", "dsmem"])
def test_forced_exchange_modes(exchange):
    """Both halo transports of the fused kernels on the same problems, fp32 and fp16 (the override is per process, hence
    the worker): stream mode with several tiles per persistent CTA, hardware clusters with several cluster tiles per image."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSPN_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "exchange_modes_worker.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok " + exchange), r.stdout[-2000:] + r.stderr[-4000:]


def test_dual_slot_forward_kernel():
    """The opt-in dual-slot forward kernel (CSPN_FWD_KERNEL=dual, read once per process, hence the worker) over what its
    planner produces, against the C oracle, plus CUDA-graph replay."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSPN_FWD_KERNEL="dual")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "dual_kernel_worker.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok dual"), r.stdout[-2000:] + r.stderr[-4000:]


def _fp16_case(seed, b, cg, h, w, density):
    g, d, s = make_inputs(seed, b, cg, 1, h, w, density=density)
    g16, d16, s16 = (a.astype(np.float16) for a in (g, d, s))
    go16 = np.random.default_rng(seed + 1).standard_normal(d.shape).astype(np.float16)
    return g16, d16, s16, go16


# BASELINE configs[2] at its own precision and size: fp16 storage, fp32 arithmetic, KITTI-shape images (stream /
# dual-slot forward with 147 tiles per half image, backward in stream mode with 132 tiles per image), and the NYU batch.
# Oracle: the C restatement in fp32 on the fp16-rounded inputs.  Tolerance: forward 1 fp16 ulp of the
# output (2^-10 relative) + 1e-4; gradients 1 fp16 ulp of the element + 1e-4 of the largest entry.
@pytest.mark.parametrize("b,cg,h,w", [(2, 8, 352, 1216), (2, 12, 352, 1216), (8, 8, 228, 304)])
def test_fp16_full_size_forward_backward_vs_oracle(b, cg, h, w):
    g16, d16, s16, go16 = _fp16_case(b * h + cg, b, cg, h, w, 0.05 if h == 352 else 0.0072)
    f32 = [a.astype(np.float32) for a in (g16, d16, s16, go16)]
    y, tg, td = _run(0, g16, d16, s16, 24, requires_grad=True, dtype=torch.float16)
    assert _lib.load().cspn_last_path() == _lib.PATH_FUSED and _lib.load().cspn_last_launch_count() == 1
    ref = c_oracle.forward(f32[0], f32[1], f32[2], 24, 3, 0, threads=0)
    err = np.abs(y.detach().float().cpu().numpy() - ref)
    assert (err <= np.abs(ref) * 2.0 ** -10 + FWD_ATOL).all(), f"fp16 forward {b}x{h}x{w} cg {cg}: max err {err.max():.3e}"
    y.backward(_cu(go16, torch.float16))
    assert _lib.load().cspn_last_path() == _lib.PATH_FUSED and _lib.load().cspn_last_launch_count() == 1
    gg, gd = c_oracle.backward(f32[0], f32[1], f32[2], f32[3], 24, 3, 0, threads=0)
    for got, want, what in ((td.grad, gd, "grad_depth"), (tg.grad, gg, "grad_guidance")):
        e = np.abs(got.float().cpu().numpy() - want)
        tol = np.abs(want) * 2.0 ** -10 + GRAD_RTOL * max(1.0, np.abs(want).max())
        assert (e <= tol).all(), f"fp16 {what} {b}x{h}x{w} cg {cg}: max err {e.max():.3e} (max |ref| {np.abs(want).max():.3e})"
    if cg > 8:
        assert torch.count_nonzero(tg.grad[:, 8:]) == 0


# One image with more tiles than the GPU has SMs (the old persistent stream deadlocked on these because
# every tile of an image advances in lockstep): the forward cuts them into resident units, the backward runs hardware
# clusters with margins.  Both must finish and match the oracle.
@pytest.mark.parametrize("h,w", [(720, 1280), (1080, 1440)])
def test_single_image_larger_than_the_gpu(h, w):
    g, d, s = make_inputs(h + w, 1, 8, 1, h, w, density=0.01)
    go = np.random.default_rng(h).standard_normal(d.shape).astype(np.float32)
    y, tg, td = _run(0, g, d, s, 24, requires_grad=True)
    assert _lib.load().cspn_last_path() == _lib.PATH_FUSED
    ref = c_oracle.forward(g, d, s, 24, 3, 0, threads=0)
    assert torch.isfinite(y).all()
    assert_close_nan(y.detach().cpu().numpy(), ref, FWD_ATOL, f"forward 1x{h}x{w}")
    y.backward(_cu(go))
    assert _lib.load().cspn_last_path() == _lib.PATH_FUSED
    gg, gd = c_oracle.backward(g, d, s, go, 24, 3, 0, threads=0)
    assert torch.isfinite(td.grad).all() and torch.isfinite(tg.grad).all()
    assert_close_nan(td.grad.cpu().numpy(), gd, GRAD_RTOL * max(1.0, np.abs(gd).max()), f"grad_depth 1x{h}x{w}")
    assert_close_nan(tg.grad.cpu().numpy(), gg, GRAD_RTOL * max(1.0, np.abs(gg).max()), f"grad_guidance 1x{h}x{w}")


def test_torch_library_ops_match_the_ctypes_path():
    """The C++ operator layer (torch.ops.cspn.*, the default route of the nn.Modules) against the ctypes autograd.Function
    over the same C ABI: bit-identical results and gradients, None / 1-channel / C-channel sparse, fp16, strided guidance,
    errors as RuntimeError."""
    from cspn_monodepth_b200 import functional
    ops = _lib.torch_ops()
    assert ops is not None
    for mode, cg, c, dtype, sc in ((0, 12, 1, torch.float32, 1), (1, 8, 1, torch.float32, None), (0, 8, 3, torch.float16, 3), (1, 24, 1, torch.float32, 1)):
        k = 5 if cg == 24 else 3
        g, d, s = make_inputs(cg + c, 2, cg, c, 70, 96, density=None if sc is None else 0.05, sparse_channels=sc or 1)
        res = []
        for use_ops in (True, False):
            tg, td, ts = _cu(g, dtype).requires_grad_(True), _cu(d, dtype).requires_grad_(True), _cu(s, dtype)
            y = ops.propagate(tg, td, ts, 12, k, mode) if use_ops else functional._CspnPropagate.apply(tg, td, ts, 12, k, mode)
            y.backward(torch.ones_like(y))
            res.append((y.detach(), tg.grad, td.grad))
        for a, b in zip(*res):
            assert torch.equal(a, b)
    wide = _cu(make_inputs(3, 2, 12, 1, 33, 48)[0])
    d, s = _cu(make_inputs(3, 2, 12, 1, 33, 48)[1]), _cu(make_inputs(3, 2, 12, 1, 33, 48)[2])
    assert torch.equal(ops.forward(wide.narrow(1, 0, 8), d, s, 24, 3, 0), ops.forward(wide[:, :8].contiguous(), d, s, 24, 3, 0))
    with pytest.raises(RuntimeError):
        ops.forward(wide[:, :7], d, s, 24, 3, 0)
    with pytest.raises(RuntimeError):
        ops.forward(wide, d, s, 24, 5, 0)                               # CSPN_new only works with prop_kernel 3
    with torch.no_grad():
        assert ops.propagate(wide, d, s, 0, 3, 0) is d or torch.equal(ops.propagate(wide, d, s, 0, 3, 0), d)


def test_two_gpu_threaded_replicas():
    """The reference's DataParallel pattern (network/libs/base/encoding.py:102-105): one Python thread per GPU, each calling
    the module on its own batch slice at the same time.  Needs two devices (gpurun --gpus 2); results must equal the
    single-device run bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    g, d, s = make_inputs(91, 8, 12, 1, 228, 304, density=0.0072)
    go = np.random.default_rng(4).standard_normal(d.shape).astype(np.float32)
    y_full, tg_full, td_full = _run(0, g, d, s, 24, requires_grad=True)
    y_full.backward(_cu(go))
    results = [None, None]

    def replica(i):
        dev = torch.device("cuda", i)
        sl = slice(4 * i, 4 * i + 4)
        with torch.cuda.device(dev):
            tg = torch.from_numpy(g[sl]).to(dev).requires_grad_(True)
            td = torch.from_numpy(d[sl]).to(dev).requires_grad_(True)
            ts = torch.from_numpy(s[sl]).to(dev)
            mod = cspn_new.AffinityPropagate(24, 3)
            for _ in range(5):                                          # several calls: per-device caches are hit concurrently
                y = mod(tg, td, ts)
            y.backward(torch.from_numpy(go[sl]).to(dev))
            torch.cuda.synchronize(dev)
            results[i] = (y.detach().cpu(), tg.grad.cpu(), td.grad.cpu())

    threads = [threading.Thread(target=replica, args=(i,)) for i in range(2)]
    [t.start() for t in threads]; [t.join() for t in threads]
    for i in range(2):
        sl = slice(4 * i, 4 * i + 4)
        assert results[i] is not None
        assert torch.equal(results[i][0], y_full.detach().cpu()[sl])
        assert torch.equal(results[i][1], tg_full.grad.cpu()[sl]) and torch.equal(results[i][2], td_full.grad.cpu()[sl])


def test_fused_backward_with_several_depth_channels():
    """C > 1 depth channels share the affinity of their image (pac.py:77-78,118-119): the fused backward runs once per
    channel and accumulates grad_guidance - C + (C - 1) launches instead of the generic path's ~50 - for both modes, one /
    C sparse channels, fp32 and fp16."""
    for mode, cg, c, sc, dtype in ((0, 12, 3, 1, torch.float32), (1, 8, 2, 2, torch.float32), (0, 8, 3, 3, torch.float16)):
        g, d, s = make_inputs(33 + c + mode, 2, cg, c, 120, 136, density=0.04, sparse_channels=sc)
        go = np.random.default_rng(c).standard_normal(d.shape).astype(np.float32)
        if dtype == torch.float16:
            g, d, s, go = (a.astype(np.float16).astype(np.float32) for a in (g, d, s, go))
        y, tg, td = _run(mode, g, d, s, 24, requires_grad=True, dtype=dtype)
        y.backward(_cu(go, dtype))
        # launch count through the raw C ABI on this thread (autograd runs the module's backward on its own thread, and the
        # library's call statistics are per host thread)
        lib = _lib.load()
        b, _, h, w = d.shape
        n = lib.cspn_bwd_workspace_bytes(b, c, h, w, 24, 3, mode)
        ws = torch.empty(n, dtype=torch.uint8, device=DEV)
        raw = [_cu(a, dtype) for a in (go, g, d, s)]
        gg_raw, gd_raw = torch.empty_like(raw[1]), torch.empty_like(raw[2])
        fn = lib.cspn_bwd_f32 if dtype == torch.float32 else lib.cspn_bwd_f16
        rc = fn(raw[0].data_ptr(), raw[1].data_ptr(), cg * h * w, cg, raw[2].data_ptr(), raw[3].data_ptr(), sc, gg_raw.data_ptr(), gd_raw.data_ptr(),
                b, c, h, w, 24, 3, mode, ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        assert rc == 0 and lib.cspn_last_path() == _lib.PATH_FUSED and lib.cspn_last_launch_count() == 2 * c - 1
        torch.cuda.synchronize()
        assert torch.equal(gg_raw, tg.grad) and torch.equal(gd_raw, td.grad)
        gg, gd = c_oracle.backward(g, d, s, go, 24, 3, mode, threads=0)
        ulp = 2.0 ** -10 if dtype == torch.float16 else 0.0
        for got, want, what in ((td.grad, gd, "grad_depth"), (tg.grad, gg, "grad_guidance")):
            e = np.abs(got.float().cpu().numpy() - want)
            tol = np.abs(want) * ulp * c + GRAD_RTOL * max(1.0, np.abs(want).max())
            assert (e <= tol).all(), f"C={c} mode {mode} {dtype} {what}: max err {e.max():.3e}"
        if cg > 8:
            assert torch.count_nonzero(tg.grad[:, 8:]) == 0


@pytest.mark.parametrize("b,cg,c,sc,h,w,iters,dtype", [(2, 24, 1, 1, 60, 80, 12, torch.float32), (1, 24, 1, 1, 100, 52, 12, torch.float32),
                                                        (2, 26, 2, 2, 45, 61, 5, torch.float32), (1, 24, 3, 1, 33, 47, 1, torch.float32),
                                                        (1, 24, 1, 1, 130, 9, 9, torch.float32), (2, 24, 1, 1, 96, 128, 12, torch.float16),
                                                        (1, 24, 1, None, 70, 203, 4, torch.float32), (2, 24, 1, 1, 480, 640, 12, torch.float32)])
def test_blocked_5x5_backward(b, cg, c, sc, h, w, iters, dtype):
    """Backward of the 5x5 variant: forward recompute and adjoint recurrence temporally blocked (4 steps per launch, weights /
    transposed weights in registers), Jacobians in one pass - ceil((T-1)/4) + ceil(T/4) + 1 launches, vs the C oracle's closed
    form (pac.py:96-121 through the loop of CSPN_ours.py:47-53): ragged shapes, widths that are not a multiple of 4, several
    depth channels with their own sparse channel, extra guidance channels (zero gradient), no sparse, fp16, one full-size image pair."""
    g, d, s = make_inputs(500 + h + iters, b, cg, c, h, w, density=None if sc is None else 0.03, sparse_channels=sc or 1)
    go = np.random.default_rng(c + w).standard_normal(d.shape).astype(np.float32)
    if dtype == torch.float16:
        g, d, s, go = (a.astype(np.float16).astype(np.float32) for a in (g, d, s, go))
    lib = _lib.load()
    n = lib.cspn_bwd_workspace_bytes(b, c, h, w, iters, 5, 1)
    ws = torch.empty(n, dtype=torch.uint8, device=DEV)
    raw = [_cu(a, dtype) for a in (go, g, d, s)]
    gg_raw, gd_raw = torch.full_like(raw[1], float("nan")), torch.full_like(raw[2], float("nan"))
    fn = lib.cspn_bwd_f32 if dtype == torch.float32 else lib.cspn_bwd_f16
    rc = fn(raw[0].data_ptr(), raw[1].data_ptr(), cg * h * w, cg, raw[2].data_ptr(), None if s is None else raw[3].data_ptr(), sc or 1,
            gg_raw.data_ptr(), gd_raw.data_ptr(), b, c, h, w, iters, 5, 1, ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and lib.cspn_last_path() == _lib.PATH_BLOCKED
    assert lib.cspn_last_launch_count() == (iters - 1 + 3) // 4 + (iters + 3) // 4 + 1
    torch.cuda.synchronize()
    gg, gd = c_oracle.backward(g[:, :24], d, s, go, iters, 5, 1, threads=0)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 0.0
    for got, want, what in ((gd_raw, gd, "grad_depth"), (gg_raw[:, :24], gg, "grad_guidance")):
        e = np.abs(got.float().cpu().numpy() - want)
        tol = np.abs(want) * ulp * c + GRAD_RTOL * max(1.0, np.abs(want).max())
        assert (e <= tol).all(), f"{what}: max err {e.max():.3e} (scale {np.abs(want).max():.3e})"
    if cg > 24:
        assert torch.count_nonzero(gg_raw[:, 24:]) == 0
    # the generic multi-launch path computes the same thing
    lib.cspn_set_path(_lib.PATH_GENERIC)
    n2 = lib.cspn_bwd_workspace_bytes(b, c, h, w, iters, 5, 1)
    ws2 = torch.empty(n2, dtype=torch.uint8, device=DEV)
    gg2, gd2 = torch.empty_like(raw[1]), torch.empty_like(raw[2])
    rc = fn(raw[0].data_ptr(), raw[1].data_ptr(), cg * h * w, cg, raw[2].data_ptr(), None if s is None else raw[3].data_ptr(), sc or 1,
            gg2.data_ptr(), gd2.data_ptr(), b, c, h, w, iters, 5, 1, ws2.data_ptr(), n2, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and lib.cspn_last_path() == _lib.PATH_GENERIC
    scale = max(1.0, float(gg2.float().abs().max()))
    assert float((gg2.float() - gg_raw.float()).abs().max()) <= (2 * GRAD_RTOL + 4 * ulp) * scale


def test_pipelined_host_entry_points():
    """cspn_fwd_host_submit_* / cspn_host_wait: several calls in flight (H2D of one overlaps kernel + D2H of the previous),
    every result checked against the oracle; fp32 with a strided 12-channel guidance and fp16; tickets are per call."""
    lib = _lib.load()
    depth = lib.cspn_host_pipeline_depth()
    assert depth >= 2
    cases = []
    for i in range(2 * depth + 1):
        g, d, s = make_inputs(200 + i, 8 if i % 2 == 0 else 3, 12, 1, 228, 304, density=0.0072)
        cases.append((g, d, s))
    pinned = [[torch.from_numpy(a).pin_memory() for a in c] for c in cases]
    outs = [torch.empty_like(p[1]).pin_memory() for p in pinned]
    tickets = []
    for (hg, hd, hs), out in zip(pinned, outs):
        t = ctypes.c_int(0)
        b = hd.shape[0]
        rc = lib.cspn_fwd_host_submit_f32(hg.data_ptr(), 12 * 228 * 304, hd.data_ptr(), hs.data_ptr(), 1, out.data_ptr(), b, 1, 228, 304, 24, 3, 0, ctypes.byref(t))
        assert rc == 0 and t.value > 0
        tickets.append(t.value)
    assert len(set(tickets)) == len(tickets)
    for t in reversed(tickets):                      # waiting out of order is fine (older calls were retired by later submits)
        assert lib.cspn_host_wait(t) == 0
    assert lib.cspn_host_wait(tickets[0]) == 0      # waiting twice is a no-op
    for (g, d, s), out in zip(cases, outs):
        assert_close_nan(out.numpy(), c_oracle.forward(g, d, s, 24, 3, 0, threads=0), FWD_ATOL, "pipelined host entry")
    g, d, s = (a.astype(np.float16) for a in make_inputs(300, 2, 8, 1, 352, 1216, density=0.05))
    hg, hd, hs = (torch.from_numpy(a).pin_memory() for a in (g, d, s))
    out = torch.empty_like(hd).pin_memory()
    t = ctypes.c_int(0)
    assert lib.cspn_fwd_host_submit_f16(hg.data_ptr(), 8 * 352 * 1216, hd.data_ptr(), hs.data_ptr(), 1, out.data_ptr(), 2, 1, 352, 1216, 24, 3, 0, ctypes.byref(t)) == 0
    assert lib.cspn_host_wait(t.value) == 0
    ref = c_oracle.forward(g.astype(np.float32), d.astype(np.float32), s.astype(np.float32), 24, 3, 0, threads=0)
    err = np.abs(out.float().numpy() - ref)
    assert (err <= np.abs(ref) * 2.0 ** -10 + FWD_ATOL).all()
    t0 = ctypes.c_int(5)
    assert lib.cspn_fwd_host_submit_f32(hg.data_ptr(), 8, hd.data_ptr(), None, 1, out.data_ptr(), 0, 1, 4, 4, 2, 3, 0, ctypes.byref(t0)) == 0 and t0.value == 0   # empty batch
    assert lib.cspn_fwd_host_submit_f32(None, 8 * 16, hd.data_ptr(), None, 1, out.data_ptr(), 1, 1, 4, 4, 2, 3, 0, ctypes.byref(t0)) == -1


def test_hybrid_row_cluster_transport():
    """Row-cluster transport of the forward kernel (every tile row of an image = one hardware cluster: DSMEM left / right, global
    inboxes up / down, row rims shipped from inside the sweep), forced with CSPN_EXCHANGE=hybrid: plan, parity vs the C oracle in
    both modes and precisions, fall-back plans when the clusters do not fit at once, CUDA-graph replay."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSPN_EXCHANGE="hybrid")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "hybrid_worker.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok hybrid"), r.stdout[-2000:] + r.stderr[-4000:]
