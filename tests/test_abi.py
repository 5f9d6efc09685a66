"""The C-ABI library loads and exports exactly what include/cspn_b200.h declares.  No GPU needed:
argument validation happens before any CUDA call."""
import ctypes
import os
import re

import pytest

from cspn_monodepth_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cspn_b200.h")).read()
    return sorted(set(re.findall(r"CSPN_API\s+[\w\s\*]+?\b(cspn_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("cspn_fwd_f32", "cspn_fwd_f16", "cspn_bwd_f32", "cspn_bwd_f16", "cspn_fwd_host_f32",
                 "cspn_fwd_workspace_bytes", "cspn_bwd_workspace_bytes", "cspn_error_string", "cspn_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.library_path()) if os.path.exists(_lib.library_path()) else _lib.load()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in cspn_b200.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_abi_version_and_error_strings():
    lib = _lib.load()
    assert lib.cspn_abi_version() == 1
    assert lib.cspn_error_string(0) == b"success"
    for code in range(-8, 0):
        assert lib.cspn_error_string(code).startswith(b"cspn:")


def test_argument_validation_without_gpu():
    lib = _lib.load()
    one = ctypes.c_float(0.0)
    p = ctypes.addressof(one)

    def fwd(guidance=p, gbs=8 * 16, depth=p, sparse=None, sc=1, out=p + 0, b=1, c=1, h=4, w=4, iters=2, k=3, mode=0, ws=None, wsb=0):
        return lib.cspn_fwd_f32(guidance, gbs, depth, sparse, sc, out, b, c, h, w, iters, k, mode, ws, wsb, None)

    assert fwd(mode=7) == -4
    assert fwd(k=5) == -3                   # CSPN_new only works with prop_kernel 3 (CSPN_new.py:122)
    assert fwd(k=4, mode=1) == -3
    assert fwd(h=0) == -2
    assert fwd(iters=-1) == -2
    assert fwd(guidance=None) == -1
    assert fwd(gbs=7 * 16) == -5            # fewer than 8 guidance channels
    assert fwd(sparse=p, sc=2) == -7
    assert fwd(b=0) == 0                    # empty batch is a no-op
    assert fwd(out=p) == -8                 # out aliases depth
    with pytest.raises(_lib.CspnError):
        _lib.check(-3)


def test_workspace_queries():
    lib = _lib.load()
    assert lib.cspn_fwd_workspace_bytes(1, 1, 8, 8, 0, 3, 0) == 0
    prev = lib.cspn_set_path(_lib.PATH_GENERIC)
    try:
        n = lib.cspn_fwd_workspace_bytes(2, 1, 16, 16, 24, 3, 0)
        assert n >= (2 * 8 + 2 * 2) * 256 * 4
        assert lib.cspn_fwd_workspace_bytes(2, 1, 16, 16, 12, 5, 1) >= (2 * 24 + 4) * 256 * 4
    finally:
        lib.cspn_set_path(prev)
    assert lib.cspn_bwd_workspace_bytes(1, 1, 16, 16, 24, 3, 0) >= (8 + 8 + 1 + 23 + 3) * 256 * 4
    assert lib.cspn_bwd_workspace_bytes(1, 1, 16, 16, 24, 4, 0) == 0


def test_planner_picks_kernel_and_workspace_by_problem_size():
    """The workspace query runs the same planner as the launch (B200 cluster capacities as defaults without a GPU):
    hardware clusters need no scratch, stream mode one 16-byte-slot inbox per tile (576 slots for a 64x80 tile) and only
    when a whole image is resident at once, the blocked 5x5 path two fp32 planes, the fused backward its history scratch
    (T x 64 x 64 floats per slot: one slot per CTA for launches of up to 192 CTAs - every stream-mode launch -, else 192 SM-id slots)."""
    lib = _lib.load()
    inbox = (4 * 80 + 4 * 2 * 32) * 16
    assert _lib.forward_plan(8, 1, 228, 304, 24)["kernel"] == _lib.KERNEL_SINGLE
    # halo transport of the single-tile kernel: 7 images fit as 5x3 hardware clusters; 8 do not and stream (the row-cluster / hybrid
    # transport is opt-in: tests/hybrid_worker.py)
    assert _lib.forward_plan(7, 1, 228, 304, 24)["transport"] == "cluster"
    for mode in (0, 1):
        p = _lib.forward_plan(8, 1, 228, 304, 24, 3, mode)
        assert (p["transport"], p["cx"], p["cy"]) == ("stream", 5, 3), p
    assert lib.cspn_fwd_workspace_bytes(8, 1, 228, 304, 24, 3, 1) == lib.cspn_fwd_workspace_bytes(8, 1, 228, 304, 24, 3, 0)
    assert lib.cspn_fwd_workspace_bytes(1, 1, 228, 304, 24, 3, 0) == 0                 # one 5x3 cluster: DSMEM
    assert lib.cspn_fwd_workspace_bytes(7, 1, 228, 304, 24, 3, 0) == 0                 # 7 clusters of 15 fit at once
    assert lib.cspn_fwd_workspace_bytes(8, 1, 228, 304, 24, 3, 0) == 256 + 8 * 15 * inbox    # the 8th would not: row clusters + global inboxes (or stream mode), status word + 120 inboxes
    assert lib.cspn_fwd_workspace_bytes(32, 1, 352, 1216, 24, 3, 0) == 0               # KITTI batch: 4x2 hardware clusters
    assert lib.cspn_fwd_workspace_bytes(8, 1, 64, 64, 24, 3, 0) == 0                   # single-tile images
    # one image with more tiles than SMs (720p: 22 x 10 = 220): never the lockstep stream (it would wait on tiles that have
    # not started), hardware clusters with margins instead
    assert lib.cspn_fwd_workspace_bytes(1, 1, 720, 1280, 24, 3, 0) == 0
    assert lib.cspn_fwd_workspace_bytes(2, 1, 1080, 1440, 24, 3, 0) == 0
    assert lib.cspn_bwd_workspace_bytes(1, 1, 1080, 1920, 24, 3, 0) == 192 * 24 * 64 * 64 * 4          # backward: history only, no inboxes
    assert _lib.forward_plan(16, 1, 480, 640, 12, 5, 1)["kernel"] == _lib.KERNEL_BLOCKED
    assert _lib.forward_plan(1, 1, 20, 30, 4, 7, 1)["kernel"] == _lib.KERNEL_GENERIC
    assert lib.cspn_fwd_workspace_bytes(16, 1, 480, 640, 12, 5, 1) == 2 * 16 * 480 * 640 * 4
    assert lib.cspn_fwd_workspace_bytes(16, 1, 480, 640, 4, 5, 1) == 0                 # one launch: no hand-over planes
    tile_hist = 24 * 64 * 64 * 4                                                       # history of one tile: T x 64 x 64 floats
    assert lib.cspn_bwd_workspace_bytes(1, 1, 60, 60, 24, 3, 0) == tile_hist           # one tile, one CTA: one history slot
    n = lib.cspn_bwd_workspace_bytes(8, 1, 228, 304, 24, 3, 0)
    # stream mode: status word + 20 tiles of 64x64 per image, history slots = the persistent grid (148 CTAs for 160 tiles)
    assert n == 148 * tile_hist + 256 + 8 * 20 * (4 * 64 + 4 * 2 * 32) * 16


def test_dual_slot_planner_is_opt_in():
    """The dual-slot forward kernel (CSPN_FWD_KERNEL=dual, read once per process): units of resident tiles, two per CTA."""
    import subprocess
    import sys
    code = ("from cspn_monodepth_b200 import _lib\n"
            "p = _lib.forward_plan(8, 1, 228, 304, 24)\n"
            "assert p == dict(kernel=_lib.KERNEL_DUAL, rows_per_warp=10, cx=5, cy=7, ntx=1, nty=1, ctas=140, rounds=1, units_per_class=4, units=8), p\n"
            "p = _lib.forward_plan(32, 1, 352, 1216, 24)\n"
            "assert (p['kernel'], p['rows_per_warp'], p['ntx'] * p['nty'], p['rounds']) == (_lib.KERNEL_DUAL, 8, 2, 32) and p['ctas'] <= 148, p\n"
            "p = _lib.forward_plan(1, 1, 1080, 1440, 24)\n"
            "assert p['kernel'] == _lib.KERNEL_DUAL and p['cx'] * p['cy'] <= 148 and p['ntx'] * p['nty'] > 1, p\n"
            "assert _lib.forward_plan(3, 1, 97, 131, 24)['kernel'] == _lib.KERNEL_SINGLE\n"
            "n = _lib.load().cspn_fwd_workspace_bytes(8, 1, 228, 304, 24, 3, 0)\n"
            "assert n >= 256 + 2 * 140 * 2 * 2 * (2 * 40 + 128) * 16, n\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CSPN_FWD_KERNEL="dual", PYTHONPATH=root), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_torch_library_layer_registers_the_operators():
    """TORCH_LIBRARY(cspn, ...) (csrc/torch_ext.cpp, built in-tree as lib/libcspn_torch.so): the three operators exist with the
    schemas SURVEY.md 8b lists, and CPU tensors are refused loudly (no CPU implementation anywhere in the product)."""
    import torch
    ops = _lib.torch_ops()
    assert ops is not None, "libcspn_torch.so could not be built / loaded"
    assert str(torch.ops.cspn.propagate.default._schema) == "cspn::propagate(Tensor guidance, Tensor depth, Tensor? sparse, int iters, int ksize, int mode) -> Tensor"
    assert str(torch.ops.cspn.backward.default._schema).endswith("-> (Tensor, Tensor)")
    g, d = torch.zeros(1, 8, 4, 4), torch.zeros(1, 1, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.cspn.forward(g, d, None, 2, 3, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.cspn.backward(d, g, d, None, 2, 3, 0)


def test_pipelined_host_api_argument_validation_without_gpu():
    lib = _lib.load()
    one = ctypes.c_float(0.0)
    p = ctypes.addressof(one)
    t = ctypes.c_int(7)
    assert lib.cspn_fwd_host_submit_f32(p, 8 * 16, p, None, 1, p, 1, 1, 4, 4, 2, 3, 9, ctypes.byref(t)) == -4 and t.value == 0
    assert lib.cspn_fwd_host_submit_f32(p, 8 * 16, p, None, 1, p, 0, 1, 4, 4, 2, 3, 0, ctypes.byref(t)) == 0 and t.value == 0     # empty batch: no ticket
    assert lib.cspn_fwd_host_submit_f32(p, 8 * 16, p, None, 1, p, 1, 1, 4, 4, 2, 3, 0, None) == -1
    assert lib.cspn_host_wait(0) == 0 and lib.cspn_host_pipeline_depth() >= 2
    assert lib.cspn_error_string(-9).startswith(b"cspn:")
