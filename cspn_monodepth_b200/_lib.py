"""ctypes binding of ``lib/libcspn_b200.so`` (C ABI declared in ``include/cspn_b200.h``).

There is deliberately no fallback: if the CUDA library is missing and cannot be built,
importing the operators raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

_c_int, _c_i64, _c_sz, _c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/cspn_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "cspn_abi_version": (_c_int, []),
    "cspn_error_string": (ctypes.c_char_p, [_c_int]),
    "cspn_set_path": (_c_int, [_c_int]),
    "cspn_last_path": (_c_int, []),
    "cspn_last_launch_count": (_c_int, []),
    "cspn_fwd_workspace_bytes": (_c_sz, [_c_int] * 7),
    "cspn_bwd_workspace_bytes": (_c_sz, [_c_int] * 7),
    "cspn_fwd_plan": (_c_int, [_c_int] * 7 + [ctypes.POINTER(_c_int)]),
    "cspn_fwd_f32": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [_c_vp, _c_sz, _c_vp]),
    "cspn_fwd_f16": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [_c_vp, _c_sz, _c_vp]),
    "cspn_bwd_f32": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_vp, _c_vp] + [_c_int] * 7 + [_c_vp, _c_sz, _c_vp]),
    "cspn_bwd_f16": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_vp, _c_vp] + [_c_int] * 7 + [_c_vp, _c_sz, _c_vp]),
    "cspn_fwd_host_f32": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [_c_vp]),
    "cspn_fwd_host_f16": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [_c_vp]),
    "cspn_fwd_host_submit_f32": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [ctypes.POINTER(_c_int)]),
    "cspn_fwd_host_submit_f16": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_int, _c_vp] + [_c_int] * 7 + [ctypes.POINTER(_c_int)]),
    "cspn_host_wait": (_c_int, [_c_int]),
    "cspn_abn_workspace_bytes": (_c_sz, [_c_int]),
    "cspn_abn_stats_f32": (_c_int, [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_abn_finalize_f32": (_c_int, [_c_vp, ctypes.c_double, _c_vp, _c_vp, _c_vp, _c_vp, ctypes.c_float, _c_int, _c_vp]),
    "cspn_abn_forward_f32": (_c_int, [_c_vp] * 5 + [_c_int] * 3 + [ctypes.c_float, _c_int, ctypes.c_float, _c_vp]),
    "cspn_abn_bwd_reduce_f32": (_c_int, [_c_vp] * 4 + [_c_int] * 3 + [ctypes.c_float, _c_int, ctypes.c_float, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_abn_bwd_apply_f32": (_c_int, [_c_vp] * 7 + [ctypes.c_double, ctypes.c_double, _c_vp, _c_vp] + [_c_int] * 3 + [ctypes.c_float, _c_int, ctypes.c_float, _c_vp]),
    "cspn_heads_workspace_bytes": (_c_sz, []),
    "cspn_heads_fwd_f32": (_c_int, [_c_vp] * 5 + [_c_int] * 8 + [_c_vp]),
    "cspn_heads_fwd_f16": (_c_int, [_c_vp] * 5 + [_c_int] * 8 + [_c_vp]),
    "cspn_heads_bwd_f32": (_c_int, [_c_vp] * 8 + [_c_int] * 8 + [_c_vp, _c_sz, _c_vp]),
    "cspn_heads_bwd_f16": (_c_int, [_c_vp] * 8 + [_c_int] * 8 + [_c_vp, _c_sz, _c_vp]),
    "cspn_legacy_workspace_bytes": (_c_sz, [_c_int] * 4),
    "cspn_legacy_fwd_f32": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp] + [_c_int] * 4 + [_c_vp, _c_sz, _c_vp]),
    "cspn_legacy_fwd_f16": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp] + [_c_int] * 4 + [_c_vp, _c_sz, _c_vp]),
    "cspn_loss_workspace_bytes": (_c_sz, []),
    "cspn_masked_l1_fwd_f32": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_masked_l1_fwd_f16": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_masked_l1_bwd_f32": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "cspn_masked_l1_bwd_f16": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "cspn_depth_metrics_f32": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_depth_metrics_f16": (_c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_sz, _c_vp]),
    "cspn_host_pipeline_depth": (_c_int, []),
}

PATH_AUTO, PATH_GENERIC, PATH_FUSED, PATH_BLOCKED = 0, 1, 2, 3
MODE_NEW, MODE_OURS = 0, 1

_lock = threading.RLock()        # torch_ops() loads the C-ABI library while holding it
_lib = None


class CspnError(RuntimeError):
    """Raised for a non-zero return code of the C ABI (RuntimeError like the reference's ATen errors)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{message} (code {code})")
        self.code = code


def library_path() -> str:
    return _build.SO


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc exists) the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        so = _build.SO
        try:
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # no nvcc on this machine: fall through to the prebuilt file if there is one
            if not os.path.exists(so):
                raise RuntimeError(
                    "libcspn_b200.so is missing and could not be built - the CSPN operators have no fallback. "
                    f"Run `python -m cspn_monodepth_b200.build` on a machine with nvcc. Cause: {exc}") from exc
        lib = ctypes.CDLL(so)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = ABI mismatch, fail loudly
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


KERNEL_GENERIC, KERNEL_SINGLE, KERNEL_DUAL, KERNEL_BLOCKED = 0, 1, 2, 3


def forward_plan(b: int, c: int, h: int, w: int, iters: int, ksize: int = 3, mode: int = MODE_NEW) -> dict:
    """What the planner picks for a forward problem on the current device (``cspn_fwd_plan``)."""
    buf = (_c_int * 10)()
    check(load().cspn_fwd_plan(b, c, h, w, iters, ksize, mode, buf))
    keys = ("kernel", "rows_per_warp", "cx", "cy", "ntx", "nty", "ctas", "rounds", "units_per_class", "units")
    plan = dict(zip(keys, list(buf)))
    if plan["kernel"] == KERNEL_SINGLE:                  # the second word is the halo transport (CSPN_TRANSPORT_*), not a tile height
        plan["transport"] = ("cluster", "stream", "hybrid")[plan.pop("rows_per_warp")]
    return plan


_torch_ext = None


def torch_ops():
    """``torch.ops.cspn`` (the TORCH_LIBRARY layer, csrc/torch_ext.cpp) or None when ``CSPN_TORCH_EXT=0`` or it cannot be
    built / loaded (no g++ and no prebuilt file): the operators then go through ctypes - same library, same kernels."""
    global _torch_ext
    if _torch_ext is None:
        with _lock:
            if _torch_ext is None:
                _torch_ext = False
                if os.environ.get("CSPN_TORCH_EXT", "1") != "0":
                    try:
                        import torch
                        try:
                            if _build.torch_ext_needs_build():
                                _build.build_torch_ext()
                        except Exception:
                            if not os.path.exists(_build.TORCH_SO):
                                raise
                        load()                                          # libcspn_b200.so first (rpath $ORIGIN resolves it anyway)
                        torch.ops.load_library(_build.TORCH_SO)
                        _torch_ext = torch.ops.cspn
                    except Exception as exc:                            # noqa: BLE001 - optional layer
                        import warnings
                        warnings.warn(f"libcspn_torch.so unavailable ({exc}); using the ctypes binding of libcspn_b200.so")
    return _torch_ext or None


def check(code: int) -> None:
    if code != 0:
        raise CspnError(code, load().cspn_error_string(code).decode())
