"""ctypes binding of the C oracle (``oracle/cspn_oracle.c``).  TEST INFRASTRUCTURE ONLY -
same import rules as ``oracle/cspn_oracle.py``."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cspn_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        _LIB.cspn_oracle_forward.restype = ctypes.c_int
        _LIB.cspn_oracle_forward.argtypes = [f32p, ctypes.c_int64, f32p, f32p, ctypes.c_int, f32p] + [ctypes.c_int] * 8
        _LIB.cspn_oracle_backward.restype = ctypes.c_int
        _LIB.cspn_oracle_backward.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, f32p, ctypes.c_int, f32p, f32p, f32p] + [ctypes.c_int] * 8
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _prep(guidance, depth, sparse):
    g = np.ascontiguousarray(guidance, dtype=np.float32)
    d = np.ascontiguousarray(depth, dtype=np.float32)
    s = None if sparse is None else np.ascontiguousarray(sparse, dtype=np.float32)
    return g, d, s


def forward(guidance, depth, sparse=None, iters=24, ksize=3, mode=0, threads=0):
    """mode 0 = CSPN_new (abs/neighbour), mode 1 = CSPN_ours (softmax/centre)."""
    g, d, s = _prep(guidance, depth, sparse)
    b, c, h, w = d.shape
    out = np.empty_like(d)
    rc = lib().cspn_oracle_forward(_p(g), g.shape[1] * h * w, _p(d), _p(s), 1 if s is None else s.shape[1],
                                   _p(out), b, c, h, w, iters, ksize, mode, threads)
    if rc != 0:
        raise ValueError("cspn_oracle_forward: bad arguments")
    return out


def backward(guidance, depth, sparse, grad_out, iters=24, ksize=3, mode=0, threads=0):
    """Returns (grad_guidance, grad_depth)."""
    g, d, s = _prep(guidance, depth, sparse)
    go = np.ascontiguousarray(grad_out, dtype=np.float32)
    b, c, h, w = d.shape
    gg = np.empty_like(g)
    gd = np.empty_like(d)
    rc = lib().cspn_oracle_backward(_p(g), g.shape[1] * h * w, g.shape[1], _p(d), _p(s),
                                    1 if s is None else s.shape[1], _p(go), _p(gg), _p(gd),
                                    b, c, h, w, iters, ksize, mode, threads)
    if rc != 0:
        raise ValueError("cspn_oracle_backward: bad arguments")
    return gg, gd
