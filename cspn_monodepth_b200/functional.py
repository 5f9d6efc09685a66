"""Host side of the CSPN hot path: a ``torch.autograd.Function`` over the C ABI.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic happens in
``libcspn_b200.so``.  The function is stateless and uses the *current* device and stream of the input
tensors, so it is safe under the reference's thread-per-GPU ``DataParallelModel``
(``network/libs/base/encoding.py:102-105``) and inside CUDA graph capture.
"""
from __future__ import annotations

import math

import torch

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float16: "f16"}


def _check_inputs(guidance, depth, sparse, taps):
    for name, t in (("guidance", guidance), ("depth", depth), ("sparse_depth", sparse)):
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise RuntimeError(f"{name} is on {t.device}: the B200 CSPN operator is CUDA-only and has no CPU fallback")
        if t.device != guidance.device:
            raise RuntimeError("guidance, depth and sparse_depth must be on the same device")
        if t.dtype != guidance.dtype:
            raise RuntimeError(f"dtype mismatch: {name} is {t.dtype}, guidance is {guidance.dtype}")
    if guidance.dtype not in _SUFFIX:
        raise RuntimeError(f"unsupported dtype {guidance.dtype}: the CSPN operator supports float32 and float16")
    if guidance.dim() != 4 or depth.dim() != 4:
        raise RuntimeError("guidance and depth must be 4-D NCHW tensors")
    b, c, h, w = depth.shape
    if guidance.shape[0] != b or tuple(guidance.shape[2:]) != (h, w):
        raise RuntimeError(f"guidance {tuple(guidance.shape)} does not match depth {tuple(depth.shape)}")
    if guidance.shape[1] < taps:
        raise RuntimeError(f"guidance has {guidance.shape[1]} channels, the propagation kernel needs {taps}")
    if sparse is not None:
        if sparse.dim() != 4 or sparse.shape[0] != b or tuple(sparse.shape[2:]) != (h, w) or sparse.shape[1] not in (1, c):
            raise RuntimeError(f"sparse_depth {tuple(sparse.shape)} does not match depth {tuple(depth.shape)}")


def _guidance_view(g, h, w):
    """Guidance may be a channel-narrowed view of a wider tensor: only its batch stride is free."""
    if g.stride(3) == 1 and g.stride(2) == w and g.stride(1) == h * w and g.stride(0) >= g.shape[1] * h * w:
        return g, g.stride(0)
    g = g.contiguous()
    return g, g.shape[1] * h * w


def _ptr(t):
    return None if t is None else t.data_ptr()


class _CspnPropagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, guidance, depth, sparse, iters, ksize, mode):
        taps = ksize * ksize - 1
        _check_inputs(guidance, depth, sparse, taps)
        lib = _lib.load()
        b, c, h, w = depth.shape
        g, gbs = _guidance_view(guidance, h, w)
        d = depth.contiguous()
        s = None if sparse is None else sparse.contiguous()
        out = torch.empty_like(d)
        sfx = _SUFFIX[d.dtype]
        with torch.cuda.device(d.device):
            nbytes = lib.cspn_fwd_workspace_bytes(b, c, h, w, iters, ksize, mode)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device) if nbytes else None
            stream = torch.cuda.current_stream(d.device).cuda_stream
            _lib.check(getattr(lib, "cspn_fwd_" + sfx)(
                g.data_ptr(), gbs, d.data_ptr(), _ptr(s), 1 if s is None else s.shape[1], out.data_ptr(),
                b, c, h, w, iters, ksize, mode, _ptr(ws), nbytes, stream))
        ctx.save_for_backward(guidance, depth, sparse)
        ctx.cfg = (iters, ksize, mode)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        guidance, depth, sparse = ctx.saved_tensors
        iters, ksize, mode = ctx.cfg
        lib = _lib.load()
        b, c, h, w = depth.shape
        g, gbs = _guidance_view(guidance, h, w)
        d = depth.contiguous()
        s = None if sparse is None else sparse.contiguous()
        go = grad_out.contiguous()
        gg = torch.empty(guidance.shape, dtype=guidance.dtype, device=guidance.device)
        gd = torch.empty_like(d)
        sfx = _SUFFIX[d.dtype]
        with torch.cuda.device(d.device):
            nbytes = lib.cspn_bwd_workspace_bytes(b, c, h, w, iters, ksize, mode)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device) if nbytes else None
            stream = torch.cuda.current_stream(d.device).cuda_stream
            _lib.check(getattr(lib, "cspn_bwd_" + sfx)(
                go.data_ptr(), g.data_ptr(), gbs, guidance.shape[1], d.data_ptr(), _ptr(s),
                1 if s is None else s.shape[1], gg.data_ptr(), gd.data_ptr(),
                b, c, h, w, iters, ksize, mode, _ptr(ws), nbytes, stream))
        return gg, gd, None, None, None, None


def cspn_propagate(guidance, depth, sparse_depth=None, *, iters: int, ksize: int = 3, mode: int = _lib.MODE_NEW):
    """``iters`` CSPN steps of ``depth`` under ``guidance`` with optional sparse re-injection.

    mode ``MODE_NEW``  = ``CSPN_new.AffinityPropagate`` semantics (reference ``CSPN_new.py:26-92``);
    mode ``MODE_OURS`` = ``CSPN_ours.AffinityPropagate`` semantics (``CSPN_ours.py:24-54``).
    ``iters == 0`` returns ``depth`` itself, like both reference loops do.
    """
    if iters == 0:
        return depth
    ops = _lib.torch_ops()
    if ops is not None:
        # C++ operator layer (TORCH_LIBRARY(cspn, ...), csrc/torch_ext.cpp): same checks, allocation and autograd formula
        # as _CspnPropagate below, without the ctypes / Python overhead per call
        for name, t in (("guidance", guidance), ("depth", depth), ("sparse_depth", sparse_depth)):
            if t is None:
                continue
            if not isinstance(t, torch.Tensor):
                raise TypeError(f"{name} must be a torch.Tensor")
            if not t.is_cuda:
                raise RuntimeError(f"{name} is on {t.device}: the B200 CSPN operator is CUDA-only and has no CPU fallback")
        return ops.propagate(guidance, depth, sparse_depth, int(iters), int(ksize), int(mode))
    return _CspnPropagate.apply(guidance, depth, sparse_depth, int(iters), int(ksize), int(mode))


def kernel_size_from_channels(channels: int) -> int:
    """K = int(sqrt(C + 1)) as in ``CSPN_ours.py:32``; rejects what the reference's reshape (:41) rejects."""
    k = int(math.sqrt(channels + 1))
    if k * k != channels + 1:
        raise RuntimeError(f"guided has {channels} channels; K*K-1 channels are required (shape invalid for K={k})")
    return k
