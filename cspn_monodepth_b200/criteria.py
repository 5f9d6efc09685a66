"""Drop-ins for what sits directly downstream of the CSPN module in the reference.

* :class:`MaskedL1Loss` - same name and ``forward(pred, target)`` as ``libs/criterion/criteria.py:27-39`` (the criterion the
  reference trains with through ``Criterion_No_DSN``, ``:170-188``): mean ``|target - pred|`` over ``target > 0``.
* :func:`evaluate` / :class:`Result` - ``libs/metrics.py:49-83`` (``Result.evaluate``): irmse, imae, mse, rmse, mae, absrel,
  lg10, delta1-3 over ``target > 0``.

Each is ONE streaming pass over (pred, target) in ``libcspn_b200.so`` (``csrc/cspn_loss.cu``) with a deterministic reduction,
instead of the reference's ~6 / ~25 ATen ops with boolean-mask gathers and a host synchronisation per ``float()``.
CUDA tensors only - like the module itself there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float16: "f16"}
_scratch = {}


def _workspace(device):
    """Per device and stream: the reduction scratch (zeroed once; the kernels leave it ready for the next call)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _scratch.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().cspn_loss_workspace_bytes(), dtype=torch.uint8, device=device)
        _scratch[key] = ws
    return ws


def _check(pred, target):
    if pred.dim() != target.dim():
        raise AssertionError("inconsistent dimensions")                  # criteria.py:32
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError(f"pred / target are on {pred.device} / {target.device}: the B200 loss kernels are CUDA-only and have no CPU fallback")
    if pred.shape != target.shape or pred.dtype != target.dtype or pred.dtype not in _SUFFIX:
        raise RuntimeError(f"pred {tuple(pred.shape)} {pred.dtype} and target {tuple(target.shape)} {target.dtype} must match (float32 or float16)")


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        _check(pred, target)
        lib = _lib.load()
        p, t = pred.contiguous(), target.contiguous()
        with torch.cuda.device(p.device):
            ws = _workspace(p.device)
            loss2 = torch.empty(2, dtype=torch.float32, device=p.device)
            _lib.check(getattr(lib, "cspn_masked_l1_fwd_" + _SUFFIX[p.dtype])(
                p.data_ptr(), t.data_ptr(), p.numel(), loss2.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(p.device).cuda_stream))
        ctx.save_for_backward(p, t, loss2)
        return loss2[0].to(pred.dtype) if pred.dtype != torch.float32 else loss2[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        p, t, loss2 = ctx.saved_tensors
        lib = _lib.load()
        gp = torch.empty_like(p)
        gl = grad_loss.to(torch.float32).contiguous()
        with torch.cuda.device(p.device):
            _lib.check(getattr(lib, "cspn_masked_l1_bwd_" + _SUFFIX[p.dtype])(
                p.data_ptr(), t.data_ptr(), p.numel(), loss2.data_ptr(), gl.data_ptr(), gp.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream))
        return gp, None


class MaskedL1Loss(nn.Module):
    """``libs/criterion/criteria.py:27-39``."""

    def __init__(self):
        super().__init__()

    def forward(self, pred, target):
        self.loss = _MaskedL1.apply(pred, target)
        return self.loss


METRIC_NAMES = ("irmse", "imae", "mse", "rmse", "mae", "absrel", "lg10", "delta1", "delta2", "delta3")


def evaluate_device(output, target):
    """The 10 metrics of ``Result.evaluate`` + the valid-pixel count as an 11-element device tensor (no host sync)."""
    _check(output, target)
    lib = _lib.load()
    o, t = output.contiguous(), target.contiguous()
    with torch.cuda.device(o.device):
        ws = _workspace(o.device)
        out = torch.empty(11, dtype=torch.float32, device=o.device)
        _lib.check(getattr(lib, "cspn_depth_metrics_" + _SUFFIX[o.dtype])(
            o.data_ptr(), t.data_ptr(), o.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(o.device).cuda_stream))
    return out


def evaluate(output, target):
    """``Result.evaluate`` as a dict of Python floats (one device-to-host copy instead of one per metric)."""
    vals = evaluate_device(output, target).tolist()
    return dict(zip(METRIC_NAMES + ("count",), vals))


class Result:
    """Same attributes as the reference's ``libs/metrics.py`` ``Result`` (:19-47), ``evaluate`` on the B200 kernel."""

    def __init__(self):
        self.irmse = self.imae = self.mse = self.rmse = self.mae = self.absrel = self.lg10 = 0
        self.delta1 = self.delta2 = self.delta3 = 0
        self.data_time = self.gpu_time = 0
        self.loss = 0

    def evaluate(self, output, target, loss=None):
        if output.shape[2:] != target.shape[2:]:                             # metrics.py:54-55
            output = torch.nn.functional.interpolate(output, size=target.shape[2:], mode="bilinear", align_corners=True)
        m = evaluate(output, target)
        for k in METRIC_NAMES:
            setattr(self, k, m[k])
        self.rmse = math.sqrt(self.mse) if self.mse == self.mse else self.rmse
        self.data_time = self.gpu_time = 0
        if loss:
            self.loss = loss
