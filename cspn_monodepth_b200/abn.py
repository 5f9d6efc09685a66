"""Drop-ins for the reference's in-place activated batch normalisation, ``network/libs/inplace_abn``.

Same names, constructor arguments, parameters / buffers (``weight``, ``bias``, ``running_mean``, ``running_var``) and
``forward(x)`` as ``bn.py:47-110`` (``InPlaceABN``), ``:113-199`` (``InPlaceABNSync``) and the two ``*Wrapper`` classes
(``:202-221``); the functional forms ``inplace_abn`` / ``inplace_abn_sync`` follow ``functions.py:70-163`` / ``:166-297``:

* training: per-channel mean / biased variance of the batch, running statistics updated with ``momentum`` and the
  ``n / (n - 1)`` correction; eval: the running statistics are used and the backward treats them as constants;
* ``gamma = |weight| + eps``, ``beta = bias``; the activated output OVERWRITES ``x`` (``ctx.mark_dirty``) and is the only
  activation tensor saved - the backward undoes the activation and recovers ``y = (z - beta) / gamma`` from it;
* ``InPlaceABNSync``: statistics over all replicas.  The reference exchanges per-GPU (mean, var) and (edz, eydz) through
  master / worker queues between ``nn.DataParallel`` threads (``functions.py:185-208``, ``:257-276``); here replicas are one
  process per GPU and the exchange is ONE all-reduce (SUM) of the [2C] double vector of per-channel sums in each direction
  (``torch.distributed``, NCCL).  Like the reference's mean of per-replica means (``functions.py:196-197``) this assumes that every
  replica holds the same number of samples (count = local count x world size).  Without an initialised process group it
  degenerates to ``InPlaceABN``, like the reference on one device.  ``devices`` is accepted and ignored.

fp32, CUDA only, contiguous input (the reference raises ``ValueError("Non-contiguous input")``, ``functions.py:65-67``).
Kernels: ``csrc/cspn_abn.cu`` behind ``cspn_abn_*`` (``include/cspn_b200.h``).
"""
from __future__ import annotations

import torch
import torch.autograd as autograd
import torch.distributed as dist
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _lib

ACT_LEAKY_RELU = "leaky_relu"
ACT_ELU = "elu"
ACT_NONE = "none"
_ACT_CODE = {ACT_NONE: 0, ACT_LEAKY_RELU: 1, ACT_ELU: 2}


def _count_samples(x):                                              # functions.py:39-44
    return x.numel() // x.size(1)


def _check_input(x, *others):
    if not x.is_cuda:
        raise RuntimeError(f"x is on {x.device}: the B200 in-place ABN kernels are CUDA-only and have no CPU fallback")
    if x.dtype != torch.float32:
        raise RuntimeError(f"unsupported dtype {x.dtype}: the in-place ABN kernels are float32 (like the reference's bn.cu)")
    if x.dim() < 2:
        raise RuntimeError("x must be [N, C, ...]")
    if not all(t is None or t.is_contiguous() for t in (x,) + others):
        raise ValueError("Non-contiguous input")                   # functions.py:65-67


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _world(group):
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class _InPlaceABN(autograd.Function):
    """``functions.py:70-163`` and, with ``group`` naming a process group, ``:166-297``."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope, group):
        _check_input(x, weight, bias, running_mean, running_var)
        if activation not in _ACT_CODE:
            raise RuntimeError(f"unknown activation {activation!r}: one of leaky_relu, elu, none")
        lib = _lib.load()
        n_, c_ = x.size(0), x.size(1)
        s_ = x.numel() // max(n_ * c_, 1)
        world = _world(group)
        ctx.dims, ctx.training, ctx.eps, ctx.act, ctx.slope, ctx.group, ctx.world = (n_, c_, s_), training, eps, _ACT_CODE[activation], slope, group, world
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if training:
                count = float(_count_samples(x)) * world
                ws = torch.empty(lib.cspn_abn_workspace_bytes(c_), dtype=torch.uint8, device=x.device)
                sums = torch.empty(2 * c_, dtype=torch.float64, device=x.device)
                _lib.check(lib.cspn_abn_stats_f32(x.data_ptr(), n_, c_, s_, sums.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                if world > 1:
                    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
                mean, var = torch.empty(c_, device=x.device), torch.empty(c_, device=x.device)
                _lib.check(lib.cspn_abn_finalize_f32(sums.data_ptr(), count, mean.data_ptr(), var.data_ptr(), _ptr(running_mean), _ptr(running_var),
                                                     float(momentum), c_, stream))
            else:
                mean, var = running_mean, running_var
            _lib.check(lib.cspn_abn_forward_f32(x.data_ptr(), mean.data_ptr(), var.data_ptr(), _ptr(weight), _ptr(bias), n_, c_, s_, float(eps),
                                                ctx.act, float(slope), stream))
        ctx.var = var
        ctx.has_affine = weight is not None
        ctx.save_for_backward(x, weight, bias)
        ctx.mark_dirty(x)
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        z, weight, bias = ctx.saved_tensors
        dz = dz.contiguous()
        lib = _lib.load()
        n_, c_, s_ = ctx.dims
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.has_affine and ctx.needs_input_grad[1], ctx.has_affine and ctx.needs_input_grad[2]
        dx = torch.empty_like(dz) if need_x else None
        dweight = torch.zeros(c_, device=dz.device) if need_w else None
        dbias = torch.zeros(c_, device=dz.device) if need_b else None
        stream = torch.cuda.current_stream(dz.device).cuda_stream
        count_local = float(n_ * s_)
        with torch.cuda.device(dz.device):
            ws = torch.empty(lib.cspn_abn_workspace_bytes(c_), dtype=torch.uint8, device=dz.device)
            sums = torch.empty(2 * c_, dtype=torch.float64, device=dz.device)
            # eval mode: the reference uses edz = eydz = 0 (functions.py:147-150), which also leaves dweight / dbias at zero - kept
            if ctx.training:
                _lib.check(lib.cspn_abn_bwd_reduce_f32(z.data_ptr(), dz.data_ptr(), _ptr(weight), _ptr(bias), n_, c_, s_, float(ctx.eps), ctx.act,
                                                       float(ctx.slope), sums.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                if ctx.world > 1:
                    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
                sums_ptr = sums.data_ptr()
            else:
                sums_ptr = 0
            _lib.check(lib.cspn_abn_bwd_apply_f32(z.data_ptr(), dz.data_ptr(), _ptr(dx), ctx.var.data_ptr(), _ptr(weight), _ptr(bias), sums_ptr,
                                                  count_local * ctx.world, count_local, _ptr(dweight) if sums_ptr else 0, _ptr(dbias) if sums_ptr else 0,
                                                  n_, c_, s_, float(ctx.eps), ctx.act, float(ctx.slope), stream))
        del ctx.var
        return dx, dweight, dbias, None, None, None, None, None, None, None, None


def inplace_abn(x, weight, bias, running_mean, running_var, training=True, momentum=0.1, eps=1e-05, activation=ACT_LEAKY_RELU, slope=0.01):
    """``functions.py:70`` (``InPlaceABN.apply``)."""
    return _InPlaceABN.apply(x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope, False)


def inplace_abn_sync(x, weight, bias, running_mean, running_var, extra=None, training=True, momentum=0.1, eps=1e-05, activation=ACT_LEAKY_RELU,
                     slope=0.01):
    """``functions.py:166`` (``InPlaceABNSync.apply``).  ``extra`` carried the reference's queues; here it may carry
    ``{"group": process_group}`` (default: the world group)."""
    group = (extra or {}).get("group") if isinstance(extra, dict) else None
    return _InPlaceABN.apply(x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope, group)


class ABN(nn.Sequential):
    """``bn.py:23-44``: plain ``nn.BatchNorm2d`` + activation module (no custom kernel; exported because the reference's package does)."""

    def __init__(self, num_features, activation=None, **kwargs):
        from collections import OrderedDict
        super().__init__(OrderedDict([("bn", nn.BatchNorm2d(num_features, **kwargs)),
                                      ("act", activation if activation is not None else nn.ReLU(inplace=True))]))


class InPlaceABN(nn.Module):
    """``bn.py:47-110``."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu", slope=0.01):
        super().__init__()
        self.num_features = num_features
        self.affine = affine
        self.eps = eps
        self.momentum = momentum
        self.activation = activation
        self.slope = slope
        if self.affine:
            self.weight = nn.Parameter(torch.Tensor(num_features))
            self.bias = nn.Parameter(torch.Tensor(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.reset_parameters()

    def reset_parameters(self):
        self.running_mean.zero_()
        self.running_var.fill_(1)
        if self.affine:
            self.weight.data.fill_(1)
            self.bias.data.zero_()

    def forward(self, x):
        return inplace_abn(x, self.weight, self.bias, self.running_mean, self.running_var, self.training, self.momentum, self.eps, self.activation,
                           self.slope)

    def extra_repr(self):
        rep = "{num_features}, eps={eps}, momentum={momentum}, affine={affine}, activation={activation}"
        if self.activation == ACT_LEAKY_RELU:
            rep += ", slope={slope}"
        return rep.format(**self.__dict__)


class InPlaceABNSync(InPlaceABN):
    """``bn.py:113-199``: statistics over every replica (one all-reduce of the per-channel sums per direction)."""

    def __init__(self, num_features, devices=None, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu", slope=0.01, process_group=None):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine, activation=activation, slope=slope)
        self.devices = devices
        self.process_group = process_group

    def forward(self, x):
        return inplace_abn_sync(x, self.weight, self.bias, self.running_mean, self.running_var, {"group": self.process_group}, self.training,
                                self.momentum, self.eps, self.activation, self.slope)


class InPlaceABNWrapper(nn.Module):
    """``bn.py:202-210``."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.bn = InPlaceABN(*args, **kwargs)

    def forward(self, input):
        return self.bn(input)


class InPlaceABNSyncWrapper(nn.Module):
    """``bn.py:213-221``."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.bn = InPlaceABNSync(*args, **kwargs)

    def forward(self, input):
        return self.bn(input)
