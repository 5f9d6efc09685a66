"""Op-for-op PyTorch restatement of the reference's CPU path - the TIMED baseline ("port").

TEST / BENCH INFRASTRUCTURE ONLY (same rules as ``oracle/cspn_oracle.py``).  The reference
implements the hot path as a sequence of ATen calls issued from Python; this module re-issues
the same sequence (same ops, same temporaries, same order) so that timing it on the host cores
reproduces what a user of the reference gets on CPU.  It is a restatement, not a copy: the
reference writes the eight shifted pads out by hand, here they come from one offset table.

* mode A: ``network/libs/post_process/CSPN_new.py:26-128`` - per step: 8 x ZeroPad2d + unsqueeze,
  ``cat`` to [B,8,C,H+2,W+2] (:94-119), ``gate * depth``, two ``conv3d`` with a ones
  (1,8,1,1,1) kernel, ``div`` (:121-128), crop (:86-88), blend (:89-90).
* mode B: ``network/libs/post_process/CSPN_ours.py:24-54`` with ``pac.Conv2dFn.forward``
  (``network/libs/base/pac.py:75-94``): ``unfold``, view * kernel, ``einsum`` reduce, ``clone``.

Checked against the reference's golden vectors in ``tests/test_oracle_golden.py``.
"""
import math

import torch
import torch.nn.functional as F

# ZeroPad2d (left, right, top, bottom) per guidance channel, CSPN_new.py:43-67
_PADS = [(l, 2 - l, t, 2 - t) for t in (0, 1, 2) for l in (0, 1, 2) if (l, t) != (1, 1)]


def _eight_shifted(x):
    return torch.cat([F.pad(x, pad).unsqueeze(1) for pad in _PADS], 1)


def mode_a_forward(guidance, blur_depth, sparse_depth=None, prop_time=24):
    gates = _eight_shifted_gates(guidance)
    result = blur_depth
    mask = sparse_depth.sign() if sparse_depth is not None else None
    for _ in range(prop_time):
        stacked = _eight_shifted(result)
        ones = torch.ones((1, 8, 1, 1, 1), device=gates.device)           # rebuilt every step, CSPN_new.py:122
        weight_sum = F.conv3d(gates, ones)                                 # loop-invariant but recomputed, :124
        total = F.conv3d(gates * stacked, ones)                            # :125
        result = torch.div(total, weight_sum).squeeze(1)[:, :, 1:-1, 1:-1]  # :127, :86-87
        if mask is not None:
            result = (1 - mask) * result + mask * blur_depth               # :90
    return result


def _eight_shifted_gates(guidance):
    return torch.cat([F.pad(torch.abs(guidance.narrow(1, k, 1)), pad).unsqueeze(1) for k, pad in enumerate(_PADS)], 1)


def mode_b_forward(x, guided, sparse_depth=None, prop_time=24):
    b, c, h, w = guided.shape
    k = int(math.sqrt(c + 1))
    soft = F.softmax(guided, dim=1)
    kernel = torch.zeros(b, c + 1, h, w, device=guided.device)
    kernel[:, :c // 2] = soft[:, :c // 2]
    kernel[:, c // 2 + 1:] = soft[:, c // 2:]
    kernel = kernel.unsqueeze(1).reshape(b, 1, k, k, h, w)
    mask = sparse_depth.sign() if sparse_depth is not None else None
    x0 = x
    for _ in range(prop_time):
        cols = F.unfold(x, (k, k), 1, k // 2, 1)                            # pac.py:89
        x = torch.einsum("ijklmn->ijmn", cols.view(b, x.shape[1], k, k, h, w) * kernel).clone()  # :91-94
        if mask is not None:
            x = mask * x0 + (1 - mask) * x                                  # CSPN_ours.py:51-53
    return x
