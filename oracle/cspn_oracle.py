"""CPU oracle for the CSPN affinity-propagation hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY. Nothing under ``cspn_monodepth_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the timed
baseline - never as the product path.

Parity pin: the reference has no golden vectors of its own, so
this restatement is pinned against outputs of the reference modules themselves,
generated in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function here against those files.

What is restated (paths relative to the reference checkout):

* mode A ("abs / neighbour-indexed"): ``network/libs/post_process/CSPN_new.py:26-128``
  - ``AffinityPropagate.forward`` with ``pad_blur_depth`` and ``eight_way_propagation``.
* mode B ("softmax / centre-indexed", K in {3, 5, ...}):
  ``network/libs/post_process/CSPN_ours.py:24-54`` on top of the pixel-adaptive
  convolution ``network/libs/base/pac.py:75-94`` (forward) and ``:96-121`` (backward).

Arithmetic is IEEE fp32 by default, summed in the reference's order (k = 0..7 for the
numerator, then the denominator, then one divide - ``CSPN_new.py:124-127``); pass
``dtype=np.float64`` for a higher-precision ground truth.
"""
from __future__ import annotations

import numpy as np

# Neighbour offset (dy, dx) read by guidance channel k in mode A.  Derived from the
# ZeroPad2d((l, r, t, b)) shifts at CSPN_new.py:43-67 followed by the [1:-1, 1:-1] crop
# at CSPN_new.py:87: a plane padded by (l, ., t, .) holds source pixel
# (y + 1 - t, x + 1 - l) at cropped position (y, x).
MODE_A_OFFSETS = (
    (+1, +1),  # ch 0  pad (0,2,0,2)
    (+1, 0),   # ch 1  pad (1,1,0,2)
    (+1, -1),  # ch 2  pad (2,0,0,2)
    (0, +1),   # ch 3  pad (0,2,1,1)
    (0, -1),   # ch 4  pad (2,0,1,1)
    (-1, +1),  # ch 5  pad (0,2,2,0)
    (-1, 0),   # ch 6  pad (1,1,2,0)
    (-1, -1),  # ch 7  pad (2,0,2,0)
)


def mode_b_offsets(ksize: int):
    """Offsets (dy, dx) of the K*K-1 softmax channels of mode B, channel order.

    ``CSPN_ours.py:37-41`` writes channels 0..C/2-1 to kernel taps 0..C/2-1, leaves
    the centre tap C/2 at zero and writes channels C/2..C-1 to taps C/2+1..C; taps are
    row-major over (K, K) and tap (iy, ix) reads input pixel (y + iy - P, x + ix - P)
    (``F.unfold`` with padding P = K // 2, ``pac.py:89``).
    """
    p = ksize // 2
    taps = [(iy - p, ix - p) for iy in range(ksize) for ix in range(ksize)]
    centre = (ksize * ksize) // 2
    return tuple(taps[:centre] + taps[centre + 1:])


def _shift(plane: np.ndarray, dy: int, dx: int) -> np.ndarray:
    """out[..., y, x] = plane[..., y + dy, x + dx], zero outside the image."""
    h, w = plane.shape[-2:]
    out = np.zeros_like(plane)
    ys0, ys1 = max(0, -dy), min(h, h - dy)
    xs0, xs1 = max(0, -dx), min(w, w - dx)
    if ys1 > ys0 and xs1 > xs0:
        out[..., ys0:ys1, xs0:xs1] = plane[..., ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
    return out


def _mask_of(sparse, like, dtype):
    if sparse is None:
        return None
    return np.sign(np.asarray(sparse, dtype=dtype))  # CSPN_new.py:77-78 / CSPN_ours.py:43-44


# --------------------------------------------------------------------------- mode A
def mode_a_weights(guidance: np.ndarray, dtype=np.float32):
    """Per-pixel gathered weights W_k(p) = |g_k(p + o_k)| and their sum S(p).

    Follows the gate preparation at ``CSPN_new.py:29-70`` (abs, shifted zero pads,
    cat) and the denominator of ``eight_way_propagation`` (``:124``).
    Returns ``W`` of shape [B, 8, 1, H, W] and ``S`` of shape [B, 1, H, W].
    """
    g = np.asarray(guidance, dtype=dtype)
    a = np.abs(g[:, :8])
    w = np.stack([_shift(a[:, k], dy, dx) for k, (dy, dx) in enumerate(MODE_A_OFFSETS)], axis=1)
    s = np.zeros_like(w[:, 0])
    for k in range(8):
        s = s + w[:, k]
    return w[:, :, None], s[:, None]


def mode_a_forward(guidance, blur_depth, sparse_depth=None, prop_time=24, dtype=np.float32,
                   return_all=False):
    """``CSPN_new.AffinityPropagate(prop_time, 3).forward`` (``CSPN_new.py:26-92``).

    guidance [B, >=8, H, W], blur_depth [B, C, H, W], sparse_depth [B, 1|C, H, W] or None.
    """
    d0 = np.asarray(blur_depth, dtype=dtype)
    w, s = mode_a_weights(guidance, dtype)
    m = _mask_of(sparse_depth, d0, dtype)
    r = d0
    hist = [r]
    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(prop_time):
            num = np.zeros_like(r)
            for k, (dy, dx) in enumerate(MODE_A_OFFSETS):
                num = num + w[:, k] * _shift(r, dy, dx)
            r = num / s
            if m is not None:
                r = (1 - m) * r + m * d0
            hist.append(r)
    return hist if return_all else r


def mode_a_backward(guidance, blur_depth, sparse_depth, grad_out, prop_time=24, dtype=np.float64):
    """Closed-form gradients of :func:`mode_a_forward`.

    Stands in for autograd through ``CSPN_new.py:80-90``.  Returns
    ``(grad_guidance [B, Cg, H, W], grad_blur_depth [B, C, H, W])``; guidance channels
    >= 8 get exactly zero.
    """
    g = np.asarray(guidance, dtype=dtype)
    d0 = np.asarray(blur_depth, dtype=dtype)
    go = np.asarray(grad_out, dtype=dtype)
    w, s = mode_a_weights(g, dtype)
    m = _mask_of(sparse_depth, d0, dtype)
    one_minus_m = 1 if m is None else (1 - m)
    with np.errstate(divide="ignore", invalid="ignore"):
        n = w / s[:, None]                                       # [B, 8, 1, H, W]
        hist = mode_a_forward(g, d0, sparse_depth, prop_time, dtype, return_all=True)
        gn = np.zeros((g.shape[0], 8) + d0.shape[1:], dtype=dtype)  # dL/dn_k per depth channel
        gd0 = np.zeros_like(d0)
        gt = go
        for t in range(prop_time - 1, -1, -1):
            u = one_minus_m * gt
            if m is not None:
                gd0 = gd0 + m * gt
            nxt = np.zeros_like(gt)
            for k, (dy, dx) in enumerate(MODE_A_OFFSETS):
                gn[:, k] += u * _shift(hist[t], dy, dx)
                nxt = nxt + _shift(u * n[:, k], -dy, -dx)        # scatter p -> p + o_k
            gt = nxt
        gd0 = gd0 + gt
        gn = gn.sum(axis=2)                                       # shared weights over C
        nn = n[:, :, 0]
        dot = (nn * gn).sum(axis=1, keepdims=True)
        gw = (gn - dot) / s                                       # dL/dW_k(p)
    gg = np.zeros_like(g)
    for k, (dy, dx) in enumerate(MODE_A_OFFSETS):
        ga = _shift(gw[:, k], -dy, -dx)                           # a_k(q) feeds W_k(q - o_k)
        gg[:, k] = np.sign(g[:, k]) * ga
    return gg, gd0


# --------------------------------------------------------------------------- mode B
def _softmax(x, axis):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def mode_b_forward(x, guided, sparse_depth=None, prop_time=24, dtype=np.float32, return_all=False):
    """``CSPN_ours.AffinityPropagate(prop_time).forward(x, guided, sparse_depth)``.

    ``CSPN_ours.py:24-54`` with ``pac.conv2d`` = ``Conv2dFn.forward`` (``pac.py:75-94``):
    softmax over the K*K-1 guidance channels, centre tap zero, weights read at the
    centre pixel, zero padding without renormalisation.
    """
    d0 = np.asarray(x, dtype=dtype)
    g = np.asarray(guided, dtype=dtype)
    c = g.shape[1]
    ksize = int(np.sqrt(c + 1))
    assert ksize * ksize == c + 1, "guided must have K*K-1 channels"
    offs = mode_b_offsets(ksize)
    s = _softmax(g, axis=1).astype(dtype)
    m = _mask_of(sparse_depth, d0, dtype)
    r = d0
    hist = [r]
    for _ in range(prop_time):
        acc = np.zeros_like(r)
        for j, (dy, dx) in enumerate(offs):
            acc = acc + s[:, j:j + 1] * _shift(r, dy, dx)
        r = acc
        if m is not None:
            r = m * d0 + (1 - m) * r
        hist.append(r)
    return hist if return_all else r


def mode_b_backward(x, guided, sparse_depth, grad_out, prop_time=24, dtype=np.float64):
    """Closed-form gradients of :func:`mode_b_forward` (``pac.py:96-121`` iterated, plus the
    softmax Jacobian of ``CSPN_ours.py:35``).  Returns ``(grad_x, grad_guided)``."""
    d0 = np.asarray(x, dtype=dtype)
    g = np.asarray(guided, dtype=dtype)
    go = np.asarray(grad_out, dtype=dtype)
    c = g.shape[1]
    ksize = int(np.sqrt(c + 1))
    offs = mode_b_offsets(ksize)
    s = _softmax(g, axis=1)
    m = _mask_of(sparse_depth, d0, dtype)
    one_minus_m = 1 if m is None else (1 - m)
    hist = mode_b_forward(d0, g, sparse_depth, prop_time, dtype, return_all=True)
    gs = np.zeros_like(s)
    gd0 = np.zeros_like(d0)
    gt = go
    for t in range(prop_time - 1, -1, -1):
        u = one_minus_m * gt
        if m is not None:
            gd0 = gd0 + m * gt
        nxt = np.zeros_like(gt)
        for j, (dy, dx) in enumerate(offs):
            gs[:, j] += (u * _shift(hist[t], dy, dx)).sum(axis=1)
            nxt = nxt + _shift(u * s[:, j:j + 1], -dy, -dx)
        gt = nxt
    gd0 = gd0 + gt
    gz = s * (gs - (s * gs).sum(axis=1, keepdims=True))
    return gd0, gz
