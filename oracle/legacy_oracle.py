"""CPU restatement (numpy fp32) of the reference's legacy max-of-8 CSPN.  TEST INFRASTRUCTURE ONLY: imported by tests/ and
bench.py's checks - never by the product.

  forward(guidance, depth, sparse, iters=16)    network/libs/post_process/CSPN.py:19-56  (AffinityPropagate)
                                                sparse=None: :132-164 (AffinityPropagate_prediction)
  backward(...)                                 autograd through the same loop (max routes the gradient to the winning gate;
                                                torch.max(a, b) splits it evenly on ties, node by node of the tree :113-123)
Pinned by tests/golden/legacy_golden.npz (tests/golden/make_legacy_golden.py runs the reference's classes).

Per step, for each of the 8 gates g_k = |guidance[:, k]| (NOT shifted, unlike CSPN_new):
    out_k(p) = box3x3(g_k * r)(p) / box3x3(g_k)(p)          zero padded, centre included   (:81-102)
    r(p)     = max_k out_k(p)                                NaN (0/0) propagates            (:48-51)
    r        = (1 - m) * r + m * sparse,  m = sign(sparse)   the SPARSE SAMPLE is re-injected (:53), and also seeds r^0 (:34)
"""
import numpy as np


def _box(a):
    """3x3 box sum with zero padding over the last two axes."""
    p = np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 1), (1, 1)])
    h, w = a.shape[-2:]
    out = np.zeros_like(a)
    for dy in range(3):
        for dx in range(3):
            out = out + p[..., dy:dy + h, dx:dx + w]
    return out


def _tree_max(e):
    """max_of_8_tensor (:113-123) with torch.max's NaN propagation; e: [..., 8, H, W] stacked on axis 1."""
    def mx(a, b):
        return np.where(np.isnan(a) | np.isnan(b), np.float32(np.nan), np.maximum(a, b))
    return mx(mx(mx(e[:, 0], e[:, 1]), mx(e[:, 2], e[:, 3])), mx(mx(e[:, 4], e[:, 5]), mx(e[:, 6], e[:, 7])))


def _tree_weights(e):
    """d max_of_8 / d e_k: 1 for the winner, split evenly at every binary node on ties."""
    def node(a, wa, b, wb):
        ga = np.where(a > b, 1.0, np.where(a == b, 0.5, 0.0)).astype(np.float32)
        return np.maximum(a, b), [w * ga for w in wa] + [w * (1 - ga) for w in wb]
    one = np.ones_like(e[:, 0])
    leaves = [(e[:, k], [one]) for k in range(8)]
    l1 = [node(leaves[i][0], leaves[i][1], leaves[i + 1][0], leaves[i + 1][1]) for i in (0, 2, 4, 6)]
    l2 = [node(l1[i][0], l1[i][1], l1[i + 1][0], l1[i + 1][1]) for i in (0, 2)]
    _, w = node(l2[0][0], l2[0][1], l2[1][0], l2[1][1])
    return np.stack(w, axis=1)


def forward(guidance, depth, sparse=None, iters=16, history=None):
    g = np.abs(np.asarray(guidance, np.float32)[:, :8])                       # :22-29
    d = np.asarray(depth, np.float32)[:, 0]
    if sparse is None:
        m = s = np.zeros_like(d)
    else:
        s = np.asarray(sparse, np.float32)[:, 0]
        m = np.sign(s)                                                        # :31
    r = (1 - m) * d + m * s                                                   # :33
    wsum = _box(g)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _ in range(iters):
            if history is not None:
                history.append(r)
            e = _box(g * r[:, None]) / wsum                                   # :99-102
            r = _tree_max(e)
            r = (1 - m) * r + m * s                                           # :53
    return r[:, None].astype(np.float32)


def backward(guidance, depth, sparse, grad_out, iters=16):
    """(grad_guidance [B,Cg,H,W], grad_depth, grad_sparse) for finite problems (float64 accumulation)."""
    guidance = np.asarray(guidance, np.float32)
    g = np.abs(guidance[:, :8]).astype(np.float64)
    hist = []
    forward(guidance, depth, sparse, iters, history=hist)
    s = np.zeros_like(hist[0]) if sparse is None else np.asarray(sparse, np.float32)[:, 0]
    m = np.sign(s).astype(np.float64)
    wsum = _box(g)
    G = np.asarray(grad_out, np.float64)[:, 0]
    gg = np.zeros_like(g)
    gs = np.zeros_like(G)
    for t in range(iters - 1, -1, -1):
        r = hist[t].astype(np.float64)
        gs += m * G
        u = (1 - m) * G
        e32 = (_box(np.abs(guidance[:, :8]) * hist[t][:, None]) / _box(np.abs(guidance[:, :8]))).astype(np.float32)
        sel = _tree_weights(e32).astype(np.float64)                           # which gate won at every pixel
        a = sel * u[:, None] / wsum                                           # d L / d numerator_k(p)
        e = _box(g * r[:, None]) / wsum
        ba = _box(a)                                                          # adjoint of a zero-padded box sum is a box sum
        gg += ba * r[:, None] - _box(a * e)                                   # numerator and denominator terms of g_k(q)
        G = (ba * g).sum(axis=1)
    gs += m * G
    gd = (1 - m) * G
    out_g = np.zeros(guidance.shape, np.float64)
    out_g[:, :8] = np.sign(guidance[:, :8]) * gg
    return out_g.astype(np.float32), gd[:, None].astype(np.float32), gs[:, None].astype(np.float32)
