"""Generate golden input/output vectors by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it imports ``/root/reference``, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

Writes ``tests/golden/cspn_golden.npz``.  The reference modules are imported
unmodified: ``network/libs/post_process/CSPN_new.py`` (mode A) and
``network/libs/post_process/CSPN_ours.py`` (mode B).  Mode B lazily imports
``network/libs/base/pac.py`` whose line 20 needs ``torch._thnn`` (removed from torch);
a shim module providing ``type2backend[...].Im2Col_updateGradInput`` via ``F.fold`` is
injected so the reference's own ``Conv2dFn.forward/backward`` run as written.
Gradients are the reference's autograd results for a fixed random ``grad_out``.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("CSPN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _install_thnn_shim():
    class _Backend:
        library_state = None

        @staticmethod
        def Im2Col_updateGradInput(state, gcol, gin, ih, iw, kh, kw, dh, dw, ph, pw, sh, sw):
            out = F.fold(gcol, (ih, iw), (kh, kw), dilation=(dh, dw), padding=(ph, pw), stride=(sh, sw))
            gin.resize_(out.shape).copy_(out)

    class _T2B(dict):
        def __missing__(self, key):
            return _Backend

    mod = types.ModuleType("torch._thnn")
    mod.type2backend = _T2B()
    sys.modules["torch._thnn"] = mod
    torch._thnn = mod


def _inputs(rng, b, cg, c, h, w, density, neg=False, scale=10.0):
    g = rng.standard_normal((b, cg, h, w)).astype(np.float32)
    d = (rng.random((b, c, h, w)) * scale).astype(np.float32)
    if density is None:
        return g, d, None
    mask = rng.random((b, 1, h, w)) < density
    s = (mask * (rng.random((b, 1, h, w)) * scale + 0.1)).astype(np.float32)
    if neg:
        s = s * np.where(rng.random(s.shape) < 0.3, -1.0, 1.0).astype(np.float32)
    return g, d, s


def main():
    sys.path.insert(0, REF)
    _install_thnn_shim()
    from network.libs.post_process import CSPN_new, CSPN_ours  # the reference, unmodified

    torch.manual_seed(0)
    torch.set_num_threads(1)
    out = {}
    rng = np.random.default_rng(20261017)

    def run_a(name, g, d, s, t, grads=True):
        tg = torch.from_numpy(g).requires_grad_(grads)
        td = torch.from_numpy(d).requires_grad_(grads)
        ts = None if s is None else torch.from_numpy(s)
        y = CSPN_new.AffinityPropagate(t, 3)(tg, td, ts)
        out[name + "/guidance"], out[name + "/depth"] = g, d
        if s is not None:
            out[name + "/sparse"] = s
        out[name + "/iters"] = np.int64(t)
        out[name + "/out"] = y.detach().numpy()
        if grads and t > 0:
            go = rng.standard_normal(y.shape).astype(np.float32)
            y.backward(torch.from_numpy(go))
            out[name + "/grad_out"] = go
            out[name + "/grad_guidance"] = tg.grad.numpy()
            out[name + "/grad_depth"] = td.grad.numpy()

    def run_b(name, x, g, s, t, grads=True):
        tx = torch.from_numpy(x).requires_grad_(grads)
        tg = torch.from_numpy(g).requires_grad_(grads)
        ts = None if s is None else torch.from_numpy(s)
        y = CSPN_ours.AffinityPropagate(prop_time=t)(tx, tg, sparse_depth=ts)
        out[name + "/guidance"], out[name + "/depth"] = g, x
        if s is not None:
            out[name + "/sparse"] = s
        out[name + "/iters"] = np.int64(t)
        out[name + "/out"] = y.detach().numpy()
        if grads:
            go = rng.standard_normal(y.shape).astype(np.float32)
            y.backward(torch.from_numpy(go))
            out[name + "/grad_out"] = go
            out[name + "/grad_guidance"] = tg.grad.numpy()
            out[name + "/grad_depth"] = tx.grad.numpy()

    # ---- mode A (CSPN_new.AffinityPropagate(T, 3))
    g, d, s = _inputs(rng, 2, 8, 1, 12, 17, 0.05);            run_a("A_small_sparse_T24", g, d, s, 24)
    g, d, s = _inputs(rng, 1, 12, 1, 9, 11, 0.10, neg=True);  run_a("A_cg12_negsparse_T6", g, d, s, 6)
    g, d, s = _inputs(rng, 1, 8, 1, 5, 7, None);               run_a("A_nosparse_T3", g, d, s, 3)
    g, d, s = _inputs(rng, 1, 8, 3, 6, 8, 0.10);               run_a("A_multichan_T5", g, d, s, 5)
    g, d, s = _inputs(rng, 2, 8, 1, 2, 3, 0.20);               run_a("A_2x3_T2", g, d, s, 2)
    g, d, s = _inputs(rng, 1, 8, 1, 1, 1, None);               run_a("A_1x1_T2", g, d, s, 2, grads=False)
    g, d, s = _inputs(rng, 1, 8, 1, 1, 9, 0.2);                run_a("A_1x9_T4", g, d, s, 4)
    g, d, s = _inputs(rng, 1, 8, 1, 7, 1, 0.2);                run_a("A_7x1_T4", g, d, s, 4)
    g, d, s = _inputs(rng, 1, 8, 1, 8, 9, 0.05)
    g[0, :, 3:6, 3:6] = 0.0                                     # a pixel whose 8 gathered weights are all zero -> 0/0
    run_a("A_zero_guidance_nan_T4", g, d, s, 4, grads=False)
    g, d, s = _inputs(rng, 1, 8, 1, 40, 70, 0.0072);           run_a("A_40x70_T24", g, d, s, 24)
    g, d, s = _inputs(rng, 1, 8, 1, 33, 130, 0.05, scale=80.0); run_a("A_33x130_T24_kitti_scale", g, d, s, 24, grads=False)

    # NYU-size known answer: inputs regenerated from the seed by the test, only the output is stored.
    rn = np.random.default_rng(304228)
    g, d, s = _inputs(rn, 1, 8, 1, 228, 304, 500.0 / 69312.0)
    y = CSPN_new.AffinityPropagate(24, 3)(torch.from_numpy(g), torch.from_numpy(d), torch.from_numpy(s))
    out["A_nyu_seed304228_T24/out"] = y.numpy()

    # ---- mode B (CSPN_ours.AffinityPropagate(prop_time=T))
    g, x, s = _inputs(rng, 2, 8, 1, 10, 13, 0.05);              run_b("B_k3_sparse_T24", x, g, s, 24)
    g, x, s = _inputs(rng, 1, 24, 1, 10, 13, 0.05);             run_b("B_k5_sparse_T12", x, g, s, 12)
    g, x, s = _inputs(rng, 1, 8, 1, 6, 5, None);                run_b("B_k3_nosparse_T2", x, g, s, 2)
    g, x, s = _inputs(rng, 1, 8, 2, 7, 9, 0.1, neg=True);       run_b("B_k3_multichan_negsparse_T4", x, g, s, 4)
    g, x, s = _inputs(rng, 1, 24, 1, 3, 4, 0.2);                run_b("B_k5_3x4_T3", x, g, s, 3)
    g, x, s = _inputs(rng, 1, 8, 1, 36, 75, 0.0072);            run_b("B_k3_36x75_T24", x, g, s, 24)

    path = os.path.join(HERE, "cspn_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays; torch", torch.__version__)


if __name__ == "__main__":
    main()
