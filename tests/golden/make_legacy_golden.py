"""Golden vectors for the legacy max-of-8 CSPN, produced by RUNNING THE REFERENCE's own classes
(build container only):

    python tests/golden/make_legacy_golden.py      ->  tests/golden/legacy_golden.npz

`network/libs/post_process/CSPN.py:14-56` (AffinityPropagate: 16 steps, re-injects the SPARSE SAMPLES) and `:126-164`
(AffinityPropagate_prediction: no sparse input) build their 3x3 box filters with `torch.ones(...).cuda()` (`:87,:96`), so the
file only runs where CUDA is present.  There is no GPU here: `Tensor.cuda` is patched to the identity for the duration of the
run - everything else of the module executes unmodified on CPU fp32, with autograd for the gradients.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CSPN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from network.libs.post_process import CSPN
    torch.manual_seed(0)
    out = {}
    rng = np.random.default_rng(1618)
    cases = (("nyu_crop", 2, 8, 57, 76, 0.02, False), ("twelve_channels", 1, 12, 33, 47, 0.05, False), ("neg_sparse", 1, 8, 20, 31, 0.2, True),
             ("tiny", 1, 8, 3, 5, 0.3, False), ("one_pixel", 1, 8, 1, 1, 0.0, False), ("row", 1, 8, 1, 9, 0.3, False),
             ("zero_gate", 1, 8, 9, 11, 0.1, False))
    for name, b, cg, h, w, density, neg in cases:
        g = rng.standard_normal((b, cg, h, w)).astype(np.float32)
        if name == "zero_gate":
            g[:, 3, 2:7, 3:9] = 0.0                                       # a 3x3 window of zeros in one gate: 0/0 = NaN spreads through max
        d = (rng.random((b, 1, h, w)) * 10).astype(np.float32)
        s = ((rng.random((b, 1, h, w)) < density) * (rng.random((b, 1, h, w)) * 10 + 0.1)).astype(np.float32)
        if neg:
            s *= np.where(rng.random(s.shape) < 0.3, -1.0, 1.0).astype(np.float32)
        tg, td, ts = (torch.from_numpy(a.copy()) for a in (g, d, s))
        tg.requires_grad_(True); td.requires_grad_(True); ts.requires_grad_(True)
        y = CSPN.AffinityPropagate()(tg, td, ts)
        go = rng.standard_normal(d.shape).astype(np.float32)
        finite = torch.isfinite(y).all().item()
        if finite:
            y.backward(torch.from_numpy(go))
        y2 = CSPN.AffinityPropagate_prediction()(torch.from_numpy(g), torch.from_numpy(d))
        for k, v in (("guidance", g), ("depth", d), ("sparse", s), ("out", y.detach().numpy()), ("out_prediction", y2.numpy()), ("grad_out", go)):
            out[f"{name}/{k}"] = v
        if finite:
            out[f"{name}/grad_guidance"] = tg.grad.numpy()
            out[f"{name}/grad_depth"] = td.grad.numpy()
            out[f"{name}/grad_sparse"] = ts.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "legacy_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
