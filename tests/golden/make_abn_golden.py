"""Golden vectors for the in-place ABN kernels (build container only):

    python tests/golden/make_abn_golden.py      ->  tests/golden/abn_golden.npz

The reference's InPlaceABN needs its own native extension (network/libs/inplace_abn/build.py: cffi over THC, torch 0.4), which
cannot be built against this image's torch, so it cannot be run.  What CAN be run is the class the reference swaps it with on one
GPU (unet_cspn_nyu.py:25-29; bn.py:23-44 `ABN` = nn.BatchNorm2d + activation) - torch's own batch norm with autograd.  The two
agree exactly when BatchNorm2d's weight is |w| + eps (bn.cu:153 `gamma = abs(weight) + eps`), so that is what is recorded here:
forward output, running statistics, and the autograd gradients w.r.t. x, the BatchNorm weight (times sign(w) = d gamma / d w) and bias.
"""
import os

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


def run(x, w, b, training, momentum, eps, activation, slope, rm, rv, dz):
    c = x.shape[1]
    bn = nn.BatchNorm2d(c, eps=eps, momentum=momentum, affine=True).double()
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(np.abs(w).astype(np.float64) + eps) if w is not None else torch.ones(c, dtype=torch.float64))
        bn.bias.copy_(torch.from_numpy(b.astype(np.float64)) if b is not None else torch.zeros(c, dtype=torch.float64))
        bn.running_mean.copy_(torch.from_numpy(rm.astype(np.float64)))
        bn.running_var.copy_(torch.from_numpy(rv.astype(np.float64)))
    bn.train(training)
    act = {"leaky_relu": nn.LeakyReLU(slope), "elu": nn.ELU(), "none": nn.Identity()}[activation]
    tx = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    z = act(bn(tx))
    z.backward(torch.from_numpy(dz.astype(np.float64)))
    sign = np.sign(w) if w is not None else None
    return {"z": z.detach().numpy(), "running_mean": bn.running_mean.numpy().copy(), "running_var": bn.running_var.numpy().copy(),
            "dx": tx.grad.numpy(), "dweight": bn.weight.grad.numpy() * sign if w is not None else np.zeros(0),
            "dbias": bn.bias.grad.numpy() if b is not None else np.zeros(0)}


def main():
    rng = np.random.default_rng(77)
    out = {}
    cases = [("leaky_train", (2, 5, 6, 7), True, "leaky_relu", 0.01, True),
             ("elu_train", (3, 4, 5, 5), True, "elu", 0.01, True),
             ("none_train", (2, 3, 4, 9), True, "none", 0.01, True),
             ("leaky_eval", (2, 5, 6, 7), False, "leaky_relu", 0.2, True),
             ("leaky_noaffine", (2, 4, 3, 5), True, "leaky_relu", 0.01, False),
             ("odd_s", (1, 2, 3, 3), True, "leaky_relu", 0.01, True)]
    for name, shape, training, activation, slope, affine in cases:
        c = shape[1]
        x = (rng.standard_normal(shape) * 1.7 + rng.standard_normal((1, c, 1, 1))).astype(np.float32)
        w = (rng.standard_normal(c) + 0.3).astype(np.float32) if affine else None             # both signs: |w| and sign(w) matter
        b = rng.standard_normal(c).astype(np.float32) if affine else None
        rm = rng.standard_normal(c).astype(np.float32) * 0.1
        rv = (rng.random(c) + 0.5).astype(np.float32)
        dz = rng.standard_normal(shape).astype(np.float32)
        res = run(x, w, b, training, 0.1, 1e-5, activation, slope, rm, rv, dz)
        out[name + "/x"], out[name + "/dz"], out[name + "/rm0"], out[name + "/rv0"] = x, dz, rm, rv
        if affine:
            out[name + "/w"], out[name + "/b"] = w, b
        out[name + "/cfg"] = np.array([float(training), {"none": 0, "leaky_relu": 1, "elu": 2}[activation], slope, float(affine)], np.float64)
        for k, v in res.items():
            out[name + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "abn_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
