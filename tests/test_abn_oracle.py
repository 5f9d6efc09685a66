"""The in-place ABN oracle (oracle/abn_oracle.py) against vectors produced by running nn.BatchNorm2d + activation with autograd -
the class the reference itself swaps InPlaceABN with on one GPU (tests/golden/make_abn_golden.py explains why the native
extension cannot be run) - and its own sync-variant algebra (functions.py:185-200: mean of means, mean of var + (mean - means)^2)."""
import os

import numpy as np
import pytest

from oracle import abn_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACTS = {0: "none", 1: "leaky_relu", 2: "elu"}


def load_abn_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "abn_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def case_args(c):
    training, act, slope, affine = c["cfg"]
    return bool(training), ACTS[int(act)], float(slope), (c["w"] if affine else None), (c["b"] if affine else None)


def test_oracle_matches_batchnorm_autograd_vectors():
    cases = load_abn_golden()
    assert set(cases) == {"leaky_train", "elu_train", "none_train", "leaky_eval", "leaky_noaffine", "odd_s"}
    for name, c in cases.items():
        training, act, slope, w, b = case_args(c)
        z, mean, var, rm, rv = abn_oracle.abn_forward(c["x"], w, b, c["rm0"], c["rv0"], training, 0.1, 1e-5, act, slope)
        assert np.abs(z - c["z"]).max() <= 1e-12 * max(1.0, np.abs(c["z"]).max()), name
        assert np.abs(rm - c["running_mean"]).max() <= 1e-12 and np.abs(rv - c["running_var"]).max() <= 1e-12, name
        dx, dw, db = abn_oracle.abn_backward(z, c["dz"], var, w, b, training, 1e-5, act, slope)
        assert np.abs(dx - c["dx"]).max() <= 1e-10 * max(1.0, np.abs(c["dx"]).max()), name
        if w is not None and training:            # in eval mode the reference zeroes the parameter gradients (functions.py:147-150)
            assert np.abs(dw - c["dweight"]).max() <= 1e-9 * max(1.0, np.abs(c["dweight"]).max()), name
            assert np.abs(db - c["dbias"]).max() <= 1e-9 * max(1.0, np.abs(c["dbias"]).max()), name
        if not training and w is not None:
            assert not dw.any() and not db.any()


def test_sync_statistics_equal_the_references_mean_of_means_rule():
    rng = np.random.default_rng(3)
    xs = [rng.standard_normal((2, 3, 4, 5)) + k for k in range(3)]                   # equal per-replica counts, as under DataParallel
    z, mean, var, _, _ = abn_oracle.abn_forward(xs[1], None, None, np.zeros(3), np.ones(3), world_x=xs)
    means = np.stack([x.mean((0, 2, 3)) for x in xs])
    vars_ = np.stack([x.var((0, 2, 3)) for x in xs])
    ref_mean = means.mean(0)                                                       # functions.py:196
    ref_var = (vars_ + (ref_mean - means) ** 2).mean(0)                            # :197
    assert np.allclose(mean, ref_mean, atol=1e-12) and np.allclose(var, ref_var, atol=1e-12)
    whole = np.concatenate(xs)
    zw = abn_oracle.abn_forward(whole, None, None, np.zeros(3), np.ones(3))[0]
    assert np.allclose(z, zw[2:4], atol=1e-12)                                     # replica 1's slice of the full-batch result


def test_activation_undo_inverts_the_activation():
    z = np.linspace(-3, 3, 41)
    for act in ("leaky_relu", "elu", "none"):
        a = abn_oracle._act(z, act, 0.01)
        back, g = abn_oracle._act_undo(a, np.ones_like(a), act, 0.01)
        assert np.allclose(back, z, atol=1e-12)
        eps = 1e-6
        num = (abn_oracle._act(z + eps, act, 0.01) - abn_oracle._act(z - eps, act, 0.01)) / (2 * eps)
        ok = np.abs(z) > 1e-3
        assert np.allclose(g[ok], num[ok], atol=1e-6)
