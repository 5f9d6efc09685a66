"""The loss / metrics oracle (oracle/loss_oracle.py) against vectors produced by the reference's own MaskedL1Loss
(libs/criterion/criteria.py:27-39, with autograd) and Result.evaluate (libs/metrics.py:49-83) - tests/golden/make_loss_golden.py."""
import os

import numpy as np
import pytest

from oracle import loss_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("irmse", "imae", "mse", "rmse", "mae", "absrel", "lg10", "delta1", "delta2", "delta3")


@pytest.fixture(scope="module")
def loss_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def test_oracle_matches_reference_loss_and_metrics(loss_golden):
    assert set(loss_golden) == {"nyu", "kitti_sparse", "tiny"}
    for name, c in loss_golden.items():
        loss, grad, n = loss_oracle.masked_l1(c["pred"], c["target"])
        assert n == int((c["target"] > 0).sum()) and n > 0
        assert abs(loss - float(c["loss"])) <= 2e-6 * max(1.0, abs(loss)), name
        assert np.abs(grad - c["grad"]).max() <= 1e-7 * max(1.0, 1.0 / n) + 1e-9, name
        m = loss_oracle.depth_metrics(c["pred"], c["target"])
        for k, ref in zip(NAMES, c["metrics"]):
            assert abs(m[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (name, k, m[k], ref)


def test_oracle_edge_cases():
    loss, grad, n = loss_oracle.masked_l1(np.ones((1, 1, 2, 2)), np.zeros((1, 1, 2, 2)))
    assert n == 0 and np.isnan(loss) and not grad.any()                          # mean of an empty selection (criteria.py:38)
    loss, grad, n = loss_oracle.masked_l1(np.array([2.0, 3.0, 1.0]), np.array([2.0, 1.0, 0.0]))
    assert n == 2 and loss == 1.0 and list(grad) == [0.0, 0.5, 0.0]               # sign(0) = 0; invalid pixel gets no gradient
