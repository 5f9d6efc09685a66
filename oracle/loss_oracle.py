"""CPU restatement (numpy, float64 accumulation) of the loss / metrics downstream of the CSPN module.  TEST INFRASTRUCTURE
ONLY: imported by tests/, bench.py's checks - never by the product.

  masked_l1(pred, target)        libs/criterion/criteria.py:27-39  mean |target - pred| over target > 0, and its gradient
  depth_metrics(output, target)  libs/metrics.py:49-83             Result.evaluate
Pinned by tests/golden/loss_golden.npz, produced by running the reference's own classes (tests/golden/make_loss_golden.py).
"""
import numpy as np


def masked_l1(pred, target):
    pred, target = np.asarray(pred, np.float64), np.asarray(target, np.float64)
    valid = target > 0                                                     # criteria.py:33
    n = valid.sum()
    diff = (target - pred)[valid]                                          # :36-37
    loss = np.abs(diff).mean() if n else np.nan                            # :38
    grad = np.where(valid, -np.sign(target - pred) / max(n, 1), 0.0)       # d mean|t - p| / dp
    return loss, grad, int(n)


def depth_metrics(output, target):
    output, target = np.asarray(output, np.float64), np.asarray(target, np.float64)
    valid = target > 0                                                     # metrics.py:57
    o, t = output[valid], target[valid]
    d = np.abs(o - t)                                                      # :61
    ratio = np.maximum(o / t, t / o)                                       # :68
    inv = np.abs(1 / o - 1 / t)                                            # :76-78
    mse = (d ** 2).mean()
    return {"irmse": np.sqrt((inv ** 2).mean()), "imae": inv.mean(), "mse": mse, "rmse": np.sqrt(mse), "mae": d.mean(),
            "absrel": (d / t).mean(), "lg10": np.abs(np.log10(o) - np.log10(t)).mean(),
            "delta1": (ratio < 1.25).mean(), "delta2": (ratio < 1.25 ** 2).mean(), "delta3": (ratio < 1.25 ** 3).mean(),
            "count": float(valid.sum())}
