"""Golden vectors for the loss / metrics kernels, produced by RUNNING THE REFERENCE's own classes (build container only):

    python tests/golden/make_loss_golden.py      ->  tests/golden/loss_golden.npz

`libs/criterion/criteria.py` imports a module that does not exist in the reference (`libs.image_processor`, SURVEY.md
appendix B #8); an empty stand-in is registered so that the file imports - MaskedL1Loss itself (:27-39) runs unmodified, with
autograd for the gradient.  `libs/metrics.py` imports as is; Result.evaluate (:49-83) runs unmodified on CPU tensors.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CSPN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    stub = types.ModuleType("libs.image_processor")
    stub.sobel_filter = None
    sys.modules["libs.image_processor"] = stub
    from libs.criterion.criteria import MaskedL1Loss
    from libs.metrics import Result
    out = {}
    rng = np.random.default_rng(2024)
    for name, shape, density in (("nyu", (1, 1, 114, 152), 1.0), ("kitti_sparse", (1, 1, 96, 352), 0.05), ("tiny", (1, 1, 5, 7), 0.5)):
        target = (rng.random(shape) * 9.5 + 0.5).astype(np.float32) * (rng.random(shape) < density)
        pred = (np.abs(target + rng.standard_normal(shape) * 0.3) + 0.05 + (target == 0) * rng.random(shape)).astype(np.float32)
        tp = torch.from_numpy(pred).requires_grad_(True)
        tt = torch.from_numpy(target.astype(np.float32))
        loss = MaskedL1Loss()(tp, tt)
        loss.backward()
        res = Result()
        res.evaluate(tp.detach(), tt)
        out[name + "/pred"], out[name + "/target"] = pred, target.astype(np.float32)
        out[name + "/loss"] = np.float32(loss.item())
        out[name + "/grad"] = tp.grad.numpy()
        out[name + "/metrics"] = np.array([res.irmse, res.imae, res.mse, res.rmse, res.mae, res.absrel, res.lg10, res.delta1, res.delta2, res.delta3], np.float64)
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
