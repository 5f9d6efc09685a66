"""The loss / metrics kernels downstream of the module (csrc/cspn_loss.cu, cspn_monodepth_b200/criteria.py) against the
reference-run vectors (tests/golden/loss_golden.npz) and the numpy oracle (oracle/loss_oracle.py).

Tolerances: the kernels accumulate in double, the reference in fp32: loss / metrics agree to 2e-5 relative (fp32 mean of
~1e5 terms), the gradient to 1e-7 absolute (it is +-1/count); fp16 inputs are compared on the fp16-rounded values."""
import os

import numpy as np
import pytest
import torch

from cspn_monodepth_b200 import criteria
from oracle import loss_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


@pytest.fixture(scope="module")
def loss_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def _cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


def test_loss_and_metrics_match_reference_vectors(loss_golden):
    for name, c in loss_golden.items():
        p = _cu(c["pred"]).requires_grad_(True)
        t = _cu(c["target"])
        loss = criteria.MaskedL1Loss()(p, t)
        loss.backward()
        assert abs(loss.item() - float(c["loss"])) <= 2e-6 * max(1.0, abs(float(c["loss"]))), name
        assert np.abs(p.grad.cpu().numpy() - c["grad"]).max() <= 1e-7, name
        res = criteria.Result()
        res.evaluate(p.detach(), t)
        for k, ref in zip(criteria.METRIC_NAMES, c["metrics"]):
            assert abs(getattr(res, k) - ref) <= 2e-5 * max(1.0, abs(ref)), (name, k, getattr(res, k), ref)


@pytest.mark.parametrize("shape,density,dtype", [((8, 1, 228, 304), 1.0, torch.float32), ((32, 1, 352, 1216), 0.05, torch.float32),
                                                  ((32, 1, 352, 1216), 0.05, torch.float16), ((1, 1, 1, 1), 1.0, torch.float32),
                                                  ((3, 1, 17, 1031), 0.5, torch.float16)])
def test_full_size_vs_oracle_and_bitwise_reproducible(shape, density, dtype):
    rng = np.random.default_rng(7)
    target = ((rng.random(shape) * 9.5 + 0.5) * (rng.random(shape) < density)).astype(np.float32)
    target.flat[0] = 3.0                                                             # at least one valid pixel
    pred = (np.abs(target + rng.standard_normal(shape) * 0.3) + 0.05).astype(np.float32)
    p, t = _cu(pred, dtype).requires_grad_(True), _cu(target, dtype)
    ph, th = p.detach().float().cpu().numpy(), t.float().cpu().numpy()               # what the kernel sees (fp16-rounded)
    loss = criteria.MaskedL1Loss()(p, t)
    loss.backward(torch.tensor(2.5, device=DEV, dtype=dtype))
    ref_loss, ref_grad, n = loss_oracle.masked_l1(ph, th)
    tol = 2e-6 if dtype == torch.float32 else 1e-3
    assert abs(loss.item() - ref_loss) <= tol * max(1.0, abs(ref_loss))
    gtol = 1e-7 if dtype == torch.float32 else 1e-3 * 2.5 / n + 6e-8                 # fp16 gradient storage (subnormal spacing 6e-8)
    assert np.abs(p.grad.float().cpu().numpy() - 2.5 * ref_grad).max() <= gtol
    m1 = criteria.evaluate_device(p.detach(), t)
    m2 = criteria.evaluate_device(p.detach(), t)
    assert torch.equal(m1, m2)                                                       # deterministic reduction
    ref = loss_oracle.depth_metrics(ph, th)
    for k, v in zip(criteria.METRIC_NAMES + ("count",), m1.tolist()):
        assert abs(v - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, v, ref[k])
    l1 = criteria.MaskedL1Loss()(p.detach(), t)
    assert torch.equal(l1, loss.detach())


def test_no_valid_pixel_is_nan_like_the_reference_and_errors_are_runtime_errors():
    p = torch.ones(1, 1, 4, 4, device=DEV, requires_grad=True)
    t = torch.zeros(1, 1, 4, 4, device=DEV)
    loss = criteria.MaskedL1Loss()(p, t)
    assert torch.isnan(loss)                                                         # mean of an empty selection (criteria.py:38)
    with pytest.raises(RuntimeError):
        criteria.MaskedL1Loss()(torch.ones(2, 2), torch.ones(2, 2))                  # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        criteria.MaskedL1Loss()(torch.ones(2, 2, device=DEV), torch.ones(2, device=DEV))   # criteria.py:32


def test_loss_under_cuda_graph_replay():
    rng = np.random.default_rng(3)
    target = (rng.random((2, 1, 64, 96)) * 9 + 0.5).astype(np.float32)
    p, t = _cu(target + 0.25), _cu(target)
    criteria.evaluate_device(p, t)                                                   # allocate the scratch outside the capture
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        criteria.evaluate_device(p, t)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = criteria.evaluate_device(p, t)
        for shift in (0.25, 0.5):
            p.copy_(_cu(target + shift))
            g.replay()
            s.synchronize()
            assert abs(out[4].item() - shift) < 1e-5                                 # mae
