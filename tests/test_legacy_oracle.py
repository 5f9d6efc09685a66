"""The legacy max-of-8 CSPN oracle (oracle/legacy_oracle.py) against vectors produced by the reference's own classes
(network/libs/post_process/CSPN.py, tests/golden/make_legacy_golden.py)."""
import os

import numpy as np
import pytest

from oracle import legacy_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def legacy_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "legacy_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def test_forward_matches_the_reference(legacy_golden):
    assert len(legacy_golden) == 7
    for name, c in legacy_golden.items():
        for sparse, ref in ((c["sparse"], c["out"]), (None, c["out_prediction"])):
            y = legacy_oracle.forward(c["guidance"], c["depth"], sparse)
            assert np.array_equal(np.isnan(y), np.isnan(ref)), name
            ok = ~np.isnan(ref)
            if ok.any():
                assert np.abs(y[ok] - ref[ok]).max() <= 1e-5 * max(1.0, np.abs(ref[ok]).max()), name
    assert np.isnan(legacy_golden["zero_gate"]["out"]).all()                      # 0/0 spreads through max in 16 steps
    m = legacy_golden["nyu_crop"]["sparse"] > 0
    assert np.array_equal(legacy_golden["nyu_crop"]["out"][m], legacy_golden["nyu_crop"]["sparse"][m])   # the SAMPLE is re-injected


def test_backward_matches_reference_autograd(legacy_golden):
    for name, c in legacy_golden.items():
        if "grad_guidance" not in c or name == "one_pixel":                       # one pixel: all 8 candidates tie up to rounding
            continue
        gg, gd, gs = legacy_oracle.backward(c["guidance"], c["depth"], c["sparse"], c["grad_out"])
        for mine, ref in ((gg, c["grad_guidance"]), (gd, c["grad_depth"]), (gs, c["grad_sparse"])):
            assert np.abs(mine - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), name
