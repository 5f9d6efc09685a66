"""The output-heads oracle (oracle/heads_oracle.py) against vectors produced by the reference's own
Simple_Gudi_UpConv_Block_Last_Layer classes with autograd (tests/golden/make_heads_golden.py)."""
import os

import numpy as np
import pytest

from oracle import heads_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def heads_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "heads_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


def test_oracle_matches_the_reference_heads(heads_golden):
    assert len(heads_golden) == 5
    for name, c in heads_golden.items():
        H, W = c["depth"].shape[2:]
        for out, wk, gk, gwk in (("depth", "w_depth", "grad_depth", "grad_w_depth"), ("guidance", "w_guid", "grad_guidance", "grad_w_guid")):
            y = heads_oracle.forward(c["x"], c[wk], H, W)
            assert np.abs(y - c[out]).max() <= 2e-6 * max(1.0, np.abs(c[out]).max()), (name, out)
        gx1, gw1 = heads_oracle.backward(c["x"], c["w_depth"], c["grad_depth"], H, W)
        gx2, gw2 = heads_oracle.backward(c["x"], c["w_guid"], c["grad_guidance"], H, W)
        assert np.abs(gx1 - c["grad_x_depth_only"]).max() <= 2e-6 * max(1.0, np.abs(c["grad_x_depth_only"]).max()), name
        assert np.abs(gx1 + gx2 - c["grad_x"]).max() <= 2e-6 * max(1.0, np.abs(c["grad_x"]).max()), name
        for mine, ref in ((gw1, c["grad_w_depth"]), (gw2, c["grad_w_guid"])):
            assert np.abs(mine - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), name


def test_zero_insertion_structure():
    """Even output positions see only the centre tap; cells cut off by an odd crop get no gradient."""
    x = np.ones((1, 1, 2, 3))
    w = np.arange(9, dtype=np.float64).reshape(1, 1, 3, 3)
    y = heads_oracle.forward(x, w, 3, 5)                      # crop of the 4 x 6 unpooled tensor
    assert y[0, 0, 0, 0] == 4 and y[0, 0, 0, 1] == 3 + 5 and y[0, 0, 1, 0] == 1 + 7 and y[0, 0, 1, 1] == 0 + 2 + 6 + 8
    gx, _ = heads_oracle.backward(np.ones((1, 1, 3, 3)), w, np.ones((1, 1, 3, 5)), 3, 5)
    assert gx[0, 0, 2].tolist() == [0, 0, 0]                  # row 2 of x would land on output row 4 >= 3: cropped away
