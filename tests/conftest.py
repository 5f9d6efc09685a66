import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by running the reference itself (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cspn_golden.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def unet_golden():
    """What the reference's own UNets hand to the CSPN module, and the reference module's results on it
    (tests/golden/make_unet_golden.py): realistic value distribution (small, smooth guidance; 12 channels in mode A)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cspn_unet_heads.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    return cases
