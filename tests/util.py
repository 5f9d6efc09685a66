"""Seeded synthetic inputs shared by the tests."""
import numpy as np

NYU_DENSITY = 500.0 / 69312.0


def make_inputs(seed, b, cg, c, h, w, density=NYU_DENSITY, neg=False, scale=10.0, sparse_channels=1):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((b, cg, h, w)).astype(np.float32)
    d = (rng.random((b, c, h, w)) * scale).astype(np.float32)
    if density is None:
        return g, d, None
    mask = rng.random((b, sparse_channels, h, w)) < density
    s = (mask * (rng.random((b, sparse_channels, h, w)) * scale + 0.1)).astype(np.float32)
    if neg:
        s = s * np.where(rng.random(s.shape) < 0.3, -1.0, 1.0).astype(np.float32)
    return g, d, s


def nyu_golden_inputs():
    """Inputs of the golden case A_nyu_seed304228_T24 (same draw order as tests/golden/make_golden.py::_inputs)."""
    rng = np.random.default_rng(304228)
    g = rng.standard_normal((1, 8, 228, 304)).astype(np.float32)
    d = (rng.random((1, 1, 228, 304)) * 10.0).astype(np.float32)
    mask = rng.random((1, 1, 228, 304)) < NYU_DENSITY
    s = (mask * (rng.random((1, 1, 228, 304)) * 10.0 + 0.1)).astype(np.float32)
    return g, d, s


def case_config(name, case):
    """(mode, ksize, iters) of a golden case."""
    mode = 0 if name.startswith("A_") else 1
    cg = case["guidance"].shape[1]
    ksize = 3 if mode == 0 else int(round((cg + 1) ** 0.5))
    return mode, ksize, int(case["iters"])


def assert_close_nan(actual, expected, atol, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} vs {expected.shape}"
    assert np.array_equal(np.isnan(actual), np.isnan(expected)), f"{what}: NaN pattern differs"
    ok = ~np.isnan(expected)
    if ok.any():
        err = np.abs(actual[ok] - expected[ok]).max()
        assert err <= atol, f"{what}: max-abs {err:.3e} > {atol:.1e}"
