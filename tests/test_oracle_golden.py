"""The oracle (numpy and C restatements) against vectors produced by the reference itself,
plus the oracle-free invariants of SURVEY.md appendix A.4.  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle, cspn_oracle
from tests.util import assert_close_nan, case_config, make_inputs, nyu_golden_inputs

FWD_ATOL = 1e-4      # north_star tolerance; observed ~3e-6 at depth scale 10 (two valid fp32 summation orders)
GRAD_RTOL = 1e-4     # relative to the largest gradient entry


def _names(golden):
    return sorted(n for n in golden if "guidance" in golden[n])


def test_golden_file_has_all_cases(golden):
    assert len(_names(golden)) >= 17
    assert "A_nyu_seed304228_T24" in golden


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_forward_matches_reference(golden, impl):
    for name in _names(golden):
        case = golden[name]
        mode, ksize, iters = case_config(name, case)
        g, d, s = case["guidance"], case["depth"], case.get("sparse")
        if impl == "numpy":
            y = cspn_oracle.mode_a_forward(g, d, s, iters) if mode == 0 else cspn_oracle.mode_b_forward(d, g, s, iters)
        else:
            y = c_oracle.forward(g, d, s, iters, ksize, mode)
        scale = max(1.0, float(np.nanmax(np.abs(d))) / 10.0)
        assert_close_nan(y, case["out"], FWD_ATOL * scale, f"{impl}:{name}")


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_backward_matches_reference_autograd(golden, impl):
    checked = 0
    for name in _names(golden):
        case = golden[name]
        if "grad_out" not in case:
            continue
        mode, ksize, iters = case_config(name, case)
        g, d, s, go = case["guidance"], case["depth"], case.get("sparse"), case["grad_out"]
        if impl == "numpy":
            if mode == 0:
                gg, gd = cspn_oracle.mode_a_backward(g, d, s, go, iters)
            else:
                gd, gg = cspn_oracle.mode_b_backward(d, g, s, go, iters)
        else:
            gg, gd = c_oracle.backward(g, d, s, go, iters, ksize, mode)
        for got, key in ((gg, "grad_guidance"), (gd, "grad_depth")):
            ref = case[key]
            assert_close_nan(got, ref, GRAD_RTOL * max(1.0, np.abs(ref).max()), f"{impl}:{name}:{key}")
        if mode == 0 and g.shape[1] > 8:
            assert np.all(gg[:, 8:] == 0)       # channels the forward never reads get exact zeros
        checked += 1
    assert checked >= 12


def test_nyu_size_known_answer(golden):
    g, d, s = nyu_golden_inputs()
    y = c_oracle.forward(g, d, s, 24, 3, 0)
    assert_close_nan(y, golden["A_nyu_seed304228_T24"]["out"], FWD_ATOL, "nyu")


def test_numpy_and_c_oracles_agree():
    g, d, s = make_inputs(7, 2, 8, 1, 31, 45, density=0.03)
    assert_close_nan(c_oracle.forward(g, d, s, 24, 3, 0), cspn_oracle.mode_a_forward(g, d, s, 24), 2e-5)
    g, d, s = make_inputs(8, 1, 24, 1, 19, 23, density=0.03)
    assert_close_nan(c_oracle.forward(g, d, s, 12, 5, 1), cspn_oracle.mode_b_forward(d, g, s, 12), 2e-5)


# ---- invariants --------------------------------------------------
def test_invariant_convex_hull_mode_a():
    g, d, s = make_inputs(1, 2, 8, 1, 20, 30, density=0.05)
    y = c_oracle.forward(g, d, s, 24, 3, 0)
    assert y.min() >= d.min() - 1e-4 and y.max() <= d.max() + 1e-4


def test_invariant_constant_depth_is_fixed_point():
    g, _, s = make_inputs(2, 1, 8, 1, 17, 19, density=0.05)
    d = np.full((1, 1, 17, 19), 3.25, np.float32)
    assert np.abs(c_oracle.forward(g, d, s, 24, 3, 0) - 3.25).max() < 1e-5
    yb = c_oracle.forward(g, d, s, 24, 3, 1)
    assert np.abs(yb[..., 12, 9] - 3.25).max() > -1        # borders of mode B lose mass ...
    assert yb[0, 0, 0, 0] < 3.25                             # ... (zero padding, no renormalisation)


def test_invariant_guidance_scale_and_softmax_shift():
    g, d, s = make_inputs(3, 1, 8, 1, 15, 21, density=0.05)
    a = c_oracle.forward(g, d, s, 24, 3, 0)
    b = c_oracle.forward(g * np.float32(-3.7), d, s, 24, 3, 0)
    assert np.abs(a - b).max() < 1e-4
    a = c_oracle.forward(g, d, s, 24, 3, 1)
    b = c_oracle.forward(g + np.float32(1.5), d, s, 24, 3, 1)
    assert np.abs(a - b).max() < 1e-4


def test_invariant_batch_slices_and_extra_channels():
    g, d, s = make_inputs(4, 3, 12, 1, 14, 18, density=0.05)
    full = c_oracle.forward(g, d, s, 12, 3, 0)
    parts = np.concatenate([c_oracle.forward(g[i:i + 1], d[i:i + 1], s[i:i + 1], 12, 3, 0) for i in range(3)])
    assert np.array_equal(full, parts)
    g2 = g.copy(); g2[:, 8:] = 99.0
    assert np.array_equal(full, c_oracle.forward(g2, d, s, 12, 3, 0))


def test_invariant_gradient_mass_mode_a():
    g, d, _ = make_inputs(5, 2, 8, 1, 10, 16, density=None)
    go = np.ones_like(d)
    _, gd = c_oracle.backward(g, d, None, go, 8, 3, 0)
    assert abs(gd.sum() - d.size) < 1e-2 * d.size ** 0.5


def test_masked_pixels_keep_blur_depth_not_sparse_value():
    g, d, s = make_inputs(6, 1, 8, 1, 12, 12, density=0.2)
    y = c_oracle.forward(g, d, s, 5, 3, 0)
    hit = s > 0
    assert hit.any() and np.array_equal(y[hit], d[hit])       # CSPN_new.py:73,90 re-injects the BLUR depth


def test_torch_port_matches_reference(golden):
    """The op-for-op PyTorch port that bench.py times as the reference arm."""
    import torch

    from oracle import torch_port
    for name in _names(golden):
        case = golden[name]
        mode, _, iters = case_config(name, case)
        g, d = torch.from_numpy(case["guidance"]), torch.from_numpy(case["depth"])
        s = torch.from_numpy(case["sparse"]) if "sparse" in case else None
        y = torch_port.mode_a_forward(g, d, s, iters) if mode == 0 else torch_port.mode_b_forward(d, g, s, iters)
        assert_close_nan(y.numpy(), case["out"], 1e-5 * max(1.0, float(np.nanmax(np.abs(case["depth"]))) / 10.0), name)


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_unet_head_vectors(unet_golden, impl):
    """Inputs with the distribution the reference UNets really produce (|guidance| ~ 1e-2, depth ~ 1e-2 at random
    initialisation): the tolerance is relative to the value range (north_star: 1e-4 at range 10 = 1e-5 relative)."""
    for name, case in sorted(unet_golden.items()):
        mode, ksize, iters = case_config(name, case)
        g, d, s, go = case["guidance"], case["depth"], case["sparse"], case["grad_out"]
        if impl == "numpy":
            y = cspn_oracle.mode_a_forward(g, d, s, iters) if mode == 0 else cspn_oracle.mode_b_forward(d, g, s, iters)
            if mode == 0:
                gg, gd = cspn_oracle.mode_a_backward(g, d, s, go, iters)
            else:
                gd, gg = cspn_oracle.mode_b_backward(d, g, s, go, iters)
        else:
            y = c_oracle.forward(g, d, s, iters, ksize, mode)
            gg, gd = c_oracle.backward(g, d, s, go, iters, ksize, mode)
        assert_close_nan(y, case["out"], 1e-5 * np.abs(case["out"]).max(), f"{impl}:{name}:out")
        assert_close_nan(gd, case["grad_depth"], GRAD_RTOL * np.abs(case["grad_depth"]).max(), f"{impl}:{name}:grad_depth")
        assert_close_nan(gg, case["grad_guidance"], GRAD_RTOL * np.abs(case["grad_guidance"]).max(), f"{impl}:{name}:grad_guidance")
