"""Host-side mirror of the reference interface (no GPU): signatures, error behaviour, drop-in hooks."""
import inspect
import sys

import pytest
import torch

import cspn_monodepth_b200 as pkg
from cspn_monodepth_b200 import cspn_new, cspn_ours, dropin, sharding
from cspn_monodepth_b200.functional import kernel_size_from_channels


def test_mode_a_signature_matches_reference():
    # CSPN_new.py:19 __init__(self, prop_time, prop_kernel); :26 forward(self, guidance, blur_depth, sparse_depth=None)
    assert list(inspect.signature(cspn_new.AffinityPropagate.__init__).parameters) == ["self", "prop_time", "prop_kernel"]
    sig = inspect.signature(cspn_new.AffinityPropagate.forward)
    assert list(sig.parameters) == ["self", "guidance", "blur_depth", "sparse_depth"]
    assert sig.parameters["sparse_depth"].default is None
    m = cspn_new.AffinityPropagate(24, 3)
    assert (m.prop_time, m.prop_kernel, m.in_feature, m.out_feature) == (24, 3, 1, 1)
    assert len(m.state_dict()) == 0 and len(list(m.parameters())) == 0


def test_mode_b_signature_matches_reference():
    # CSPN_ours.py:20 __init__(self, prop_time); :24 forward(self, x, guided, sparse_depth=None)
    assert list(inspect.signature(cspn_ours.AffinityPropagate.__init__).parameters) == ["self", "prop_time"]
    assert list(inspect.signature(cspn_ours.AffinityPropagate.forward).parameters) == ["self", "x", "guided", "sparse_depth"]
    m = cspn_ours.AffinityPropagate(prop_time=24)
    assert m.times == 24 and len(m.state_dict()) == 0


def test_zero_iterations_return_the_input_object():
    d = torch.rand(1, 1, 4, 5)
    assert cspn_new.AffinityPropagate(0, 3)(torch.randn(1, 8, 4, 5), d) is d
    assert cspn_ours.AffinityPropagate(0)(d, torch.randn(1, 8, 4, 5)) is d


def test_cpu_tensors_fail_loudly_no_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cspn_new.AffinityPropagate(2, 3)(torch.randn(1, 8, 4, 5), torch.rand(1, 1, 4, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cspn_ours.AffinityPropagate(2)(torch.rand(1, 1, 4, 5), torch.randn(1, 8, 4, 5))


def test_reference_error_cases():
    with pytest.raises(RuntimeError, match="prop_kernel=3"):       # CSPN_new.py:122 breaks for 5
        cspn_new.AffinityPropagate(2, 5)(torch.randn(1, 8, 4, 5), torch.rand(1, 1, 4, 5))
    with pytest.raises(RuntimeError, match="K\\*K-1"):              # CSPN_ours.py:41 reshape fails
        cspn_ours.AffinityPropagate(2)(torch.rand(1, 1, 4, 5), torch.randn(1, 7, 4, 5))
    assert kernel_size_from_channels(8) == 3 and kernel_size_from_channels(24) == 5 and kernel_size_from_channels(48) == 7


def test_dropin_install_registers_reference_module_names():
    saved = {k: sys.modules.get(k) for k in ("network.libs.post_process.CSPN_new", "network.libs.post_process.CSPN_ours")}
    try:
        dropin.install()
        from network.libs.post_process.CSPN_new import AffinityPropagate as A   # unet_cspn_nyu.py:9
        from network.libs.post_process.CSPN_ours import AffinityPropagate as B_  # unet_ours.py:16
        assert A is cspn_new.AffinityPropagate and B_ is cspn_ours.AffinityPropagate
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_patch_model_swaps_post_process_layer():
    class AffinityPropagate(torch.nn.Module):      # stands in for the reference class (same name / attributes)
        def __init__(self):
            super().__init__()
            self.prop_time, self.prop_kernel = 24, 3

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(1, 1, 1)
            self.post_process_layer = AffinityPropagate()

    net = Net()
    assert dropin.patch_model(net) == 1
    assert isinstance(net.post_process_layer, cspn_new.AffinityPropagate)
    assert dropin.patch_model(net) == 0


def test_batch_slices_cover_batch_exactly_once():
    for gb in (1, 7, 8, 32, 256):
        for ws in (1, 2, 4, 8):
            seen = []
            for r in range(ws):
                s = sharding.batch_slice(gb, ws, r)
                seen += list(range(gb))[s]
            assert seen == list(range(gb))
    with pytest.raises(ValueError):
        sharding.batch_slice(8, 2, 2)
    assert pkg.MODE_NEW == 0 and pkg.MODE_OURS == 1
