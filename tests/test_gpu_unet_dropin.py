"""The reference's own UNets with the B200 module dropped in (SURVEY.md 8b / VERDICT r1 item 7): builds the unmodified
`network/unet_cspn_nyu.py` and `network/unet_ours.py` from the staged reference (baseline/_ref, see baseline/fetch_ref.py -
/root/reference itself does not exist on the GPU box), swaps `post_process_layer` with `dropin.patch_model`, and compares the
whole network's output and one training step's gradients with the un-patched model on the same GPU and the same weights.

Tolerance: the two runs share every cuDNN kernel up to the CSPN module; the module's own deviation is <= 1e-4 at depth
scale 10.  Output: 1e-4 of the output range.  Gradients: the backward of the ResNet-50 encoder (batch norm in training
mode) amplifies rounding noise, so they are compared per parameter tensor relative to that tensor's largest entry (1e-2, or
20x the reference's own run-to-run spread) and as a cosine similarity over all parameters (> 0.9999).
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from cspn_monodepth_b200 import _lib, cspn_new, cspn_ours, dropin

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DEV = "cuda:0"


def _import_reference():
    if not os.path.isdir(os.path.join(REF, "network")):
        pytest.skip("baseline/_ref is not staged (run `python baseline/fetch_ref.py` where /root/reference is mounted)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import _install_thnn_shim              # pac.py:20 imports the removed torch._thnn
    _install_thnn_shim()
    from network import unet_cspn_nyu, unet_ours
    return unet_cspn_nyu, unet_ours


def _rgbd(batch, seed):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(batch, 3, 228, 304, generator=g)
    dense = torch.rand(batch, 1, 228, 304, generator=g) * 9 + 0.5
    mask = torch.rand(batch, 1, 228, 304, generator=g) < 500.0 / 69312.0
    target = torch.rand(batch, 1, 228, 304, generator=g) * 10
    return torch.cat([rgb, dense * mask], dim=1).to(DEV), target.to(DEV)      # the reference's rgbd input: channel 3 = sparse depth


@pytest.mark.parametrize("heads", [False, True])
@pytest.mark.parametrize("which", ["unet_cspn_nyu", "unet_ours"])
def test_reference_unet_with_b200_module(which, heads):
    """heads=True: the two output heads upstream of the module (unet_cspn_nyu.py:331-332,383-384) are replaced as well - one
    fused launch of csrc/cspn_heads.cu for both; the reference's own heads then run as exact fp32 convolutions (TF32 off) so
    that the comparison is against the reference's arithmetic, not cuDNN's TF32 shortcut."""
    unet_cspn_nyu, unet_ours = _import_reference()
    torch.manual_seed(7)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = not heads
    ref_model = (unet_cspn_nyu if which == "unet_cspn_nyu" else unet_ours).resnet50(pretrained=False).to(DEV).train()
    our_model = copy.deepcopy(ref_model)
    assert dropin.patch_model(our_model, heads=heads) == 1
    if heads:
        from cspn_monodepth_b200.heads import Simple_Gudi_UpConv_Block_Last_Layer as OurHead
        assert isinstance(our_model.gud_up_proj_layer5, OurHead) and isinstance(our_model.gud_up_proj_layer6, OurHead)
    assert isinstance(our_model.post_process_layer, (cspn_new.AffinityPropagate, cspn_ours.AffinityPropagate))
    assert type(ref_model.post_process_layer).__module__.startswith("network.libs.post_process")
    assert our_model.state_dict().keys() == ref_model.state_dict().keys()       # the module has no parameters or buffers
    x, target = _rgbd(2, 3)

    def run(model):
        model.zero_grad(set_to_none=True)
        y = model(x)
        if isinstance(y, (list, tuple)):                                        # unet_ours.py:335 returns [depth, guidance]
            y = y[0]
        valid = target > 0
        loss = (y - target)[valid].abs().mean()                                 # MaskedL1Loss (libs/criterion/criteria.py:27-39)
        loss.backward()
        return y.detach().float(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}

    def compare(a, b):
        """(largest per-tensor deviation relative to the tensor's largest entry, cosine over all parameters)"""
        worst = max(float((a[n] - b[n]).abs().max()) / max(float(a[n].abs().max()), 1e-12) for n in a)
        dot = sum(float((a[n] * b[n]).sum()) for n in a)
        na = sum(float((a[n] ** 2).sum()) for n in a) ** 0.5
        nb = sum(float((b[n] ** 2).sum()) for n in a) ** 0.5
        return worst, dot / (na * nb)

    out_ref, g_ref = run(ref_model)
    out_ref2, g_ref2 = run(ref_model)                # the reference against itself: cuDNN / atomics noise floor of this network
    out_our, g_our = run(our_model)
    assert _lib.load().cspn_last_path() == _lib.PATH_FUSED and _lib.load().cspn_last_launch_count() == 1
    assert torch.isfinite(out_our).all()
    span = float(out_ref.abs().max())
    err, noise = float((out_ref - out_our).abs().max()), float((out_ref - out_ref2).abs().max())
    assert err <= 1e-4 * max(1.0, span) + 10 * noise, f"{which}: output max-abs {err:.3e} (range {span:.3e}, reference run-to-run {noise:.3e})"
    assert g_ref.keys() == g_our.keys() and len(g_ref) > 100
    self_worst, self_cos = compare(g_ref, g_ref2)
    worst, cos = compare(g_ref, g_our)
    print(f"{which}: output err {err:.3e} (noise {noise:.3e}); grads worst {worst:.3e} cos {cos:.7f} (reference vs itself: {self_worst:.3e}, {self_cos:.7f})")
    # The backward of a ResNet-50 with batch norm in training mode amplifies the 1e-6 relative difference of two valid fp32
    # summation orders inside the module: tolerance = 1e-3 of the tensor's largest entry, or 20x what the reference shows
    # against itself, whichever is larger; the overall gradient direction must agree to 1e-4.
    assert cos > 1.0 - 1e-4, f"{which}: gradient cosine {cos:.7f}"
    assert worst <= max(1e-2, 20 * self_worst), f"{which}: a parameter gradient differs by {worst:.3e} of its largest entry (reference vs itself {self_worst:.3e})"
    # one optimiser step on both keeps the weights together
    for model in (ref_model, our_model):
        torch.optim.SGD(model.parameters(), lr=1e-3).step()
    drift = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(ref_model.parameters(), our_model.parameters()))
    assert np.isfinite(drift) and drift <= 1e-4


def test_reference_unet_imports_b200_inplace_abn_on_multi_gpu_boxes(monkeypatch):
    """On a box with more than one GPU the reference's UNet files import `InPlaceABNSync` at module level
    (unet_cspn_nyu.py:19-25) and BasicBlock (resnet18) normalises with `partial(InPlaceABNSync, activation='none')`.  The
    reference's own package needs a cffi extension built against torch 0.4; `dropin.install_inplace_abn()` registers the B200
    classes under its name.  Simulated here with device_count() == 2: the multi-GPU flavour of resnet18 must build with the B200
    class and its BasicBlock stages agree with the single-GPU flavour (nn.BatchNorm2d) on the same weights - they differ only by
    gamma = |w| + eps.  (Forward only: BasicBlock's `out += residual` overwrites the tensor in-place ABN saves for its backward -
    the reference's own multi-GPU resnet18 cannot train for the same reason, see the note at unet_cspn_nyu.py:14-18.)"""
    import importlib.util
    from cspn_monodepth_b200 import abn
    unet_cspn_nyu, _ = _import_reference()
    dropin.install_inplace_abn()
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 2)
    spec = importlib.util.spec_from_file_location("unet_cspn_nyu_multigpu", os.path.join(REF, "network", "unet_cspn_nyu.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    monkeypatch.undo()
    torch.manual_seed(3)
    single_model = unet_cspn_nyu.resnet18(pretrained=False).to(DEV).eval()
    multi_model = multi.resnet18(pretrained=False).to(DEV).eval()
    assert isinstance(multi_model.layer1[0].bn1, abn.InPlaceABNSync) and multi_model.layer1[0].bn1.activation == "none"
    missing = multi_model.load_state_dict(single_model.state_dict(), strict=False)
    assert not missing.missing_keys and all(k.endswith("num_batches_tracked") for k in missing.unexpected_keys)
    # (the reference's resnet18 does not run end to end - its decoder is wired for the 2048 channels of resnet50 - so the encoder
    # stages built from BasicBlock are compared)
    x = torch.randn(2, 64, 57, 76, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        a = single_model.layer2(single_model.layer1(x))
        b = multi_model.layer2(multi_model.layer1(x.clone()))
    assert torch.isfinite(b).all()
    assert (a - b).abs().max().item() <= 2e-3 * max(1.0, a.abs().max().item())
    # training mode: batch statistics through the B200 kernels against nn.BatchNorm2d's, forward only
    single_model.train(); multi_model.train()
    with torch.no_grad():
        a, b = single_model.layer1[0].bn1(single_model.layer1[0].conv1(x)), multi_model.layer1[0].bn1(multi_model.layer1[0].conv1(x.clone()))
    assert (a - b).abs().max().item() <= 1e-4 * max(1.0, a.abs().max().item())
    assert torch.allclose(single_model.layer1[0].bn1.running_var, multi_model.layer1[0].bn1.running_var, rtol=1e-5, atol=1e-6)
