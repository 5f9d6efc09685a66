"""The fused output heads (csrc/cspn_heads.cu through cspn_monodepth_b200/heads.py) against the reference-run vectors and the
numpy oracle.  Tolerances: forward 1e-5 of the output range (fp32 sums of 64..576 products in a different order than the
reference's convolution), gradients 1e-5 of their largest entry (grad_weight sums ~1e5 terms: 1e-4); fp16: + 2 fp16 ulps."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from cspn_monodepth_b200 import _lib, heads
from oracle import heads_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


def _close(got, want, rtol, what):
    got, want = got.detach().float().cpu().numpy().astype(np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    err, scale = np.abs(got - want).max(), max(1.0, np.abs(want).max())
    assert err <= rtol * scale, f"{what}: max err {err:.3e} (scale {scale:.3e})"


def test_matches_reference_vectors_fused_and_single():
    z = np.load(os.path.join(ROOT, "tests", "golden", "heads_golden.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 5
    for n in names:
        H, W = z[n + "/depth"].shape[2:]
        x = _cu(z[n + "/x"]).requires_grad_(True)
        wd, wg = _cu(z[n + "/w_depth"]).requires_grad_(True), _cu(z[n + "/w_guid"]).requires_grad_(True)
        d, g = heads.guidance_depth_heads(x, wd, wg, H, W)
        _close(d, z[n + "/depth"], 1e-5, n + " depth")
        _close(g, z[n + "/guidance"], 1e-5, n + " guidance")
        torch.autograd.backward([d, g], [_cu(z[n + "/grad_depth"]), _cu(z[n + "/grad_guidance"])])
        _close(x.grad, z[n + "/grad_x"], 1e-5, n + " grad_x")
        _close(wd.grad, z[n + "/grad_w_depth"], 1e-5, n + " grad_w_depth")
        _close(wg.grad, z[n + "/grad_w_guid"], 1e-5, n + " grad_w_guid")
        # a single head through the drop-in class (same constructor as the reference's)
        head = heads.Simple_Gudi_UpConv_Block_Last_Layer(x.shape[1], 1, H, W).to(DEV)
        with torch.no_grad():
            head.conv1.weight.copy_(wd)
        x2 = _cu(z[n + "/x"]).requires_grad_(True)
        y = head(x2)
        _close(y, z[n + "/depth"], 1e-5, n + " single head")
        y.backward(_cu(z[n + "/grad_depth"]))
        _close(x2.grad, z[n + "/grad_x_depth_only"], 1e-5, n + " single head grad_x")
        _close(head.conv1.weight.grad, z[n + "/grad_w_depth"], 1e-5, n + " single head grad_w")


@pytest.mark.parametrize("b,cin,h,w,H,W,ng,dtype", [(2, 64, 114, 152, 228, 304, 12, torch.float32), (2, 64, 114, 152, 228, 304, 8, torch.float32),
                                                     (1, 64, 176, 608, 352, 1216, 8, torch.float16), (2, 70, 19, 37, 37, 73, 12, torch.float32),
                                                     (1, 130, 9, 40, 18, 79, 15, torch.float32), (3, 3, 5, 3, 9, 5, 2, torch.float32)])
def test_full_size_vs_oracle(b, cin, h, w, H, W, ng, dtype):
    rng = np.random.default_rng(cin + W)
    x = rng.standard_normal((b, cin, h, w)).astype(np.float32)
    wd = (rng.standard_normal((1, cin, 3, 3)) / np.sqrt(cin)).astype(np.float32)
    wg = (rng.standard_normal((ng, cin, 3, 3)) / np.sqrt(cin)).astype(np.float32)
    god, gog = rng.standard_normal((b, 1, H, W)).astype(np.float32), rng.standard_normal((b, ng, H, W)).astype(np.float32)
    tx, twd, twg, tgod, tgog = (_cu(a, dtype) for a in (x, wd, wg, god, gog))
    xr, wdr, wgr, godr, gogr = (t.float().cpu().numpy() for t in (tx, twd, twg, tgod, tgog))      # what the kernels see
    tx.requires_grad_(True); twd.requires_grad_(True); twg.requires_grad_(True)
    d, g = heads.guidance_depth_heads(tx, twd, twg, H, W)
    lib = _lib.load()
    assert lib.cspn_last_launch_count() == 1
    ulp = 2.0 ** -9 if dtype == torch.float16 else 0.0
    _close(d, heads_oracle.forward(xr, wdr, H, W), 1e-5 + ulp, "depth")
    _close(g, heads_oracle.forward(xr, wgr, H, W), 1e-5 + ulp, "guidance")
    torch.autograd.backward([d, g], [tgod, tgog])
    gx1, gw1 = heads_oracle.backward(xr, wdr, godr, H, W)
    gx2, gw2 = heads_oracle.backward(xr, wgr, gogr, H, W)
    _close(tx.grad, gx1 + gx2, 1e-5 + ulp, "grad_x")
    _close(twd.grad, gw1, 1e-4 + ulp, "grad_w_depth")
    _close(twg.grad, gw2, 1e-4 + ulp, "grad_w_guid")
    # deterministic (fixed-order split-K reduction, no floating-point atomics)
    tx2, twd2, twg2 = (t.detach().clone().requires_grad_(True) for t in (tx, twd, twg))
    d2, g2 = heads.guidance_depth_heads(tx2, twd2, twg2, H, W)
    torch.autograd.backward([d2, g2], [tgod, tgog])
    assert torch.equal(twg2.grad, twg.grad) and torch.equal(tx2.grad, tx.grad) and torch.equal(d2, d)


def test_paired_heads_share_one_launch_in_a_model():
    class Tail(nn.Module):                                                       # the tail of unet_cspn_nyu.ResNet.forward (:383-384)
        def __init__(self):
            super().__init__()
            self.gud_up_proj_layer5 = heads.Simple_Gudi_UpConv_Block_Last_Layer(64, 1, 57, 75)
            self.gud_up_proj_layer6 = heads.Simple_Gudi_UpConv_Block_Last_Layer(64, 12, 57, 75)

        def forward(self, x):
            guidance = self.gud_up_proj_layer6(x)
            return self.gud_up_proj_layer5(x), guidance

    torch.manual_seed(0)
    plain = Tail().to(DEV)
    x = torch.randn(2, 64, 29, 38, device=DEV, requires_grad=True)
    d0, g0 = plain(x)
    (d0.sum() + (g0 * g0).sum()).backward()
    ref = [x.grad.clone(), plain.gud_up_proj_layer5.conv1.weight.grad.clone(), plain.gud_up_proj_layer6.conv1.weight.grad.clone()]
    keys = list(plain.state_dict().keys())
    fused = heads.fuse_heads(plain)
    assert list(fused.state_dict().keys()) == keys
    x.grad = None
    fused.zero_grad(set_to_none=True)
    d1, g1 = fused(x)
    assert torch.equal(d0, d1) and torch.equal(g0, g1)                           # same kernel, stacked weights: bit-identical
    (d1.sum() + (g1 * g1).sum()).backward()
    got = [x.grad, fused.gud_up_proj_layer5.conv1.weight.grad, fused.gud_up_proj_layer6.conv1.weight.grad]
    for a, b_ in zip(got, ref):
        assert float((a - b_).abs().max()) <= 1e-5 * max(1.0, float(b_.abs().max()))
    x2 = torch.randn(2, 64, 29, 38, device=DEV)                                  # a new input must not pick up a parked result
    with torch.no_grad():
        d2, _ = fused(x2)
        d3 = fused.gud_up_proj_layer5(x2)
    assert torch.equal(d2, d3) and not torch.equal(d2, d1.detach())


def test_errors():
    x = torch.randn(1, 64, 8, 8, device=DEV)
    w1 = torch.randn(1, 64, 3, 3, device=DEV)
    with pytest.raises(RuntimeError):
        heads.guidance_depth_heads(x.cpu(), w1.cpu(), None, 16, 16)
    with pytest.raises(RuntimeError):
        heads.guidance_depth_heads(x, w1, None, 17, 16)                          # not a crop of the 16 x 16 unpooled tensor
    with pytest.raises(RuntimeError):
        heads.guidance_depth_heads(x, torch.randn(1, 32, 3, 3, device=DEV), None, 16, 16)
    with pytest.raises(RuntimeError):
        heads.guidance_depth_heads(x, w1, torch.randn(16, 64, 3, 3, device=DEV), 16, 16)    # 17 outputs > 16
    d, g = heads.guidance_depth_heads(x, w1, None, 15, 16)
    assert g is None and tuple(d.shape) == (1, 1, 15, 16)
