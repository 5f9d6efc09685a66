"""Golden vectors with the value distribution the reference's OWN networks feed into the CSPN module.

Run in the build container only (imports ``/root/reference``):

    python tests/golden/make_unet_golden.py

Builds the reference UNets unmodified (``network/unet_cspn_nyu.py`` -> CSPN_new, ``network/unet_ours.py`` -> CSPN_ours,
random initialisation with a fixed seed, no checkpoint), runs one forward on a seeded RGB-D input and records what
``ResNet.forward`` hands to ``self.post_process_layer`` (``unet_cspn_nyu.py:386`` / ``unet_ours.py:333``): the
guidance head's output (12 resp. 8 channels), the blur-depth head's output and the sparse-depth channel.  A window
at the image corner (two image borders) is cut out of these tensors and the reference module itself is re-run on
the window - forward and autograd gradients - so the fixture stays small.  Writes ``tests/golden/cspn_unet_heads.npz``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _install_thnn_shim  # noqa: E402

WIN_H, WIN_W = 72, 100


def capture(model, x):
    got = {}

    def hook(mod, args, kwargs, output):
        got["args"], got["kwargs"], got["out"] = args, kwargs, output
    h = model.post_process_layer.register_forward_hook(hook, with_kwargs=True)
    with torch.no_grad():
        model(x)
    h.remove()
    return got


def main():
    sys.path.insert(0, REF)
    _install_thnn_shim()
    from network import unet_cspn_nyu, unet_ours
    from network.libs.post_process import CSPN_new, CSPN_ours
    torch.manual_seed(1234)
    torch.set_num_threads(8)
    rgb = torch.rand(1, 3, 228, 304)
    dense = torch.rand(1, 1, 228, 304) * 9 + 0.5
    mask = torch.rand(1, 1, 228, 304) < 500.0 / 69312.0
    x = torch.cat([rgb, dense * mask], dim=1)                      # the reference's rgbd input: channel 3 = sparse depth
    out = {}
    rng = np.random.default_rng(7)

    # ---- unet_cspn_nyu: post_process_layer(guidance[12 ch], blur_depth, sparse_depth)   (CSPN_new, mode A)
    model = unet_cspn_nyu.resnet50(pretrained=False).eval()
    got = capture(model, x)
    guidance, blur, sparse = got["args"][:3]
    assert guidance.shape[1] == 12 and model.post_process_layer.prop_time == 24
    g, d, s = (t[:, :, :WIN_H, :WIN_W].contiguous().clone() for t in (guidance, blur, sparse))
    tg, td = g.clone().requires_grad_(True), d.clone().requires_grad_(True)
    y = CSPN_new.AffinityPropagate(24, 3)(tg, td, s)
    go = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(go)
    out.update({"A_unet/guidance": g.numpy(), "A_unet/depth": d.numpy(), "A_unet/sparse": s.numpy(), "A_unet/iters": np.int64(24),
                "A_unet/out": y.detach().numpy(), "A_unet/grad_out": go.numpy(), "A_unet/grad_guidance": tg.grad.numpy(), "A_unet/grad_depth": td.grad.numpy()})
    print("mode A: guidance |g| mean %.3g max %.3g, blur range [%.3g, %.3g], %d sparse hits in the window" % (
        g.abs().mean(), g.abs().max(), d.min(), d.max(), int((s != 0).sum())))
    del model

    # ---- unet_ours: post_process_layer(blur_depth, guidance[8 ch], sparse_depth=...)    (CSPN_ours, mode B)
    model = unet_ours.resnet50(pretrained=False).eval()
    got = capture(model, x)
    blur, guidance = got["args"][:2]
    sparse = got["kwargs"]["sparse_depth"]
    assert guidance.shape[1] == 8 and model.post_process_layer.times == 24
    g, d, s = (t[:, :, :WIN_H, :WIN_W].contiguous().clone() for t in (guidance, blur, sparse))
    tg, td = g.clone().requires_grad_(True), d.clone().requires_grad_(True)
    y = CSPN_ours.AffinityPropagate(prop_time=24)(td, tg, sparse_depth=s)
    go = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(go)
    out.update({"B_unet/guidance": g.numpy(), "B_unet/depth": d.numpy(), "B_unet/sparse": s.numpy(), "B_unet/iters": np.int64(24),
                "B_unet/out": y.detach().numpy(), "B_unet/grad_out": go.numpy(), "B_unet/grad_guidance": tg.grad.numpy(), "B_unet/grad_depth": td.grad.numpy()})
    print("mode B: guidance mean %.3g max %.3g, blur range [%.3g, %.3g]" % (g.abs().mean(), g.abs().max(), d.min(), d.max()))

    path = os.path.join(HERE, "cspn_unet_heads.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; torch", torch.__version__)


if __name__ == "__main__":
    main()
