"""Golden vectors for the two output heads upstream of the CSPN module, produced by RUNNING THE
REFERENCE's own classes (build container only):

    python tests/golden/make_heads_golden.py      ->  tests/golden/heads_golden.npz

`Simple_Gudi_UpConv_Block_Last_Layer` of network/unet_cspn_nyu.py:195-218 (nearest upsample + crop + Python mask loop + conv3x3)
and of network/unet_ours.py:194-202 over MyBlock._up_pooling :138-150 (conv_transpose2d zero insertion + crop + conv3x3) run
unmodified on CPU fp32 with autograd; the two heads of a model see the same x, as at unet_cspn_nyu.py:383-384.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _install_thnn_shim  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _install_thnn_shim()
    from network import unet_cspn_nyu, unet_ours
    torch.manual_seed(99)
    out = {}
    # (module file, guidance channels, Cin, h, w, oheight, owidth): even / odd crops, the NYU aspect, a few-channel case
    cases = (("nyu_12", unet_cspn_nyu, 12, 64, 9, 12, 18, 24), ("nyu_odd_crop", unet_cspn_nyu, 12, 64, 7, 10, 13, 19),
             ("ours_8", unet_ours, 8, 64, 10, 9, 20, 18), ("ours_odd_crop", unet_ours, 8, 64, 6, 11, 11, 21), ("thin", unet_ours, 8, 5, 4, 35, 8, 70))
    for name, mod, ng, cin, h, w, oh, ow in cases:
        depth_head = mod.Simple_Gudi_UpConv_Block_Last_Layer(cin, 1, oh, ow)
        guid_head = mod.Simple_Gudi_UpConv_Block_Last_Layer(cin, ng, oh, ow)
        x = torch.randn(2, cin, h, w, requires_grad=True)
        d, g = depth_head(x), guid_head(x)
        god, gog = torch.randn_like(d), torch.randn_like(g)
        (d * god).sum().backward(retain_graph=True)
        gx_d = x.grad.clone()
        ((g * gog).sum()).backward()
        for k, v in (("x", x), ("w_depth", depth_head.conv1.weight), ("w_guid", guid_head.conv1.weight), ("depth", d), ("guidance", g),
                     ("grad_depth", god), ("grad_guidance", gog), ("grad_x", x.grad), ("grad_x_depth_only", gx_d),
                     ("grad_w_depth", depth_head.conv1.weight.grad), ("grad_w_guid", guid_head.conv1.weight.grad)):
            out[f"{name}/{k}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "heads_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
