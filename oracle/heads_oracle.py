"""CPU restatement (numpy, float64 accumulation) of the two output heads upstream of the CSPN module.  TEST INFRASTRUCTURE
ONLY: imported by tests/ and bench.py's checks - never by the product.

  forward(x, weight, H, W)              Simple_Gudi_UpConv_Block_Last_Layer.forward, network/unet_cspn_nyu.py:214-217 (= unet_ours.py
                                        :199-202): _up_pooling (:201-212: nearest 2x upsample, crop to H x W, keep even rows / columns)
                                        followed by conv1 (3x3, padding 1, no bias)
  backward(x, weight, grad_out, H, W)   its autograd: (grad_x, grad_weight)
The restatement follows the reference literally (it builds the zero-inserted tensor and convolves all 9 taps); the CUDA kernels
never materialise it.  Pinned by tests/golden/heads_golden.npz (tests/golden/make_heads_golden.py runs the reference's classes).
"""
import numpy as np


def _unpool(x, H, W):
    b, c, h, w = x.shape
    u = np.zeros((b, c, H, W), np.float64)
    u[:, :, 0::2, 0::2] = x[:, :, :(H + 1) // 2, :(W + 1) // 2]          # :201-212
    return u


def forward(x, weight, H, W):
    x, weight = np.asarray(x, np.float64), np.asarray(weight, np.float64)
    u = np.pad(_unpool(x, H, W), [(0, 0), (0, 0), (1, 1), (1, 1)])
    out = np.zeros((x.shape[0], weight.shape[0], H, W), np.float64)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("oc,bchw->bohw", weight[:, :, ky, kx], u[:, :, ky:ky + H, kx:kx + W])
    return out


def backward(x, weight, grad_out, H, W):
    x, weight, go = np.asarray(x, np.float64), np.asarray(weight, np.float64), np.asarray(grad_out, np.float64)
    u = np.pad(_unpool(x, H, W), [(0, 0), (0, 0), (1, 1), (1, 1)])
    gu = np.zeros_like(u)
    gw = np.zeros_like(weight)
    for ky in range(3):
        for kx in range(3):
            gw[:, :, ky, kx] = np.einsum("bohw,bchw->oc", go, u[:, :, ky:ky + H, kx:kx + W])
            gu[:, :, ky:ky + H, kx:kx + W] += np.einsum("oc,bohw->bchw", weight[:, :, ky, kx], go)
    gx = np.zeros_like(x)
    gx[:, :, :(H + 1) // 2, :(W + 1) // 2] = gu[:, :, 1:-1, 1:-1][:, :, 0::2, 0::2]
    return gx, gw
