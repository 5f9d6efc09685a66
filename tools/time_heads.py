"""Timing of the fused output heads (csrc/cspn_heads.cu) against the reference's formulation on the same GPU.

    python tools/time_heads.py [--cfg nyu|kitti]

Ours: cspn_heads_fwd / cspn_heads_bwd through cspn_monodepth_b200.heads (both heads, one forward launch).  Reference formulation:
unet_ours.py's heads (conv_transpose2d zero insertion + crop + conv3x3 per head, cuDNN; TF32 on = PyTorch's default, and off) -
unet_cspn_nyu.py's Python mask loop (:208-212, 17k one-element assignments per call) is timed once for context.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cspn_monodepth_b200 import heads  # noqa: E402

CFGS = {"nyu": (8, 64, 114, 152, 228, 304, 12), "kitti": (8, 64, 176, 608, 352, 1216, 8)}


def cuda_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def ref_heads(x, wd, wg, H, W):
    c = x.shape[1]
    k = torch.zeros(c, 1, 2, 2, device=x.device, dtype=x.dtype)
    k[:, :, 0, 0] = 1
    u = F.conv_transpose2d(x, k, stride=2, groups=c)[:, :, :H, :W]
    return F.conv2d(u, wd, padding=1), F.conv2d(u, wg, padding=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="nyu")
    a = ap.parse_args()
    b, cin, h, w, H, W, ng = CFGS[a.cfg]
    dev = "cuda:0"
    torch.manual_seed(0)
    x = torch.randn(b, cin, h, w, device=dev, requires_grad=True)
    wd = (torch.randn(1, cin, 3, 3, device=dev) / 24).requires_grad_(True)
    wg = (torch.randn(ng, cin, 3, 3, device=dev) / 24).requires_grad_(True)
    god, gog = torch.randn(b, 1, H, W, device=dev), torch.randn(b, ng, H, W, device=dev)
    flops = 2.0 * b * h * w * 9 * cin * (1 + ng)
    bytes_fwd = 4.0 * (b * cin * h * w + b * (1 + ng) * H * W)

    def ours_fwd():
        with torch.no_grad():
            return heads.guidance_depth_heads(x, wd, wg, H, W)

    def ours_fwd_bwd():
        d, g = heads.guidance_depth_heads(x, wd, wg, H, W)
        torch.autograd.backward([d, g], [god, gog])

    def ref_fwd():
        with torch.no_grad():
            return ref_heads(x, wd, wg, H, W)

    def ref_fwd_bwd():
        d, g = ref_heads(x, wd, wg, H, W)
        torch.autograd.backward([d, g], [god, gog])

    t = cuda_time(ours_fwd)
    print(f"{a.cfg}: ours fwd       {t:9.1f} us   {flops / t / 1e6:7.2f} TFLOP/s (useful)   {bytes_fwd / t / 1e3:7.1f} GB/s algorithmic")
    t2 = cuda_time(ours_fwd_bwd)
    print(f"{a.cfg}: ours fwd+bwd   {t2:9.1f} us")
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        print(f"{a.cfg}: torch fwd      {cuda_time(ref_fwd):9.1f} us   (unet_ours formulation, cuDNN, TF32 {'on' if tf32 else 'off'})")
        print(f"{a.cfg}: torch fwd+bwd  {cuda_time(ref_fwd_bwd):9.1f} us")
    d0, g0 = ours_fwd()
    d1, g1 = ref_fwd()
    print(f"max-abs vs torch fp32: depth {(d0 - d1).abs().max().item():.2e} guidance {(g0 - g1).abs().max().item():.2e}")
    # the NYU variant's mask loop, once (unet_cspn_nyu.py:208-212)
    t0 = time.perf_counter()
    mask = torch.zeros(b, cin, H, W, device=dev)
    for hh in range(0, H, 2):
        for ww in range(0, W, 2):
            mask[:, :, hh, ww] = 1
    torch.cuda.synchronize()
    print(f"{a.cfg}: unet_cspn_nyu.py mask loop alone: {(time.perf_counter() - t0) * 1e3:.0f} ms per head and forward")


if __name__ == "__main__":
    main()
