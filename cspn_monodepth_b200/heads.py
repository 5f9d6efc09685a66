"""Drop-in for the two output heads that feed the CSPN module in the reference's UNets.

``Simple_Gudi_UpConv_Block_Last_Layer(in_channels, out_channels, oheight, owidth)`` - same name, constructor and ``conv1``
parameter as ``network/unet_cspn_nyu.py:195-218`` / ``network/unet_ours.py:194-202``, so checkpoints load unchanged.  The
reference unpools by 2 (nearest upsample, crop to ``oheight x owidth``, multiply by a mask of the even positions that an
O(H W) Python loop fills one element at a time, ``unet_cspn_nyu.py:208-212``) and runs a dense 3x3 convolution over a tensor
that is 3/4 zeros.  ``csrc/cspn_heads.cu`` computes the same result from the half-resolution input directly (9/4 taps per
output pixel, no unpooled tensor) and - :func:`guidance_depth_heads` / :func:`fuse_heads` - BOTH heads of a model in one pass
over ``x`` (``unet_cspn_nyu.py:383-384`` applies them to the same tensor).  Forward and backward (x and both weights) are CUDA
kernels behind the C ABI (``cspn_heads_fwd_*`` / ``cspn_heads_bwd_*``); CUDA tensors only, no CPU fallback.
"""
from __future__ import annotations

import threading

import torch
import torch.nn as nn

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float16: "f16"}
_scratch = {}


def _workspace(device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _scratch.get(key)
    if ws is None:
        ws = torch.empty(_lib.load().cspn_heads_workspace_bytes(), dtype=torch.uint8, device=device)
        _scratch[key] = ws
    return ws


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Heads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, w2, oheight, owidth):
        for name, t in (("x", x), ("weight", w1), ("weight", w2)):
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError(f"{name} is on {t.device}: the B200 head kernels are CUDA-only and have no CPU fallback")
            if t.dtype != x.dtype or t.device != x.device:
                raise RuntimeError(f"{name}: dtype / device differ from x ({t.dtype} on {t.device})")
        if x.dtype not in _SUFFIX:
            raise RuntimeError(f"unsupported dtype {x.dtype}: float32 and float16 only")
        if x.dim() != 4 or w1.dim() != 4 or tuple(w1.shape[1:]) != (x.shape[1], 3, 3) or (w2 is not None and tuple(w2.shape[1:]) != (x.shape[1], 3, 3)):
            raise RuntimeError(f"x {tuple(x.shape)} / weights {tuple(w1.shape)} {None if w2 is None else tuple(w2.shape)}: need [B,Cin,h,w] and [n,Cin,3,3]")
        b, cin, h, w = x.shape
        if not (1 <= oheight <= 2 * h and 1 <= owidth <= 2 * w):
            raise RuntimeError(f"output size {oheight} x {owidth} is not a crop of the 2x unpooled {2 * h} x {2 * w}")
        lib = _lib.load()
        xc, w1c = x.contiguous(), w1.contiguous()
        w2c = None if w2 is None else w2.contiguous()
        n1, n2 = w1.shape[0], 0 if w2 is None else w2.shape[0]
        out1 = torch.empty(b, n1, oheight, owidth, dtype=x.dtype, device=x.device)
        out2 = None if w2 is None else torch.empty(b, n2, oheight, owidth, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(getattr(lib, "cspn_heads_fwd_" + _SUFFIX[x.dtype])(
                xc.data_ptr(), w1c.data_ptr(), _ptr(w2c), out1.data_ptr(), _ptr(out2), b, cin, h, w, oheight, owidth, n1, n2,
                torch.cuda.current_stream(x.device).cuda_stream))
        ctx.save_for_backward(xc, w1c, w2c)
        ctx.size = (oheight, owidth)
        return out1, out2

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go1, go2):
        x, w1, w2 = ctx.saved_tensors
        oheight, owidth = ctx.size
        lib = _lib.load()
        b, cin, h, w = x.shape
        n1, n2 = w1.shape[0], 0 if w2 is None else w2.shape[0]
        go1 = go1.contiguous()
        go2 = None if w2 is None else go2.contiguous()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or (w2 is not None and ctx.needs_input_grad[2])
        gx = torch.empty_like(x) if need_x else None
        gw1 = torch.empty_like(w1) if need_w else None
        gw2 = torch.empty_like(w2) if (need_w and w2 is not None) else None
        with torch.cuda.device(x.device):
            ws = _workspace(x.device)
            _lib.check(getattr(lib, "cspn_heads_bwd_" + _SUFFIX[x.dtype])(
                x.data_ptr(), w1.data_ptr(), _ptr(w2), go1.data_ptr(), _ptr(go2), _ptr(gx), _ptr(gw1), _ptr(gw2),
                b, cin, h, w, oheight, owidth, n1, n2, ws.data_ptr(), ws.numel(), torch.cuda.current_stream(x.device).cuda_stream))
        return gx, gw1, gw2, None, None


def guidance_depth_heads(x, w_depth, w_guidance, oheight, owidth):
    """Both heads in one pass over ``x``: returns ``(blur_depth [B,n1,H,W], guidance [B,n2,H,W])``."""
    return _Heads.apply(x, w_depth, w_guidance, int(oheight), int(owidth))


class Simple_Gudi_UpConv_Block_Last_Layer(nn.Module):
    """``unet_cspn_nyu.py:195-218`` / ``unet_ours.py:194-202``: unpool x2 (zero insertion, cropped) + conv3x3, no bias."""

    def __init__(self, in_channels, out_channels, oheight=0, owidth=0):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.oheight = oheight
        self.owidth = owidth
        self._pair = None          # set by fuse_heads: (shared state, my role)

    def forward(self, x):
        if self._pair is not None:
            return self._pair[0].get(self._pair[1], x)
        return _Heads.apply(x, self.conv1.weight, None, self.oheight, self.owidth)[0]

    def extra_repr(self):
        return f"oheight={self.oheight}, owidth={self.owidth}, fused={self._pair is not None}"


class _Pair:
    """Two heads that the model applies to the same tensor one after the other (``unet_cspn_nyu.py:383-384``): the first call
    computes both with one launch and parks the sibling's result, the second call picks it up.  Per-thread state, so the
    reference's thread-per-GPU DataParallel replicas (``encoding.py:102-105``) do not see each other."""

    def __init__(self, depth_head, guid_head):
        self.heads = {"depth": depth_head, "guid": guid_head}
        self.local = threading.local()

    def get(self, role, x):
        key = (x.data_ptr(), x._version, tuple(x.shape), x.device)
        parked = getattr(self.local, "parked", None)
        if parked is not None and parked[0] == key and parked[1] == role:
            self.local.parked = None
            return parked[2]
        d, g = self.heads["depth"], self.heads["guid"]
        depth, guidance = _Heads.apply(x, d.conv1.weight, g.conv1.weight, d.oheight, d.owidth)
        self.local.parked = (key, "guid" if role == "depth" else "depth", guidance if role == "depth" else depth)
        return depth if role == "depth" else guidance


def fuse_heads(model, depth_name="gud_up_proj_layer5", guid_name="gud_up_proj_layer6"):
    """Replace the reference model's two last-layer heads (``unet_cspn_nyu.py:331-332`` / ``unet_ours.py:278-279``) by B200
    heads that share one launch.  Parameters are moved, not copied, so optimiser state and checkpoints are unaffected."""
    old_d, old_g = getattr(model, depth_name), getattr(model, guid_name)
    new = []
    for old in (old_d, old_g):
        cin, cout = old.conv1.in_channels, old.conv1.out_channels
        head = Simple_Gudi_UpConv_Block_Last_Layer(cin, cout, old.oheight, old.owidth)
        head.conv1 = old.conv1
        new.append(head)
    if (new[0].oheight, new[0].owidth) != (new[1].oheight, new[1].owidth):
        raise RuntimeError("the two heads produce different output sizes")
    pair = _Pair(new[0], new[1])
    new[0]._pair, new[1]._pair = (pair, "depth"), (pair, "guid")
    setattr(model, depth_name, new[0])
    setattr(model, guid_name, new[1])
    return model
