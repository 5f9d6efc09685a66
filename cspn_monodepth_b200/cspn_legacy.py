"""Drop-in for the reference's legacy max-of-8 CSPN, ``network/libs/post_process/CSPN.py``.

Same class names, constructors and ``forward`` signatures as the reference (``CSPN.py:14-19`` and ``:126-132``):

    AffinityPropagate(spn=False).forward(guidance, blur_depth, sparse_depth)
    AffinityPropagate_prediction(spn=False).forward(guidance, blur_depth)

16 steps (``CSPN.py:35``; ``prop_time`` is exposed as an attribute), each ``r = max_k box3x3(|g_k| r) / box3x3(|g_k|)`` followed
by the re-injection of the sparse samples.  The reference module only runs on CUDA (``.cuda()`` at ``:87``) and issues ~700 ATen
launches; here it is 4 launches of ``csrc/cspn_legacy.cu`` (4 steps each).  Forward only: the module is dead code in the
reference (nothing imports it) and was never trained through in this repository - a tensor that requires grad raises instead of
silently detaching.
"""
import torch
import torch.nn as nn

from . import _lib
from .functional import _SUFFIX, _guidance_view, _ptr


def legacy_propagate(guidance, blur_depth, sparse_depth=None, iters=16):
    for name, t in (("guidance", guidance), ("blur_depth", blur_depth), ("sparse_depth", sparse_depth)):
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{name} is on {t.device}: the B200 CSPN operator is CUDA-only and has no CPU fallback")
        if t.dtype != guidance.dtype or t.device != guidance.device:
            raise RuntimeError(f"{name}: dtype / device differ from guidance ({t.dtype} on {t.device})")
        if torch.is_grad_enabled() and t.requires_grad:
            raise RuntimeError("the legacy max-of-8 CSPN kernel is forward-only; wrap the call in torch.no_grad() or use CSPN_new")
    if guidance.dtype not in _SUFFIX:
        raise RuntimeError(f"unsupported dtype {guidance.dtype}: float32 and float16 only")
    if guidance.dim() != 4 or blur_depth.dim() != 4 or blur_depth.shape[1] != 1:
        raise RuntimeError("guidance must be [B,>=8,H,W] and blur_depth [B,1,H,W]")
    b, _, h, w = blur_depth.shape
    if guidance.shape[0] != b or tuple(guidance.shape[2:]) != (h, w) or guidance.shape[1] < 8:
        raise RuntimeError(f"guidance {tuple(guidance.shape)} does not match blur_depth {tuple(blur_depth.shape)} (needs 8 channels)")
    if sparse_depth is not None and tuple(sparse_depth.shape) != tuple(blur_depth.shape):
        raise RuntimeError(f"sparse_depth {tuple(sparse_depth.shape)} does not match blur_depth {tuple(blur_depth.shape)}")
    lib = _lib.load()
    g, gbs = _guidance_view(guidance, h, w)
    d = blur_depth.contiguous()
    s = None if sparse_depth is None else sparse_depth.contiguous()
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        nbytes = lib.cspn_legacy_workspace_bytes(b, h, w, iters)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device) if nbytes else None
        _lib.check(getattr(lib, "cspn_legacy_fwd_" + _SUFFIX[d.dtype])(
            g.data_ptr(), gbs, d.data_ptr(), _ptr(s), out.data_ptr(), b, h, w, iters, _ptr(ws), nbytes,
            torch.cuda.current_stream(d.device).cuda_stream))
    return out


class AffinityPropagate(nn.Module):
    """``CSPN.py:14-56``."""

    def __init__(self, spn=False):
        super().__init__()
        self.spn = spn
        self.prop_time = 16

    def forward(self, guidance, blur_depth, sparse_depth):
        return legacy_propagate(guidance, blur_depth, sparse_depth, self.prop_time)


class AffinityPropagate_prediction(nn.Module):
    """``CSPN.py:126-164``."""

    def __init__(self, spn=False):
        super().__init__()
        self.spn = spn
        self.prop_time = 16

    def forward(self, guidance, blur_depth):
        return legacy_propagate(guidance, blur_depth, None, self.prop_time)
