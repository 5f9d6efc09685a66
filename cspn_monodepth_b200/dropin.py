"""Ways to put the B200 modules into the reference's networks without editing them.

* :func:`install` registers this package's modules under the reference's module names
  (``network.libs.post_process.CSPN_new`` / ``CSPN_ours``) so that
  ``from network.libs.post_process.CSPN_new import AffinityPropagate`` (``unet_cspn_nyu.py:9``) and
  ``from network.libs.post_process.CSPN_ours import AffinityPropagate`` (``unet_ours.py:16``) pick them up.
  Call it before importing the UNet files.
* :func:`install_inplace_abn` registers :mod:`cspn_monodepth_b200.abn` as ``network.libs.inplace_abn`` (needed on multi-GPU boxes,
  where the UNet files import ``InPlaceABNSync``); :func:`install` does it as well.
* :func:`patch_model` (``heads=True``: also the two output heads upstream, see :mod:`cspn_monodepth_b200.heads`) swaps ``model.post_process_layer`` on an already-built reference model
  (``unet_cspn_nyu.py:357-358`` / ``unet_ours.py:304-305``).  The module has no parameters or buffers,
  so checkpoints are unaffected.
"""
import sys

from . import cspn_new, cspn_ours


def install(inplace_abn=True):
    sys.modules["network.libs.post_process.CSPN_new"] = cspn_new
    sys.modules["network.libs.post_process.CSPN_ours"] = cspn_ours
    pkg = sys.modules.get("network.libs.post_process")
    if pkg is not None:
        pkg.CSPN_new, pkg.CSPN_ours = cspn_new, cspn_ours
    if inplace_abn:
        install_inplace_abn()


def install_inplace_abn():
    """Register :mod:`cspn_monodepth_b200.abn` as ``network.libs.inplace_abn``: on a box with more than one GPU both UNet files do
    ``from network.libs.inplace_abn import InPlaceABNSync`` at import time (``unet_cspn_nyu.py:19-25``, ``unet_ours.py:23-28``), and
    the reference's own package needs a cffi extension built against torch 0.4.  Call it before importing the UNet files."""
    from . import abn
    sys.modules["network.libs.inplace_abn"] = abn
    pkg = sys.modules.get("network.libs")
    if pkg is not None:
        pkg.inplace_abn = abn


def patch_model(model, heads=False):
    """Replace every reference ``AffinityPropagate`` inside ``model`` by the B200 one. Returns the count.

    ``heads=True`` also replaces the two ``Simple_Gudi_UpConv_Block_Last_Layer`` heads that feed it
    (``gud_up_proj_layer5`` / ``gud_up_proj_layer6``, ``unet_cspn_nyu.py:331-332`` / ``unet_ours.py:278-279``) by the fused B200
    heads of :mod:`cspn_monodepth_b200.heads` (one launch for both, parameters kept)."""
    if heads and hasattr(model, "gud_up_proj_layer5") and hasattr(model, "gud_up_proj_layer6"):
        from .heads import fuse_heads
        fuse_heads(model)
    n = 0
    for parent in model.modules():
        for name, child in list(parent.named_children()):
            if type(child).__name__ != "AffinityPropagate" or isinstance(child, (cspn_new.AffinityPropagate, cspn_ours.AffinityPropagate)):
                continue
            if hasattr(child, "times"):        # CSPN_ours.py:20-22
                setattr(parent, name, cspn_ours.AffinityPropagate(child.times))
            else:                              # CSPN_new.py:19-24
                setattr(parent, name, cspn_new.AffinityPropagate(child.prop_time, child.prop_kernel))
            n += 1
    return n
