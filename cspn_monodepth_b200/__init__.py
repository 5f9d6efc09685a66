"""cspn_monodepth_b200 - B200-native (sm_100a) CSPN affinity propagation.

Drop-in replacements for the reference's ``AffinityPropagate`` modules backed by hand-written CUDA
kernels behind a C ABI (``include/cspn_b200.h``, ``cspn_monodepth_b200/lib/libcspn_b200.so``).
"""
from . import criteria, cspn_legacy, cspn_new, cspn_ours, dropin, heads, sharding  # noqa: F401
from ._lib import MODE_NEW, MODE_OURS, PATH_AUTO, PATH_FUSED, PATH_GENERIC, CspnError  # noqa: F401
from .functional import cspn_propagate  # noqa: F401

__all__ = ["cspn_new", "cspn_ours", "cspn_legacy", "heads", "criteria", "dropin", "sharding", "cspn_propagate", "CspnError",
           "MODE_NEW", "MODE_OURS", "PATH_AUTO", "PATH_FUSED", "PATH_GENERIC"]
