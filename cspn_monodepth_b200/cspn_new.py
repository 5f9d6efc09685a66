"""Drop-in for ``network/libs/post_process/CSPN_new.py`` of the reference (mode A).

Same class name, constructor and ``forward`` signature as the reference module
(``CSPN_new.py:17-24`` and ``:26``), so ``unet_cspn_nyu.py:357-358,386`` works unchanged:

    self.post_process_layer = AffinityPropagate(24, 3)
    x = self.post_process_layer(guidance, x, sparse_depth)
"""
import torch.nn as nn

from . import _lib
from .functional import cspn_propagate


class AffinityPropagate(nn.Module):
    def __init__(self, prop_time, prop_kernel):
        super().__init__()
        self.prop_time = prop_time
        self.prop_kernel = prop_kernel
        self.in_feature = 1
        self.out_feature = 1

    def forward(self, guidance, blur_depth, sparse_depth=None):
        if self.prop_time > 0 and self.prop_kernel != 3:
            # the reference only works for 3 (CSPN_new.py:122 builds a (k//2)-sized ones kernel; 5 -> shape error)
            raise RuntimeError(f"CSPN_new.AffinityPropagate supports prop_kernel=3 only, got {self.prop_kernel}")
        return cspn_propagate(guidance, blur_depth, sparse_depth, iters=self.prop_time, ksize=3, mode=_lib.MODE_NEW)

    def extra_repr(self):
        return f"prop_time={self.prop_time}, prop_kernel={self.prop_kernel}"
