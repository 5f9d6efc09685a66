"""Drop-in for ``network/libs/post_process/CSPN_ours.py`` of the reference (mode B, PAC variant).

Same class name, constructor and ``forward`` signature (``CSPN_ours.py:18-24``; note the argument
order differs from CSPN_new: ``forward(x, guided, sparse_depth=None)``), so ``unet_ours.py:304-305,333``
works unchanged.  The kernel size follows the guidance channel count, K = int(sqrt(C+1)) (``:32``);
3x3 and 5x5 (and 7x7) are supported.
"""
import torch.nn as nn

from . import _lib
from .functional import cspn_propagate, kernel_size_from_channels


class AffinityPropagate(nn.Module):
    def __init__(self, prop_time):
        super().__init__()
        self.times = prop_time

    def forward(self, x, guided, sparse_depth=None):
        ksize = kernel_size_from_channels(guided.shape[1])
        return cspn_propagate(guided, x, sparse_depth, iters=self.times, ksize=ksize, mode=_lib.MODE_OURS)

    def extra_repr(self):
        return f"prop_time={self.times}"
