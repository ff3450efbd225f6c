"""crick_b200 -- a B200-native (sm_100a) implementation of crick's sketches.

Same Python API as the reference (crick/__init__.py:1-16): ``TDigest``, ``SummaryStats``,
``SpaceSaving``, plus the grouped entry point (``TDigestGroup`` / ``grouped_update``).
All numeric work runs in hand-written CUDA kernels behind the C ABI of
``libcrick_cuda.so`` (include/crick_cuda.h); there is no CPU fallback.
"""
import numpy as _np

from . import _lib
from ._arrays import DeviceArray, PinnedArray
from .space_saving import SpaceSaving, TopKResult
from .stats import SummaryStats
from .tdigest import TDigest, TDigestGroup, grouped_update

__version__ = "0.1.0"
__git_revision__ = None
__numpy_version__ = _np.__version__.encode()


def device_count():
    return _lib.lib.crick_device_count()


def launch_count():
    """Kernels launched by libcrick_cuda since it was loaded."""
    return int(_lib.lib.crick_launch_count())


__all__ = ["TDigest", "SummaryStats", "SpaceSaving", "TopKResult", "TDigestGroup",
           "grouped_update", "DeviceArray", "PinnedArray", "device_count", "launch_count"]
