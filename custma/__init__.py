"""custma - drop-in for lzhnb/CuStereoMatching's Python package (custma/__init__.py:2-6), backed by the B200-native
kernels in custereomatching_b200.  Same exports: __version__, stereo_matching, Timer; `__all__` is every public
global, as in the reference.  New surface (banded / batched / fused WTA / sharding) is exported additionally."""
import torch as _torch  # noqa: F401  (the native module below needs torch's libraries and type casters loaded first)

from .version import __version__
try:
    from . import src  # noqa: F401  (native module, built by setup.py / custereomatching_b200.build)
except ImportError as _e:  # no Python or CPU stand-in exists for it
    raise ImportError("custma.src (the native torch module) is not built: run `python setup.py build_ext --inplace` "
                      "or `python -m custereomatching_b200.build`") from _e
from .stereo_matching_wrapper import stereo_matching
from .utils import Timer
from .stereo_matching_wrapper import (stereo_matching_banded, stereo_matching_wta, stereo_matching_with_wta,
                                      stereo_matching_masked, stereo_matching_soft_disparity,
                                      stereo_matching_projector_grad, cost_volume_mask)

__all__ = [k for k in globals().keys() if not k.startswith("_")]
