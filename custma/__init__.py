"""custma - drop-in for lzhnb/CuStereoMatching's Python package (custma/__init__.py:2-6), backed by the B200-native
kernels in custereomatching_b200.  Same exports: __version__, stereo_matching, Timer; `__all__` is every public
global, as in the reference.  New surface (banded / batched / fused WTA / sharding) is exported additionally."""
from .version import __version__
from .stereo_matching_wrapper import stereo_matching
from .utils import Timer
from .stereo_matching_wrapper import (stereo_matching_banded, stereo_matching_wta, stereo_matching_with_wta,
                                      cost_volume_mask)

__all__ = [k for k in globals().keys() if not k.startswith("_")]
