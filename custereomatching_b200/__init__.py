"""custereomatching_b200 - B200-native (sm_100a only) ZNCC cost-volume hot path of lzhnb/CuStereoMatching.

  csrc/         hand-written CUDA kernels + the C ABI (include/custma_b200.h) -> libcustma_b200.so
  build.py      in-tree nvcc build of that library
  binding.py    ctypes binding of the C ABI (no torch types cross it)
  functional.py torch-facing host layer: allocation, checks, autograd
  sharding.py   batch / row-band partitioning across the GPUs of one box (torch.distributed)

The drop-in Python surface of the reference lives in the top-level `custma` package.
"""
from . import binding
from .functional import (FLAG_DIRECT, FLAG_TENSOR, INVALID_COST, backward, backward_projector, confidence_mask, cost_volume, cost_volume_and_wta,
                         forward, ingest_u8, prepare_backward, soft_disparity, wta, wta_masked)

__all__ = ["binding", "forward", "backward", "prepare_backward", "backward_projector", "cost_volume", "wta", "wta_masked", "cost_volume_and_wta", "confidence_mask",
           "ingest_u8", "soft_disparity", "FLAG_DIRECT", "FLAG_TENSOR", "INVALID_COST"]
