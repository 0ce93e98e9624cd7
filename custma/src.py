"""custma.src - the native module of the reference (custma/src/bindings.cpp:4-7) with the same two functions,
positional signatures and error behaviour:

    stereo_matching_forward(camera, projector, D, kernel_size) -> Tensor[H, W, W]
    stereo_matching_backward(cost_volume_grad, camera, projector, kernel_size) -> Tensor[H, W]

Both go straight to libcustma_b200.so through its C ABI (include/custma_b200.h); there is no other code path.
As in the reference, D is accepted and ignored (custma/src/stereo_matching_kernel.cu:14): the volume is [H, W, W]
with the projector column on the last axis.  Honouring D is `custma.stereo_matching_banded`.
"""
import torch

from custereomatching_b200 import functional as _F


def stereo_matching_forward(camera: torch.Tensor, projector: torch.Tensor, D: int, kernel_size: int) -> torch.Tensor:
    """custma/src/stereo_matching.cpp:16-42.  CHECK_INPUT on camera and projector; returns a new [H,W,W] tensor."""
    if camera.dim() != 2:
        raise RuntimeError(f"camera must be a 2-D [H, W] tensor, got {tuple(camera.shape)}")
    cost, _, _ = _F.forward(camera, projector, 0, kernel_size, want_cost=True, want_wta=False)
    return cost


def stereo_matching_backward(cost_volume_grad: torch.Tensor, camera: torch.Tensor, projector: torch.Tensor,
                             kernel_size: int) -> torch.Tensor:
    """custma/src/stereo_matching.cpp:45-73.  cost_volume_grad must be a contiguous [H,W,W] CUDA tensor (:52);
    returns camera_grad [H,W]."""
    if cost_volume_grad.dim() != 3:
        raise RuntimeError(f"cost_volume_grad must be a 3-D [H, W, W] tensor, got {tuple(cost_volume_grad.shape)}")
    return _F.backward(cost_volume_grad, camera, projector, kernel_size, 0)
