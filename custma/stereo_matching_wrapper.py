"""The autograd entry point of the reference (custma/stereo_matching_wrapper.py:7-35), kept verbatim in behaviour:
`stereo_matching(camera_image, projector_image, D, kernel_size)` is `Function.apply` (positional only), saves both
images, and returns gradients `(camera_grad, None, None, None)`."""
from typing import Tuple

import torch

from custereomatching_b200 import functional as _F
from .src import stereo_matching_backward, stereo_matching_forward


class _StereoMatching(torch.autograd.Function):
    @staticmethod
    def forward(ctx, camera_image: torch.Tensor, projector_image: torch.Tensor, D: int,
                kernel_size: int) -> torch.Tensor:
        # camera_image, projector_image: [H, W] float32 CUDA contiguous -> cost volume [H, W, W]
        ctx.save_for_backward(camera_image, projector_image)
        ctx.D = D
        ctx.kernel_size = kernel_size
        return stereo_matching_forward(camera_image, projector_image, D, kernel_size)

    @staticmethod
    def backward(ctx, cost_volume_grad: torch.Tensor) -> Tuple:
        camera_image, projector_image = ctx.saved_tensors
        # the reference hands the incoming gradient to CHECK_INPUT unchanged (stereo_matching.cpp:52):
        # a non-contiguous gradient raises there and here
        grad = stereo_matching_backward(cost_volume_grad, camera_image, projector_image, ctx.kernel_size)
        return grad, None, None, None


stereo_matching = _StereoMatching.apply


# ---- extensions beyond the reference surface (SURVEY.md section 8a "banded extension", 8f) ----------------------
def stereo_matching_banded(camera_image: torch.Tensor, projector_image: torch.Tensor, D: int,
                           kernel_size: int) -> torch.Tensor:
    """Differentiable banded volume [..., H, W, D]: band[h, w, s] = full[h, w, w - s]; accepts [H,W] or [B,H,W]."""
    if D <= 0:
        raise RuntimeError(f"D must be positive for the banded volume, got {D}")
    return _F.cost_volume(camera_image, projector_image, D, kernel_size)


def stereo_matching_wta(camera_image: torch.Tensor, projector_image: torch.Tensor, D: int, kernel_size: int):
    """(best, disparity) over D disparities without materialising the volume (examples/verify.py:72-74 fused).
    D == 0 searches every projector column and returns the correspondence column instead of a disparity."""
    return _F.wta(camera_image, projector_image, D, kernel_size)


def stereo_matching_with_wta(camera_image: torch.Tensor, projector_image: torch.Tensor, D: int, kernel_size: int):
    """(cost_volume, best, index) in one pass; not differentiable."""
    return _F.cost_volume_and_wta(camera_image, projector_image, D, kernel_size)


def stereo_matching_masked(camera_image: torch.Tensor, projector_image: torch.Tensor, D: int, kernel_size: int,
                           threshold: float = 0.6):
    """(best, index, mask, masked_disparity): examples/verify.py:72-74 and examples/test.py:78-84 in one fused call -
    mask = best > threshold, masked_disparity = (column - correspondence) * mask; no volume is materialised."""
    return _F.wta_masked(camera_image, projector_image, D, kernel_size, threshold)


def stereo_matching_soft_disparity(camera_image: torch.Tensor, projector_image: torch.Tensor, D: int, kernel_size: int,
                                   beta: float = 50.0, threshold: float = 0.6):
    """(soft_disparity * mask, best, index, mask), differentiable w.r.t. the camera image: the soft_argmax of
    examples/verify.py:31-39 (softargmax_beta = 50, :11) as the disparity of examples/test.py:85-86, fused into the
    kernels - forward and backward never materialise the cost volume.  kernel_size 3 or 5."""
    return _F.soft_disparity(camera_image, projector_image, D, kernel_size, beta, threshold)


def stereo_matching_projector_grad(cost_volume_grad: torch.Tensor, camera_image: torch.Tensor,
                                   projector_image: torch.Tensor, D: int, kernel_size: int) -> torch.Tensor:
    """Gradient with respect to the projector image (the reference returns None for it,
    custma/stereo_matching_wrapper.py:33)."""
    return _F.backward_projector(cost_volume_grad, camera_image, projector_image, kernel_size, D)


def cost_volume_mask(best: torch.Tensor, threshold: float = 0.6) -> torch.Tensor:
    """examples/verify.py:13,74."""
    return _F.confidence_mask(best, threshold)
