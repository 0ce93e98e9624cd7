"""Torch-facing host layer over the C ABI: allocation, input checks, autograd glue.

Mirrors the reference's host entry points for the hot path and nothing else:
  stereo_matching_forward   custma/src/stereo_matching.cpp:16-42  (CHECK_INPUT on both images, callee allocates)
  stereo_matching_backward  custma/src/stereo_matching.cpp:45-73  (CHECK_INPUT on the gradient only, H/W from it)
  _StereoMatching           custma/stereo_matching_wrapper.py:7-35 (camera-only gradient)
and adds the banded / batched / fused-WTA entry points BASELINE.json's configs name.  Differences from the
reference that a caller can observe, all deliberate:
  * kernels run on torch's CURRENT stream of the inputs' device (the reference uses legacy stream 0),
  * outputs are written exactly once (no torch::zeros pre-fill), sizes are 64-bit,
  * the backward is deterministic (no atomics),
  * shape / dtype mismatches raise RuntimeError instead of being undefined behaviour (the reference's assert at
    stereo_matching.cpp:28 is compiled out).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import binding

INVALID_COST = binding.INVALID_COST
FLAG_DIRECT = binding.FLAG_DIRECT
FLAG_TENSOR = binding.FLAG_TENSOR


def _check_input(x: torch.Tensor, name: str) -> None:
    # same messages as CHECK_CUDA / CHECK_CONTIGUOUS (custma/include/stereo_matching.hpp:20-25)
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(x).__name__}")
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if x.dtype != torch.float32:
        # the reference's data_ptr<float>() throws for any other dtype (stereo_matching_kernel.cu:209-212)
        raise RuntimeError(f"{name} must be a float32 tensor, got {x.dtype}")


def _shape_bhw(camera: torch.Tensor, projector: torch.Tensor) -> Tuple[int, int, int, bool]:
    if camera.dim() not in (2, 3):
        raise RuntimeError(f"camera must be [H, W] or [B, H, W], got {tuple(camera.shape)}")
    if camera.shape != projector.shape:
        raise RuntimeError(f"camera {tuple(camera.shape)} and projector {tuple(projector.shape)} must have the same shape")
    if camera.device != projector.device:
        raise RuntimeError("camera and projector must be on the same device")
    batched = camera.dim() == 3
    B = camera.shape[0] if batched else 1
    H, W = camera.shape[-2], camera.shape[-1]
    if B == 0 or H == 0 or W == 0:
        raise RuntimeError(f"empty input {tuple(camera.shape)}")
    return B, H, W, batched


def _aligned16(x: torch.Tensor) -> torch.Tensor:
    """The banded kernels move the volume with 128-bit accesses (include/custma_b200.h, "Alignment"): a contiguous view
    whose storage offset is not a multiple of 16 bytes is copied to a fresh (512-byte aligned) allocation."""
    return x if x.data_ptr() % 16 == 0 else x.clone(memory_format=torch.contiguous_format)


def _workspace(nbytes: int, device) -> Tuple[Optional[torch.Tensor], int]:
    if nbytes == 0:
        return None, 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)  # caching allocator: 512-byte aligned
    return ws, ws.data_ptr()


def forward(camera: torch.Tensor, projector: torch.Tensor, D: int = 0, kernel_size: int = 5, *,
            want_cost: bool = True, want_wta: bool = False, flags: int = 0, want_mask: bool = False,
            mask_threshold: float = 0.6):
    """ZNCC cost volume and / or winner-take-all.

    camera, projector: float32 CUDA, [H,W] or [B,H,W].  D == 0: reference-shaped volume [...,H,W,W] whose last axis
    is the projector column; D > 0: banded volume [...,H,W,D] whose last axis is the disparity s (projector column
    w - s; cells with w - s < 0 hold INVALID_COST).  Returns (cost | None, best | None, index | None); index is
    int32: first maximal projector column (D == 0) or the disparity of it (D > 0).
    want_mask (needs want_wta): two more results from the same decode kernel - the confidence mask best > mask_threshold
    (examples/verify.py:74) and the masked disparity (column - correspondence) * mask (examples/test.py:83-84), both fp32.
    """
    _check_input(camera, "camera")
    _check_input(projector, "projector")
    B, H, W, batched = _shape_bhw(camera, projector)
    if not (want_cost or want_wta):
        raise RuntimeError("nothing to compute: want_cost and want_wta are both False")
    if want_mask and not want_wta:
        raise RuntimeError("want_mask needs want_wta")
    D = int(D)
    k = int(kernel_size)
    C = D if D > 0 else W
    lead = (B,) if batched else ()
    with torch.cuda.device(camera.device):
        cost = torch.empty(lead + (H, W, C), dtype=torch.float32, device=camera.device) if want_cost else None
        best = torch.empty(lead + (H, W), dtype=torch.float32, device=camera.device) if want_wta else None
        index = torch.empty(lead + (H, W), dtype=torch.int32, device=camera.device) if want_wta else None
        mask = torch.empty(lead + (H, W), dtype=torch.float32, device=camera.device) if want_mask else None
        mdisp = torch.empty(lead + (H, W), dtype=torch.float32, device=camera.device) if want_mask else None
        nbytes = binding.forward_workspace_bytes(B, H, W, D, k, flags)
        if nbytes == 0:  # the query validates the arguments; every valid problem needs a non-empty workspace
            binding.check(binding.ERR_INVALID_ARGUMENT, "custma_forward_workspace_bytes")
        ws, ws_ptr = _workspace(nbytes, camera.device)
        stream = torch.cuda.current_stream(camera.device).cuda_stream
        if want_mask:
            binding.forward_wta(camera.data_ptr(), projector.data_ptr(), cost.data_ptr() if cost is not None else 0,
                                best.data_ptr(), index.data_ptr(), mask.data_ptr(), mdisp.data_ptr(), mask_threshold,
                                B, H, W, D, k, flags, ws_ptr, nbytes, stream)
        else:
            binding.forward(camera.data_ptr(), projector.data_ptr(),
                            cost.data_ptr() if cost is not None else 0,
                            best.data_ptr() if best is not None else 0,
                            index.data_ptr() if index is not None else 0,
                            B, H, W, D, k, flags, ws_ptr, nbytes, stream)
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream(camera.device))
    if want_mask:
        return cost, best, index, mask, mdisp
    return cost, best, index


class PreparedBackward:
    """custma_backward_prepare run ahead of time: the workspace that holds the image-dependent half of the backward
    and the stream it was filled on.  backward(..., prepared=...) waits for that stream and skips the preparation."""

    def __init__(self, ws: torch.Tensor, nbytes: int, stream: "torch.cuda.Stream", key):
        self.ws, self.nbytes, self.stream, self.key = ws, nbytes, stream, key


_prepare_streams = {}


def prepare_backward(camera: torch.Tensor, projector: torch.Tensor, kernel_size: int, D: int = 0, *,
                     flags: int = 0) -> PreparedBackward:
    """Starts the image-dependent part of backward() (pivots, band copies, window statistics, verdict) on a side stream,
    so that it runs beside the forward and the loss.  The images must not change until the backward has run."""
    _check_input(camera, "camera")
    _check_input(projector, "projector")
    B, H, W, _ = _shape_bhw(camera, projector)
    D, k = int(D), int(kernel_size)
    dev = camera.device
    with torch.cuda.device(dev):
        nbytes = binding.backward_workspace_bytes(B, H, W, D, k, flags)
        if nbytes == 0:
            binding.check(binding.ERR_INVALID_ARGUMENT, "custma_backward_workspace_bytes")
        main = torch.cuda.current_stream(dev)
        side = _prepare_streams.get(dev.index)
        if side is None:
            side = _prepare_streams[dev.index] = torch.cuda.Stream(dev)
        side.wait_stream(main)      # the images are ready, and the workspace is not older work's any more
        ws, ws_ptr = _workspace(nbytes, dev)
        binding.backward_prepare(camera.data_ptr(), projector.data_ptr(), B, H, W, D, k, flags, ws_ptr, nbytes,
                                 side.cuda_stream)
        ws.record_stream(side)
    return PreparedBackward(ws, nbytes, side, (camera.data_ptr(), projector.data_ptr(), B, H, W, D, k, int(flags)))


def backward(cost_volume_grad: torch.Tensor, camera: torch.Tensor, projector: torch.Tensor, kernel_size: int,
             D: int = 0, *, flags: int = 0, rows: Optional[Tuple[int, int]] = None,
             prepared: Optional[PreparedBackward] = None) -> torch.Tensor:
    """Gradient of sum(cost * cost_volume_grad) with respect to the camera image; same leading shape as camera.

    rows = (row_begin, row_end): the upstream gradient exists only on those volume rows and cost_volume_grad is
    [..., row_end - row_begin, W, C] (custma_backward_rows: what a row-band shard passes for its owned rows).
    prepared = prepare_backward(camera, projector, ...) of these very images: its half of the work is not repeated."""
    _check_input(cost_volume_grad, "cost_volume_grad")
    _check_input(camera, "camera")
    _check_input(projector, "projector")
    B, H, W, batched = _shape_bhw(camera, projector)
    D = int(D)
    C = D if D > 0 else W
    lead = (B,) if batched else ()
    r0, r1 = (0, H) if rows is None else (int(rows[0]), int(rows[1]))
    if not 0 <= r0 < r1 <= H:
        raise RuntimeError(f"rows {rows} must be a non-empty range inside [0, {H})")
    if tuple(cost_volume_grad.shape) != lead + (r1 - r0, W, C):
        raise RuntimeError(f"cost_volume_grad must have shape {lead + (r1 - r0, W, C)}, got {tuple(cost_volume_grad.shape)}")
    if cost_volume_grad.device != camera.device:
        raise RuntimeError("cost_volume_grad must be on the images' device")
    k = int(kernel_size)
    cost_volume_grad = _aligned16(cost_volume_grad)
    with torch.cuda.device(camera.device):
        camera_grad = torch.empty_like(camera)
        nbytes = binding.backward_workspace_bytes(B, H, W, D, k, flags)
        if nbytes == 0:
            binding.check(binding.ERR_INVALID_ARGUMENT, "custma_backward_workspace_bytes")
        if prepared is not None:
            if prepared.key != (camera.data_ptr(), projector.data_ptr(), B, H, W, D, k, int(flags)):
                raise RuntimeError("prepared belongs to other images, another shape or other flags")
            torch.cuda.current_stream(camera.device).wait_stream(prepared.stream)
            ws, ws_ptr, call_flags = prepared.ws, prepared.ws.data_ptr(), flags | binding.FLAG_PREPARED
        else:
            (ws, ws_ptr), call_flags = _workspace(nbytes, camera.device), flags
        stream = torch.cuda.current_stream(camera.device).cuda_stream
        if rows is None:
            binding.backward(cost_volume_grad.data_ptr(), camera.data_ptr(), projector.data_ptr(),
                             camera_grad.data_ptr(), B, H, W, D, k, call_flags, ws_ptr, nbytes, stream)
        else:
            binding.backward_rows(cost_volume_grad.data_ptr(), camera.data_ptr(), projector.data_ptr(),
                                  camera_grad.data_ptr(), B, H, W, D, k, r0, r1, call_flags, ws_ptr, nbytes, stream)
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream(camera.device))
    return camera_grad


def backward_projector(cost_volume_grad: torch.Tensor, camera: torch.Tensor, projector: torch.Tensor, kernel_size: int,
                       D: int = 0, *, flags: int = 0) -> torch.Tensor:
    """Gradient of sum(cost * cost_volume_grad) with respect to the PROJECTOR image (the reference has none:
    custma/stereo_matching_wrapper.py:33 returns None); same leading shape as projector."""
    _check_input(cost_volume_grad, "cost_volume_grad")
    _check_input(camera, "camera")
    _check_input(projector, "projector")
    B, H, W, batched = _shape_bhw(camera, projector)
    D, k = int(D), int(kernel_size)
    C = D if D > 0 else W
    lead = (B,) if batched else ()
    if tuple(cost_volume_grad.shape) != lead + (H, W, C):
        raise RuntimeError(f"cost_volume_grad must have shape {lead + (H, W, C)}, got {tuple(cost_volume_grad.shape)}")
    with torch.cuda.device(camera.device):
        projector_grad = torch.empty_like(projector)
        nbytes = binding.backward_projector_workspace_bytes(B, H, W, D, k, flags)
        if nbytes == 0:
            binding.check(binding.ERR_INVALID_ARGUMENT, "custma_backward_projector_workspace_bytes")
        ws, ws_ptr = _workspace(nbytes, camera.device)
        stream = torch.cuda.current_stream(camera.device)
        binding.backward_projector(cost_volume_grad.data_ptr(), camera.data_ptr(), projector.data_ptr(),
                                   projector_grad.data_ptr(), B, H, W, D, k, flags, ws_ptr, nbytes, stream.cuda_stream)
        ws.record_stream(stream)
    return projector_grad


class _CostVolume(torch.autograd.Function):
    """Differentiable cost volume, banded or full, batched or not.  The camera always gets its gradient (as in the
    reference); the projector gets one too when it requires it (an extension: the reference returns None for it)."""

    @staticmethod
    def forward(ctx, camera, projector, D, kernel_size, flags):
        ctx.save_for_backward(camera, projector)
        ctx.D, ctx.kernel_size, ctx.flags = int(D), int(kernel_size), int(flags)
        # the camera gradient's image-dependent half starts now, beside the forward and whatever the loss does
        # (not for tiny volumes: below ~32 Mcell the second stream's events cost more than they hide)
        C = int(D) if int(D) > 0 else camera.shape[-1]
        ctx.prepared = prepare_backward(camera, projector, kernel_size, D, flags=flags) \
            if camera.requires_grad and camera.numel() * C >= 32e6 else None
        cost, _, _ = forward(camera, projector, D, kernel_size, want_cost=True, want_wta=False, flags=flags)
        return cost

    @staticmethod
    def backward(ctx, cost_volume_grad):
        camera, projector = ctx.saved_tensors
        cost_volume_grad = cost_volume_grad.contiguous()
        g = backward(cost_volume_grad, camera, projector, ctx.kernel_size, ctx.D, flags=ctx.flags, prepared=ctx.prepared) \
            if ctx.needs_input_grad[0] else None
        ctx.prepared = None
        gp = backward_projector(cost_volume_grad, camera, projector, ctx.kernel_size, ctx.D) \
            if ctx.needs_input_grad[1] else None
        return g, gp, None, None, None


def cost_volume(camera, projector, D: int = 0, kernel_size: int = 5, flags: int = 0) -> torch.Tensor:
    """Differentiable (camera only) ZNCC volume: [...,H,W,W] for D == 0, banded [...,H,W,D] for D > 0."""
    return _CostVolume.apply(camera, projector, D, kernel_size, flags)


def wta(camera, projector, D: int = 0, kernel_size: int = 5, flags: int = 0):
    """Fused forward + winner-take-all without materialising the volume: (best fp32, index int32).

    Equals torch.max(cost_volume, -1) of examples/verify.py:72 (first maximal projector column); for D > 0 the
    index is the disparity w - column (examples/test.py:83)."""
    _, best, index = forward(camera, projector, D, kernel_size, want_cost=False, want_wta=True, flags=flags)
    return best, index


def cost_volume_and_wta(camera, projector, D: int = 0, kernel_size: int = 5, flags: int = 0):
    """(cost, best, index) from one pass over the volume (not differentiable; use cost_volume for autograd)."""
    return forward(camera, projector, D, kernel_size, want_cost=True, want_wta=True, flags=flags)


def wta_masked(camera, projector, D: int = 0, kernel_size: int = 5, threshold: float = 0.6, flags: int = 0):
    """Fused forward + winner-take-all + confidence mask, no volume: (best, index, mask, masked_disparity).

    mask = best > threshold (examples/verify.py:72-74, cost_volume_threshold = 0.6 at :13) and
    masked_disparity = (column - correspondence) * mask (examples/test.py:83-84) come out of the kernel that decodes the
    WTA keys - no torch op touches the results."""
    _, best, index, mask, mdisp = forward(camera, projector, D, kernel_size, want_cost=False, want_wta=True, flags=flags,
                                          want_mask=True, mask_threshold=threshold)
    return best, index, mask, mdisp


class _SoftDisparity(torch.autograd.Function):
    """Fused head: masked soft-argmax disparity straight from the images; backward through the same kernels, no volume."""

    @staticmethod
    def forward(ctx, camera, projector, D, kernel_size, beta, threshold):
        _check_input(camera, "camera")
        _check_input(projector, "projector")
        B, H, W, batched = _shape_bhw(camera, projector)
        D, k = int(D), int(kernel_size)
        lead = (B,) if batched else ()
        with torch.cuda.device(camera.device):
            dev = camera.device
            soft = torch.empty(lead + (H, W), dtype=torch.float32, device=dev)
            best = torch.empty(lead + (H, W), dtype=torch.float32, device=dev)
            index = torch.empty(lead + (H, W), dtype=torch.int32, device=dev)
            mask = torch.empty(lead + (H, W), dtype=torch.float32, device=dev)
            state = torch.empty(lead + (H, W, 4), dtype=torch.float32, device=dev)
            nbytes = binding.head_workspace_bytes(B, H, W, D, k, 0)
            if nbytes == 0:
                raise RuntimeError(f"the fused disparity head needs kernel_size 3 or 5 and valid sizes (got kernel_size={k}, "
                                   f"shape {tuple(camera.shape)}, D={D})")
            ws, ws_ptr = _workspace(nbytes, dev)
            stream = torch.cuda.current_stream(dev)
            binding.forward_head(camera.data_ptr(), projector.data_ptr(), soft.data_ptr(), best.data_ptr(), index.data_ptr(),
                                 mask.data_ptr(), state.data_ptr(), beta, threshold, B, H, W, D, k, 0, ws_ptr, nbytes,
                                 stream.cuda_stream)
            ws.record_stream(stream)
        ctx.save_for_backward(camera, projector, state)
        ctx.cfg = (B, H, W, D, k, float(beta))
        ctx.mark_non_differentiable(best, index, mask)
        return soft, best, index, mask

    @staticmethod
    def backward(ctx, soft_grad, _gb, _gi, _gm):
        camera, projector, state = ctx.saved_tensors
        B, H, W, D, k, beta = ctx.cfg
        soft_grad = soft_grad.contiguous()
        if soft_grad.dtype != torch.float32:
            raise RuntimeError(f"soft_disparity gradient must be float32, got {soft_grad.dtype}")
        with torch.cuda.device(camera.device):
            camera_grad = torch.empty_like(camera)
            nbytes = binding.head_workspace_bytes(B, H, W, D, k, 0)
            ws, ws_ptr = _workspace(nbytes, camera.device)
            stream = torch.cuda.current_stream(camera.device)
            binding.backward_head(soft_grad.data_ptr(), camera.data_ptr(), projector.data_ptr(), state.data_ptr(), beta,
                                  camera_grad.data_ptr(), B, H, W, D, k, 0, ws_ptr, nbytes, stream.cuda_stream)
            ws.record_stream(stream)
        return camera_grad, None, None, None, None, None


def soft_disparity(camera, projector, D: int = 0, kernel_size: int = 5, beta: float = 50.0, threshold: float = 0.6):
    """Fused differentiable disparity head: (soft_disparity * mask, best, index, mask), differentiable w.r.t. the camera.

    soft_disparity = sum_s softmax_s(beta * cost)[s] * s (examples/verify.py:31-39 with softargmax_beta = 50 at :11; the
    disparity of examples/test.py:85-86), mask = best > threshold (verify.py:72-74).  The cost volume is never written:
    forward and backward recompute it tile by tile inside the kernels.  threshold < -1 gives the unmasked soft disparity."""
    return _SoftDisparity.apply(camera, projector, D, kernel_size, beta, threshold)


def ingest_u8(image_u8: torch.Tensor, channel: int = 0, scale: float = 1.0 / 255.0) -> torch.Tensor:
    """uint8 CUDA image [H,W], [H,W,Ch], [B,H,W,Ch] -> float32 plane(s) of one channel times scale
    (examples/verify.py:138-142,149: cv2.imread(...) / 255, then [:, :, 0])."""
    if not isinstance(image_u8, torch.Tensor) or not image_u8.is_cuda:
        raise RuntimeError("image must be a CUDA tensor")
    if image_u8.dtype != torch.uint8:
        raise RuntimeError(f"image must be a uint8 tensor, got {image_u8.dtype}")
    if not image_u8.is_contiguous():
        raise RuntimeError("image must be contiguous")
    if image_u8.dim() == 2:
        lead, (H, W), ch = (), image_u8.shape, 1
    elif image_u8.dim() == 3:
        lead, (H, W, ch) = (), image_u8.shape
    elif image_u8.dim() == 4:
        lead, (H, W, ch) = (image_u8.shape[0],), image_u8.shape[1:]
    else:
        raise RuntimeError(f"image must be [H,W], [H,W,Ch] or [B,H,W,Ch], got {tuple(image_u8.shape)}")
    B = lead[0] if lead else 1
    if B == 0 or H == 0 or W == 0:
        raise RuntimeError(f"empty input {tuple(image_u8.shape)}")
    with torch.cuda.device(image_u8.device):
        out = torch.empty(lead + (H, W), dtype=torch.float32, device=image_u8.device)
        binding.ingest_u8(image_u8.data_ptr(), out.data_ptr(), B, H, W, ch, int(channel), scale,
                          torch.cuda.current_stream(image_u8.device).cuda_stream)
    return out


def confidence_mask(best: torch.Tensor, threshold: float = 0.6) -> torch.Tensor:
    """examples/verify.py:74 as a torch expression (for results that already exist); the fused form is wta_masked()."""
    return (best > threshold).to(best.dtype)
