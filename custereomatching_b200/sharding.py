"""Partitioning of the hot path across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no multi-GPU code at all; the two partitionings BASELINE.json names are:

  * batch sharding      - stereo pairs are independent: rank r owns a contiguous slice of the batch; nothing is
                          exchanged in forward or backward; results ([B,H,W] best / disparity / camera_grad) are
                          gathered with one all_gather each.
  * row-band sharding   - for one very large pair: output row h depends only on image rows h-r .. h+k-1-r
                          (custma/src/stereo_matching_kernel.cu:44,60), so rank r owns rows [h0,h1) of the volume and
                          needs the image rows [h0-r, h1+k-1-r) (the window-radius halo).  Forward needs no exchange;
                          the camera gradient of a band reaches the same haloed rows, so overlapping halo rows of
                          neighbouring ranks are summed when the bands are assembled.

The cost volume and its gradient stay sharded; volume-sized data never crosses GPUs.  Collectives run on whatever
backend the process group has (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def split_even(n: int, parts: int) -> List[Tuple[int, int]]:
    """[begin, end) of each of `parts` contiguous slices of range(n); sizes differ by at most one, larger first."""
    if parts <= 0:
        raise ValueError("parts must be positive")
    base, extra = divmod(n, parts)
    out, pos = [], 0
    for p in range(parts):
        size = base + (1 if p < extra else 0)
        out.append((pos, pos + size))
        pos += size
    return out


def batch_slice(B: int, rank: int, world: int) -> Tuple[int, int]:
    return split_even(B, world)[rank]


@dataclass(frozen=True)
class RowBand:
    h0: int      # first owned volume row
    h1: int      # one past the last owned volume row
    lo: int      # first image row needed (h0 - r, clamped)
    hi: int      # one past the last image row needed (h1 + k - 1 - r, clamped)

    @property
    def rows(self) -> int:
        return self.h1 - self.h0

    @property
    def top_halo(self) -> int:
        return self.h0 - self.lo


def row_band(H: int, kernel_size: int, rank: int, world: int) -> RowBand:
    if world > H:
        # checked identically on every rank BEFORE any collective: an empty band would fail on its own rank while the
        # others wait in all_gather
        raise ValueError(f"row-band sharding needs at least one volume row per rank: H={H} < world={world}")
    h0, h1 = split_even(H, world)[rank]
    r = kernel_size // 2
    if h1 == h0:
        return RowBand(h0, h1, h0, h0)
    return RowBand(h0, h1, max(0, h0 - r), min(H, h1 + kernel_size - 1 - r))


def crop_rows_with_halo(img: torch.Tensor, band: RowBand) -> torch.Tensor:
    """Rows [lo, hi) of an [H,W] image (contiguous copy).

    The cropped image is zero-padded by the kernels exactly where the full image is out of range only at the true
    image border; inside the image the halo rows supply the real neighbours, so rows [h0,h1) computed on the crop
    equal the same rows computed on the full image."""
    return img[band.lo:band.hi].contiguous()


def band_rows_of(result_on_crop: torch.Tensor, band: RowBand) -> torch.Tensor:
    """Picks the owned rows [h0,h1) out of a tensor computed on the haloed crop (first axis = crop rows)."""
    return result_on_crop[band.top_halo:band.top_halo + band.rows]


def band_gradient_mask_rows(band: RowBand) -> Tuple[int, int]:
    """Crop-relative [begin,end) of the owned rows: the upstream gradient of halo rows must be zero on this rank
    (those volume rows belong to the neighbours)."""
    return band.top_halo, band.top_halo + band.rows


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _all_gather_ragged(local: torch.Tensor, sizes: Sequence[int], group=None) -> List[torch.Tensor]:
    """all_gather of tensors whose first axis differs per rank (sizes[r] rows on rank r): pad to the largest, gather
    into one buffer with a single collective, hand back trimmed views."""
    world = len(sizes)
    top = max(sizes)
    tail = tuple(local.shape[1:])
    if local.shape[0] == top:
        padded = local.contiguous()
    else:
        padded = torch.zeros((top,) + tail, dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    out = torch.empty((world * top,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return [out[r * top:r * top + sizes[r]] for r in range(world)]


def all_gather_batch(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """Gathers batch-sharded results: local is [b_r, ...] with b_r = size of batch_slice(B, rank, world)."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [e - b for b, e in split_even(B, world)]
    if local.shape[0] != sizes[rank]:
        raise RuntimeError(f"rank {rank} holds {local.shape[0]} pairs, expected {sizes[rank]}")
    if len(set(sizes)) == 1:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    return torch.cat(_all_gather_ragged(local, sizes, group), dim=0)


def all_gather_row_bands(local_rows: torch.Tensor, H: int, kernel_size: int, group=None) -> torch.Tensor:
    """Gathers row-band-sharded [rows_r, W] results (best score, disparity) into [H, W]."""
    rank, world = _world(group)
    if world == 1:
        return local_rows
    sizes = [e - b for b, e in split_even(H, world)]
    if local_rows.shape[0] != sizes[rank]:
        raise RuntimeError(f"rank {rank} holds {local_rows.shape[0]} rows, expected {sizes[rank]}")
    if len(set(sizes)) == 1:
        out = torch.empty((H,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    return torch.cat(_all_gather_ragged(local_rows, sizes, group), dim=0)


def assemble_row_band_gradient(grad_on_crop: torch.Tensor, H: int, kernel_size: int, group=None) -> torch.Tensor:
    """Sums the per-rank camera gradients (each defined on the rank's haloed crop rows [lo,hi)) into [H, W].

    Only the halo rows overlap between neighbours; the exchange is an all_gather of the (rows + halo) bands followed
    by a local overlap-add, i.e. (H + 2*world*r) * W * 4 bytes per rank - image-sized, never volume-sized."""
    rank, world = _world(group)
    bands = [row_band(H, kernel_size, r, world) for r in range(world)]
    mine = bands[rank]
    if grad_on_crop.shape[0] != mine.hi - mine.lo:
        raise RuntimeError(f"rank {rank}: gradient has {grad_on_crop.shape[0]} rows, expected {mine.hi - mine.lo}")
    W = grad_on_crop.shape[1]
    if world == 1:
        return grad_on_crop
    parts = _all_gather_ragged(grad_on_crop, [b.hi - b.lo for b in bands], group)
    out = torch.zeros((H, W), dtype=grad_on_crop.dtype, device=grad_on_crop.device)
    for b, part in zip(bands, parts):      # fixed order -> every rank computes identical bits
        out[b.lo:b.hi] += part
    return out


# ---- GPU-side drivers (need the CUDA library; exercised by the -m gpu tests and bench.py) ------------------------
def batch_sharded_step(camera: torch.Tensor, projector: torch.Tensor, D: int, kernel_size: int,
                       cost_volume_grad_fn=None, gather: bool = True, group=None, global_batch: Optional[int] = None):
    """One forward(+WTA)(+backward) over THIS rank's batch slice [b_r,H,W] already resident on its GPU.

    cost_volume_grad_fn(cost) -> upstream gradient (the caller's loss); None skips the backward.
    global_batch: the size of the whole batch (the slices are batch_slice(global_batch, rank, world)); when None it is
    found with one small all_gather + host read per call - pass it in a training loop.
    Returns (best, disparity, camera_grad | None), gathered over ranks to the full batch when gather=True."""
    from . import functional as F
    rank, world = _world(group)
    # the backward's image-dependent half runs on a side stream beside the forward and the loss
    extra = {"prepared": F.prepare_backward(camera, projector, kernel_size, D)} \
        if cost_volume_grad_fn is not None and camera.is_cuda else {}
    cost, best, disp = F.forward(camera, projector, D, kernel_size, want_cost=cost_volume_grad_fn is not None,
                                 want_wta=True)
    grad = None
    if cost_volume_grad_fn is not None:
        grad = F.backward(cost_volume_grad_fn(cost), camera, projector, kernel_size, D, **extra)
    if gather and world > 1:
        B = int(global_batch) if global_batch is not None else sum(_gather_sizes(camera.shape[0], group))
        best = all_gather_batch(best, B, group)
        disp = all_gather_batch(disp, B, group)
        if grad is not None:
            grad = all_gather_batch(grad, B, group)
    return best, disp, grad


def _gather_sizes(local_b: int, group=None) -> Sequence[int]:
    rank, world = _world(group)
    t = torch.tensor([local_b], dtype=torch.int64)
    if world == 1:
        return [local_b]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = t.to(dev)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return [int(p.item()) for p in parts]


def row_band_sharded_step(camera: torch.Tensor, projector: torch.Tensor, D: int, kernel_size: int,
                          cost_volume_grad_fn=None, group=None):
    """One very large pair split by rows: camera/projector are the FULL [H,W] images on this rank's GPU (images are
    small next to the volume: 133 MB at 8K against 68 GB); each rank computes volume rows [h0,h1) from the haloed
    crop.  Returns (best [H,W], disparity [H,W], camera_grad [H,W] | None), identical on every rank."""
    from . import functional as F
    rank, world = _world(group)
    H = camera.shape[0]
    band = row_band(H, kernel_size, rank, world)
    cam_c = crop_rows_with_halo(camera, band)
    proj_c = crop_rows_with_halo(projector, band)
    extra = {"prepared": F.prepare_backward(cam_c, proj_c, kernel_size, D)} \
        if cost_volume_grad_fn is not None and cam_c.is_cuda else {}
    cost, best, disp = F.forward(cam_c, proj_c, D, kernel_size, want_cost=cost_volume_grad_fn is not None,
                                 want_wta=True)
    grad = None
    if cost_volume_grad_fn is not None:
        # gradient of the OWNED rows only, handed over as it is: custma_backward_rows treats the halo rows of the crop as
        # zero, so no volume-sized zero fill or copy happens here
        g = cost_volume_grad_fn(band_rows_of(cost, band), band)
        grad = F.backward(g.contiguous(), cam_c, proj_c, kernel_size, D, rows=band_gradient_mask_rows(band), **extra)
        grad = assemble_row_band_gradient(grad, H, kernel_size, group)
    best = all_gather_row_bands(band_rows_of(best, band), H, kernel_size, group)
    disp = all_gather_row_bands(band_rows_of(disp, band), H, kernel_size, group)
    return best, disp, grad
