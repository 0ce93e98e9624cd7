"""CPU oracle for the ZNCC cost-volume hot path of lzhnb/CuStereoMatching.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this
module; the product (`custereomatching_b200/`, `custma/`) never does and has no CPU fallback.

Parity status: PINNED.  The reference ships no golden vectors, so the pins were
produced by running the unmodified reference CUDA extension on a B200 (tests/golden/make_golden_from_reference.py)
and are committed as tests/golden/*.npz; tests/test_oracle_golden.py checks every function below against them.

What is restated (all file:line relative to /root/reference):
  * patch extraction with zero padding         examples/verify.py:18-28  (F.pad + unfold)
  * centring, correlation, normalisation       examples/verify.py:107-120 (mean, bmm, EX2, EY2, eps = 1e-8 at :112)
  * conventions where the CUDA kernel rules    custma/src/stereo_matching_kernel.cu
      - window offsets  i - k/2, j - k/2  (integer division; asymmetric for even k)     :44-46
      - out-of-image pixels read as 0 and still count in the mean's divisor k*k          :6-12, :53-54
      - axis order [h, w_camera, d_projector], d = absolute projector column            :35-37
      - cost = (exy + eps) / sqrt(ex2 * ey2 + eps)                                       :71
      - camera-only gradient, out-of-image patch gradients dropped                      :135-151, :177
        custma/stereo_matching_wrapper.py:33
  * WTA / mask / soft-argmax (example level)   examples/verify.py:31-39, :72-74; examples/test.py:78-86

Everything runs in torch on the CPU, fp64 by default (fp32 on request), so it doubles as the "pure-PyTorch CPU
reimplementation" that BASELINE.json's north_star names as the reported CPU baseline.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-8                 # stereo_matching_kernel.cu:4 ; examples/verify.py:112
INVALID_COST = -2.0        # banded extension: value of cells whose projector column w - s falls left of the image
MASK_THRESHOLD = 0.6       # examples/verify.py:13
SOFTARGMAX_BETA = 50.0     # examples/verify.py:12


def _as_tensor(x, dtype):
    return torch.as_tensor(x).to(dtype=dtype, device="cpu")


def extract_patches(img: torch.Tensor, k: int) -> torch.Tensor:
    """[H,W] -> [H,W,k*k]; element (i,j) of patch (h,w) is img[h+i-k//2, w+j-k//2], 0 outside the image.

    examples/verify.py:18-28 pads int((k-1)/2) on every side, which for even k yields a smaller output; the CUDA
    kernel (stereo_matching_kernel.cu:44-46, query_ij :6-12) defines the even-k behaviour, so padding here is
    k//2 before and k-1-k//2 after (identical to verify.py for odd k).
    """
    r = k // 2
    x = F.pad(img[None, None], (r, k - 1 - r, r, k - 1 - r), mode="constant", value=0.0)
    x = x.unfold(2, k, 1).unfold(3, k, 1)          # [1,1,H,W,k(i),k(j)]
    H, W = img.shape
    return x[0, 0].reshape(H, W, k * k)


def centred_patches(img: torch.Tensor, k: int):
    """Returns (patches - mean, mean).  verify.py:107-113; the mean divides by k*k including padded zeros (:53-54)."""
    p = extract_patches(img, k)
    m = p.sum(-1, keepdim=True) / float(k * k)
    return p - m, m[..., 0]


def cost_volume_full(camera, projector, kernel_size: int, dtype=torch.float64) -> torch.Tensor:
    """Reference-shaped ZNCC volume [H, W, W] (stereo_matching_kernel.cu:17-72 == examples/verify.py:107-120)."""
    cam = _as_tensor(camera, dtype)
    proj = _as_tensor(projector, dtype)
    cc, _ = centred_patches(cam, kernel_size)
    pc, _ = centred_patches(proj, kernel_size)
    exy = torch.bmm(cc, pc.transpose(1, 2))                       # verify.py:116
    ex2 = (cc * cc).sum(-1)[:, :, None]                           # verify.py:118
    ey2 = (pc * pc).sum(-1)[:, None, :]                           # verify.py:119
    return (exy + EPS) / torch.sqrt(ex2 * ey2 + EPS)              # verify.py:120 ; kernel.cu:71


def camera_grad_full_autograd(camera, projector, cost_volume_grad, kernel_size: int, dtype=torch.float64):
    """d(sum(cost*g))/d(camera) through autograd, exactly what examples/verify.py:122-123 does."""
    cam = _as_tensor(camera, dtype).clone().requires_grad_(True)
    cv = cost_volume_full(cam, projector, kernel_size, dtype)
    cv.backward(_as_tensor(cost_volume_grad, dtype))
    return cam.grad.detach()


def camera_grad_full_kernel_formula(camera, projector, cost_volume_grad, kernel_size: int, dtype=torch.float64):
    """The reference kernels' own gradient formula, restated literally.

    get_patches_grad_kernel (stereo_matching_kernel.cu:75-152):
        patch_grad[h,w,i,j] = sum_d g[h,w,d] * (proj_c/den - ey2 * cam_c * (exy+eps) / den^3),  den = sqrt(ex2*ey2+eps)
    patches_grad_to_image_kernel (:155-179): overlap-add into the image, out-of-image targets dropped (:177).
    The Jacobian of the mean subtraction is skipped by the reference; it contributes exactly zero.
    """
    k = kernel_size
    cam = _as_tensor(camera, dtype)
    proj = _as_tensor(projector, dtype)
    g = _as_tensor(cost_volume_grad, dtype)
    H, W = cam.shape
    cc, _ = centred_patches(cam, k)
    pc, _ = centred_patches(proj, k)
    exy = torch.bmm(cc, pc.transpose(1, 2))
    ex2 = (cc * cc).sum(-1)[:, :, None]
    ey2 = (pc * pc).sum(-1)[:, None, :]
    den = torch.sqrt(ex2 * ey2 + EPS)
    a = g / den                                                    # coefficient of proj_c   (:135, :145)
    b = g * ey2 * (exy + EPS) / den ** 3                           # coefficient of cam_c    (:135, :147)
    patch_grad = torch.bmm(a, pc) - b.sum(-1, keepdim=True) * cc   # [H,W,k*k]
    r = k // 2
    out = torch.zeros(H + k - 1, W + k - 1, dtype=dtype)
    pg = patch_grad.reshape(H, W, k, k)
    for i in range(k):
        for j in range(k):
            out[i:i + H, j:j + W] += pg[:, :, i, j]                # target (h+i-r, w+j-r) in padded coordinates
    return out[r:r + H, r:r + W].contiguous()


# ----------------------------------------------------------------------------------------------------------------
# Banded extension: band[h, w, s] = full[h, w, w - s], s in [0, D); invalid (w - s < 0) cells hold INVALID_COST,
# are excluded from WTA and receive/propagate zero gradient.  Convention from examples/test.py:83
# (disparity = column - correspondence).
# ----------------------------------------------------------------------------------------------------------------
def band_from_full(full: torch.Tensor, D: int, invalid=INVALID_COST) -> torch.Tensor:
    H, W, _ = full.shape
    w = torch.arange(W)[:, None]
    s = torch.arange(D)[None, :]
    d = w - s
    valid = d >= 0
    band = full[:, w.expand(W, D), d.clamp(min=0)]                 # [H,W,D]
    return torch.where(valid[None], band, torch.full_like(band, invalid))


def full_from_band_grad(band_grad: torch.Tensor, W: int) -> torch.Tensor:
    """Scatter an upstream gradient on the band back to reference-shaped [H,W,W] (zeros elsewhere)."""
    H, W_, D = band_grad.shape
    assert W_ == W
    full = torch.zeros(H, W, W, dtype=band_grad.dtype)
    w = torch.arange(W)[:, None].expand(W, D)
    d = w - torch.arange(D)[None, :]
    valid = d >= 0
    full[:, w[valid], d[valid]] = band_grad[:, valid]
    return full


def cost_volume_banded(camera, projector, D: int, kernel_size: int, dtype=torch.float64, invalid=INVALID_COST):
    """[H, W, D] banded volume computed diagonal by diagonal (memory O(H*W*D), never forms [H,W,W])."""
    cam = _as_tensor(camera, dtype)
    proj = _as_tensor(projector, dtype)
    H, W = cam.shape
    cc, _ = centred_patches(cam, kernel_size)
    pc, _ = centred_patches(proj, kernel_size)
    ex2 = (cc * cc).sum(-1)
    ey2 = (pc * pc).sum(-1)
    cols = []
    for s in range(D):
        col = torch.full((H, W), invalid, dtype=dtype)
        if s < W:
            exy = (cc[:, s:, :] * pc[:, :W - s, :]).sum(-1)
            col = torch.cat([col[:, :s], (exy + EPS) / torch.sqrt(ex2[:, s:] * ey2[:, :W - s] + EPS)], dim=1)
        cols.append(col)
    return torch.stack(cols, dim=-1)


def camera_grad_banded_autograd(camera, projector, band_grad, D: int, kernel_size: int, dtype=torch.float64):
    cam = _as_tensor(camera, dtype).clone().requires_grad_(True)
    band = cost_volume_banded(cam, projector, D, kernel_size, dtype)
    band.backward(_as_tensor(band_grad, dtype))
    return cam.grad.detach()


def projector_grad_autograd(camera, projector, cost_volume_grad, D: int, kernel_size: int, dtype=torch.float64):
    """Gradient of sum(cost * cost_volume_grad) with respect to the PROJECTOR image by autograd (the oracle of
    custma_backward_projector; the reference itself returns None for it, custma/stereo_matching_wrapper.py:33).
    D > 0: banded volume (invalid cells carry no gradient); D == 0: the [H,W,W] volume."""
    proj = _as_tensor(projector, dtype).clone().requires_grad_(True)
    g = _as_tensor(cost_volume_grad, dtype)
    if D > 0:
        vol = cost_volume_banded(camera, proj, D, kernel_size, dtype=dtype, invalid=0.0)
        W = proj.shape[1]
        valid = (torch.arange(W)[:, None] - torch.arange(D)[None, :]) >= 0
        (vol * g * valid[None].to(dtype)).sum().backward()
    else:
        (cost_volume_full(camera, proj, kernel_size, dtype=dtype) * g).sum().backward()
    return proj.grad.detach()


# ----------------------------------------------------------------------------------------------------------------
# Winner-take-all, confidence mask, soft-argmax (examples/verify.py:31-39, 72-74; examples/test.py:78-86)
# ----------------------------------------------------------------------------------------------------------------
def wta_full(full: torch.Tensor):
    """(best [H,W], correspondence [H,W] int64 = first maximal projector column, disparity = w - correspondence)."""
    best, corr = torch.max(full, dim=-1)                           # torch.max returns the first maximal index
    W = full.shape[1]
    disparity = torch.arange(W)[None, :] - corr                    # examples/test.py:83
    return best, corr, disparity


def wta_banded(band: torch.Tensor):
    """(best, disparity).  Ties resolve to the LOWEST projector column == LARGEST disparity."""
    D = band.shape[-1]
    flipped = torch.flip(band, dims=[-1])                          # ascending projector column
    best, idx = torch.max(flipped, dim=-1)
    disparity = (D - 1) - idx
    return best, disparity


def top2_gap(vol: torch.Tensor) -> torch.Tensor:
    """best minus runner-up along the last axis: disparities are compared only where this exceeds the tolerance."""
    t = vol.topk(2, dim=-1).values
    return t[..., 0] - t[..., 1]


def confidence_mask(best: torch.Tensor, threshold: float = MASK_THRESHOLD) -> torch.Tensor:
    return (best > threshold).to(best.dtype)                       # examples/verify.py:74


def soft_argmax(vol: torch.Tensor, beta: float = SOFTARGMAX_BETA) -> torch.Tensor:
    """examples/verify.py:31-39 on the last axis: sum_i softmax(beta * x)_i * i."""
    sm = torch.softmax(vol * beta, dim=-1)
    idx = torch.arange(vol.shape[-1], dtype=vol.dtype)
    return (sm * idx).sum(-1)


def soft_disparity_head(camera, projector, D: int, kernel_size: int, beta: float = SOFTARGMAX_BETA,
                        threshold: float = MASK_THRESHOLD, soft_grad=None, dtype=torch.float64):
    """The example-level disparity head on top of the volume (the oracle of custma_forward_head / custma_backward_head):
    soft_argmax of examples/verify.py:31-39 on the last axis, disparity = column - correspondence and the product with the
    confidence mask of examples/test.py:79-86 (mask = best > threshold, verify.py:72-74, detached as there).
    D > 0: banded volume, the axis is the disparity itself and cells with w - s < 0 take no part; D == 0: the reference's
    [H,W,W] volume, soft disparity = column - sum_d softmax * d.
    Returns (soft * mask, best, mask, weights' mean absolute deviation of s, camera_grad or None); the camera gradient is
    autograd's for the loss sum(soft * mask * soft_grad)."""
    cam = _as_tensor(camera, dtype).clone().requires_grad_(soft_grad is not None)
    proj = _as_tensor(projector, dtype)
    H, W = cam.shape
    if D > 0:
        vol = cost_volume_banded(cam, proj, D, kernel_size, dtype=dtype, invalid=float("-inf"))
        s = torch.arange(D, dtype=dtype)[None, None, :].expand(H, W, D)
    else:
        vol = cost_volume_full(cam, proj, kernel_size, dtype=dtype)
        s = torch.arange(W, dtype=dtype)[None, :, None] - torch.arange(W, dtype=dtype)[None, None, :]
        s = s.expand(H, W, W)
    w = torch.softmax(vol * beta, dim=-1)                          # verify.py:34 (exp(-inf) = 0 for missing cells)
    soft = (w * s).sum(-1)                                         # verify.py:36-38 / test.py:85
    best = vol.detach().max(dim=-1).values
    mask = (best > threshold).to(dtype)                            # verify.py:74
    out = soft * mask                                              # test.py:86
    mad = (w.detach() * (s - soft.detach()[..., None]).abs()).sum(-1)
    grad = None
    if soft_grad is not None:
        (out * _as_tensor(soft_grad, dtype)).sum().backward()
        grad = cam.grad.detach()
    return out.detach(), best, mask, mad, grad


# ----------------------------------------------------------------------------------------------------------------
# Timed CPU baseline (north_star: "a pure-PyTorch CPU reimplementation ... timed on the box's own host cores")
# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline_step_banded(camera: torch.Tensor, projector: torch.Tensor, band_grad: torch.Tensor, D: int, k: int):
    """One fwd + WTA + bwd pass in fp32 torch on the host; returns (best, disparity, camera_grad)."""
    cam = camera.detach().clone().requires_grad_(True)
    band = cost_volume_banded(cam, projector, D, k, dtype=torch.float32)
    best, disp = wta_banded(band.detach())
    band.backward(band_grad)
    return best, disp, cam.grad


def cpu_baseline_step_full(camera: torch.Tensor, projector: torch.Tensor, grad: torch.Tensor, k: int):
    cam = camera.detach().clone().requires_grad_(True)
    cv = cost_volume_full(cam, projector, k, dtype=torch.float32)
    best, corr, _ = wta_full(cv.detach())
    cv.backward(grad)
    return best, corr, cam.grad
