#!/usr/bin/env python
"""How the conditioning verdict reacts to the input distribution: forward / backward time of one KITTI-size pair for
several synthetic image families (slow = tiles handed to the direct-arithmetic fallback)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import custereomatching_b200 as cb  # noqa: E402

H, W, D, k = 375, 1242, 192, 5
g = torch.Generator(device="cuda").manual_seed(0)
yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
rnd = lambda: torch.rand(H, W, device="cuda", generator=g)


def shifted(p, noise):
    c = torch.zeros_like(p)
    c[:, 40:] = p[:, :-40]
    return (c + noise * torch.randn(H, W, device="cuda", generator=g)).contiguous()


proj = rnd()
ramp = 0.2 + 0.6 * xx / W
cases = {
    "uniform random pair": (rnd(), proj),
    "camera = projector shifted by 40 + 1% noise": (shifted(proj, 0.01), proj),
    "random speckle, contrast 0.5 on a 0.25 pedestal": (0.25 + 0.5 * rnd(), 0.25 + 0.5 * proj),
    "horizontal brightness ramp 0.2..0.8 + texture std 0.06": (ramp + 0.2 * (rnd() - 0.5), ramp + 0.2 * (proj - 0.5)),
    "ramp + texture std 0.015 (low texture)": (ramp + 0.05 * (rnd() - 0.5), ramp + 0.05 * (proj - 0.5)),
    "flat 0.8 + 0.2% noise (no texture)": (0.8 + 0.002 * rnd(), 0.8 + 0.002 * proj),
}


def pink(alpha, seed):
    """1/f^alpha noise scaled to [0, 1]: the spectrum of natural images is close to alpha = 1"""
    gg = torch.Generator(device="cuda").manual_seed(seed)
    spec = torch.fft.rfft2(torch.randn(H, W + D, device="cuda", generator=gg))
    fy = torch.fft.fftfreq(H, device="cuda")[:, None]
    fx = torch.fft.rfftfreq(W + D, device="cuda")[None, :]
    f = torch.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    img = torch.fft.irfft2(spec / f ** alpha, s=(H, W + D))
    img = (img - img.min()) / (img.max() - img.min())
    return img


for alpha in (1.0, 1.5, 2.0):
    scene = pink(alpha, 3)
    # a rectified pair of the same scene: constant disparity 40 plus 1 % sensor noise in each view
    left = (scene[:, D - 40:D - 40 + W] + 0.01 * torch.randn(H, W, device="cuda", generator=g)).contiguous()
    right = (scene[:, D:D + W] + 0.01 * torch.randn(H, W, device="cuda", generator=g)).contiguous()
    cases[f"natural-like pair: 1/f^{alpha} scene, disparity 40, 1% noise"] = (right, left)
blocks = torch.rand(H // 25 + 1, W // 54 + 1, device="cuda", generator=g).repeat_interleave(25, 0).repeat_interleave(54, 1)[:H, :W]
cases["piecewise-constant blocks + 1% noise (cartoon)"] = (blocks + 0.01 * torch.randn(H, W, device="cuda", generator=g),
                                                          blocks.roll(-20, 1) + 0.01 * torch.randn(H, W, device="cuda", generator=g))
gin = torch.randn(H, W, D, device="cuda", generator=g)
for name, (cam, prj) in cases.items():
    cam, prj = cam.contiguous(), prj.contiguous()
    out = []
    for fn in (lambda: cb.forward(cam, prj, D, k, want_cost=True, want_wta=True), lambda: cb.backward(gin, cam, prj, k, D)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 3)
    print(f"{name:58s} fwd {out[0]:8.3f} ms   bwd {out[1]:8.3f} ms", flush=True)
