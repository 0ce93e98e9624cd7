#!/usr/bin/env python
"""Times the fused disparity head (forward, backward) against the unfused route (volume + torch soft-argmax, backward
through the volume) on P KITTI-size pairs: python tools/run_head.py [--pairs 8] [--iters 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import custereomatching_b200 as cb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
P, H, W, D, k = a.pairs, 375, 1242, 192, 5
g = torch.Generator(device="cuda").manual_seed(0)
cam = torch.rand(P, H, W, device="cuda", generator=g)
proj = torch.rand(P, H, W, device="cuda", generator=g)
gd = torch.randn(P, H, W, device="cuda", generator=g)
cells = P * H * W * D


def timed(fn, name):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{name:58s} {ms:8.3f} ms  {cells / ms / 1e6:8.1f} Gcell/s", flush=True)


def fused_fwd():
    with torch.no_grad():
        cb.soft_disparity(cam, proj, D, k)


def fused_fwd_bwd():
    c = cam.detach().requires_grad_(True)
    s, *_ = cb.soft_disparity(c, proj, D, k)
    s.backward(gd)


valid = (torch.arange(W, device="cuda")[:, None] - torch.arange(D, device="cuda")[None, :]) >= 0
sidx = torch.arange(D, device="cuda", dtype=torch.float32)


def unfused(backward):
    c = cam.detach().requires_grad_(backward)
    vol = cb.cost_volume(c, proj, D, k)
    logits = torch.where(valid, vol * 50.0, torch.full_like(vol, float("-inf")))
    w = torch.softmax(logits, dim=-1)
    soft = (w * sidx).sum(-1)
    best = logits.detach().max(dim=-1).values / 50.0
    out = soft * (best > 0.6).float()
    if backward:
        out.backward(gd)


timed(fused_fwd, "fused head forward (no volume)")
timed(fused_fwd_bwd, "fused head forward + backward (no volume)")
timed(lambda: unfused(False), "volume + torch soft-argmax + mask, forward")
timed(lambda: unfused(True), "volume + torch soft-argmax + mask, forward + backward")

# projector gradient: the mirrored sliding-window route against the direct kernels
gvol = torch.randn(P, H, W, D, device="cuda", generator=g)
timed(lambda: cb.backward_projector(gvol, cam, proj, k, D), "projector gradient, mirrored sliding-window route")
timed(lambda: cb.backward_projector(gvol, cam, proj, k, D, flags=cb.FLAG_DIRECT), "projector gradient, direct kernels")
timed(lambda: cb.backward(gvol, cam, proj, k, D), "camera gradient (for comparison)")
