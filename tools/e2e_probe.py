#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-resident step?  Runs forward + backward of 8 KITTI
pairs on one stream and adds the host traffic of custma_host_submit piece by piece on side streams (no dependencies
between the copies and the kernels, so only contention can slow the kernels down)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custereomatching_b200 import binding  # noqa: E402

P, H, W, D, k = 8, 375, 1242, 192, 5
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
cam = torch.rand(P, H, W, generator=g).to(dev)
proj = torch.rand(P, H, W, generator=g).to(dev)
cost = torch.empty(P, H, W, D, device=dev)
gin = torch.randn(P, H, W, D, device=dev)
best = torch.empty(P, H, W, device=dev)
idx = torch.empty(P, H, W, dtype=torch.int32, device=dev)
grad = torch.empty(P, H, W, device=dev)
wsb = max(binding.forward_workspace_bytes(P, H, W, D, k, 0), binding.backward_workspace_bytes(P, H, W, D, k, 0))
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
h_in = [torch.rand(P, H, W).pin_memory() for _ in range(2)]
d_in = [torch.empty(P, H, W, device=dev) for _ in range(2)]
h_out = [torch.empty(P, H, W).pin_memory() for _ in range(3)]
main = torch.cuda.Stream()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, steps=40):
    sp = main.cuda_stream
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        if h2d:
            with torch.cuda.stream(s_in):
                for a, b in zip(d_in, h_in):
                    a.copy_(b, non_blocking=True)
        binding.forward(cam.data_ptr(), proj.data_ptr(), cost.data_ptr(), best.data_ptr(), idx.data_ptr(), P, H, W, D, k, 0,
                        ws.data_ptr(), wsb, sp)
        binding.backward(gin.data_ptr(), cam.data_ptr(), proj.data_ptr(), grad.data_ptr(), P, H, W, D, k, 0, ws.data_ptr(), wsb, sp)
        if d2h:
            with torch.cuda.stream(s_out):
                for a, b in zip(h_out, (best, idx.view(torch.float32), grad)):
                    a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for name, a, b in (("kernels only", 0, 0), ("+ H2D 30 MB", 1, 0), ("+ D2H 45 MB", 0, 1), ("+ both", 1, 1), ("kernels only", 0, 0)):
    run(a, b, 5)
    print(f"{name:14s} {run(a, b):.4f} ms/step", flush=True)
