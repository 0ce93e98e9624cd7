#!/usr/bin/env python
"""custma_host_submit pipeline timing under different host waiting patterns (8 KITTI pairs)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custereomatching_b200 import binding  # noqa: E402

P, H, W, D, k = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 375, 1242, 192, 5
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
h_cam = torch.rand(P, H, W, generator=g).pin_memory()
h_proj = torch.rand(P, H, W, generator=g).pin_memory()
gin = torch.randn(P, H, W, D, device=dev)
hb = [torch.empty(P, H, W).pin_memory() for _ in range(2)]
hi = [torch.empty(P, H, W, dtype=torch.int32).pin_memory() for _ in range(2)]
hg = [torch.empty(P, H, W).pin_memory() for _ in range(2)]


def submit(i, with_grad=True):
    return binding.host_submit(h_cam.data_ptr(), h_proj.data_ptr(), hb[i & 1].data_ptr(), hi[i & 1].data_ptr(),
                               hg[i & 1].data_ptr() if with_grad else 0, 0, gin.data_ptr() if with_grad else 0, P, H, W, D, k, 0)


def loop(mode, steps=40, with_grad=True):
    for i in range(3):
        binding.host_wait(submit(i, with_grad))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = None
    tsub = 0.0
    for i in range(steps):
        a = time.perf_counter()
        t = submit(i, with_grad)
        tsub += time.perf_counter() - a
        if mode == "lag1" and prev is not None:
            binding.host_wait(prev)
        if mode == "sync":
            binding.host_wait(t)
        prev = t
    binding.host_wait(0)
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f"{mode:6s} grad={int(with_grad)}: {dt:.4f} ms/step   (host time inside submit: {tsub / steps * 1e3:.4f} ms/step)", flush=True)


if len(sys.argv) > 2:
    loop(sys.argv[2])
else:
    for m in ("lag1", "free", "sync"):
        loop(m)
    loop("lag1", with_grad=False)
    loop("free", with_grad=False)
