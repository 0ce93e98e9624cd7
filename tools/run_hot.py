#!/usr/bin/env python
"""Runs the hot path a few times on one GPU (for ncu / quick timing): python tools/run_hot.py [--phase fwd|bwd|both]
[--workload kitti|cfg2|cfg3|cfg1] [--pairs P] [--iters N] [--direct] [--no-wta] [--no-cost]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from custereomatching_b200 import binding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--phase", default="both")
ap.add_argument("--workload", default="kitti")
ap.add_argument("--pairs", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--direct", action="store_true")
ap.add_argument("--tensor", action="store_true", help="force the tensor-core forward")
ap.add_argument("--no-wta", action="store_true")
ap.add_argument("--no-cost", action="store_true")
ap.add_argument("--full", action="store_true", help="reference-shaped [H,W,W] volume (D = 0)")
a = ap.parse_args()
H, W, D, k, P0, _ = WORKLOADS[a.workload]
if a.full:
    D = 0
P = a.pairs or P0
C = D if D > 0 else W
flags = binding.FLAG_DIRECT if a.direct else binding.FLAG_TENSOR if a.tensor else 0
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = torch.Generator().manual_seed(0)
cam = torch.rand(P, H, W, generator=g).to(dev)
proj = torch.rand(P, H, W, generator=g).to(dev)
cost = torch.empty(P, H, W, C, device=dev)
best = torch.empty(P, H, W, device=dev)
idx = torch.empty(P, H, W, dtype=torch.int32, device=dev)
grad = torch.empty(P, H, W, device=dev)
gin = torch.randn(P, H, W, C, device=dev) if a.phase != "fwd" else None
wsb = max(binding.forward_workspace_bytes(P, H, W, D, k, flags), binding.backward_workspace_bytes(P, H, W, D, k, flags))
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
cells = P * H * W * C


def fwd():
    binding.forward(cam.data_ptr(), proj.data_ptr(), 0 if a.no_cost else cost.data_ptr(),
                    0 if a.no_wta else best.data_ptr(), 0 if a.no_wta else idx.data_ptr(), P, H, W, D, k, flags,
                    ws.data_ptr(), wsb, s)


def bwd():
    binding.backward(gin.data_ptr(), cam.data_ptr(), proj.data_ptr(), grad.data_ptr(), P, H, W, D, k, flags,
                     ws.data_ptr(), wsb, s)


for name, fn in (("fwd", fwd), ("bwd", bwd)):
    if a.phase not in (name, "both"):
        continue
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{name}: {ms:.4f} ms/iter  {cells / ms / 1e6:.1f} Gcell/s  {4 * cells / ms / 1e6:.1f} GB/s algorithmic "
          f"({4 * cells / ms / 1e6 / 6550.1 * 100:.1f}% of 6550 GB/s)  [P={P} H={H} W={W} D={D} k={k} flags={flags}]", flush=True)
