#!/usr/bin/env python
"""Small forward/backward cases for compute-sanitizer (memcheck): every kernel body (all-valid / masked / scalar,
fallback, finalize) on shapes with awkward remainders."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import custereomatching_b200 as cb  # noqa: E402

torch.manual_seed(0)
CASES = [  # B, H, W, D, k
    (1, 70, 330, 192, 5), (2, 37, 131, 64, 5), (1, 19, 75, 100, 5), (1, 33, 97, 0, 5), (1, 12, 40, 0, 5),
    (1, 66, 520, 256, 5), (1, 9, 21, 32, 5), (1, 140, 200, 128, 5), (1, 20, 24, 0, 15), (1, 13, 17, 7, 3),
]
for B, H, W, D, k in CASES:
    shape = (B, H, W) if B > 1 else (H, W)
    cam = torch.rand(*shape, device="cuda")
    proj = torch.rand(*shape, device="cuda")
    for flat in (False, True):
        c = cam * 0.001 + 0.8 if flat else cam          # the flat variant sends tiles to the fallback kernels
        cost, best, idx = cb.forward(c, proj, D, k, want_cost=True, want_wta=True)
        g = torch.randn_like(cost)
        grad = cb.backward(g, c, proj, k, D)
        torch.cuda.synchronize()
        assert torch.isfinite(cost).all() and torch.isfinite(grad).all()
    print("ok", B, H, W, D, k, flush=True)
print("done")
