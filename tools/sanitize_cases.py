#!/usr/bin/env python
"""Small cases that reach every kernel body (all-valid / masked / scalar, fallback, finalize) and every entry point on
shapes with awkward remainders.  Written for compute-sanitizer memcheck (round 1 ran it clean); the tool is closed on
the round-2 GPU pool, so here it only checks that every call runs and returns finite results."""
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

# the calls added in round 2: fused mask, gradient row window, prepared backward / forward, fused head, projector gradient,
# 8-bit ingestion and the host entry points (both pipeline shapes)
from custereomatching_b200 import binding  # noqa: E402

for B, H, W, D, k in [(2, 37, 131, 64, 5), (1, 30, 97, 100, 3)]:
    shape = (B, H, W) if B > 1 else (H, W)
    cam, proj = torch.rand(*shape, device="cuda"), torch.rand(*shape, device="cuda")
    cb.wta_masked(cam, proj, D, k)
    g = torch.randn(*shape, D, device="cuda")
    prep = cb.prepare_backward(cam, proj, k, D)
    cb.backward(g, cam, proj, k, D, prepared=prep)
    cb.backward(g[..., 5:20, :, :].contiguous(), cam, proj, k, D, rows=(5, 20))
    cb.backward_projector(g, cam, proj, k, D)
    cam_r = cam.clone().requires_grad_(True)
    cb.soft_disparity(cam_r, proj, D, k)[0].sum().backward()
    hc, hp = cam.reshape(B, H, W).cpu().pin_memory(), proj.reshape(B, H, W).cpu().pin_memory()
    hb, hi, hg = torch.empty(B, H, W).pin_memory(), torch.empty(B, H, W, dtype=torch.int32).pin_memory(), torch.empty(B, H, W).pin_memory()
    gg = g.reshape(B, H, W, D).contiguous()
    for shape_env in ("single", "slots"):
        os.environ["CUSTMA_HOST_PIPELINE"] = shape_env
        binding.host_release()
        t = [binding.host_submit(hc.data_ptr(), hp.data_ptr(), hb.data_ptr(), hi.data_ptr(), hg.data_ptr(), 0, gg.data_ptr(),
                                 B, H, W, D, k) for _ in range(3)]
        binding.host_wait(t[-1])
        binding.host_step(hc.data_ptr(), hp.data_ptr(), hb.data_ptr(), hi.data_ptr(), hg.data_ptr(), 0, gg.data_ptr(), B, H, W, D, k)
    binding.host_release()
    u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
    cb.ingest_u8(u8, 1)
    torch.cuda.synchronize()
    print("ok round-2 calls", B, H, W, D, k, flush=True)
print("done")
