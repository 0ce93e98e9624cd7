#!/usr/bin/env python
"""One-off fuzz of the default path (sliding-window kernels, per-tile fallback, tensor-core hand-over) against the direct
kernels over random shapes, banded and reference-shaped: python tools/fuzz_default.py [cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import custereomatching_b200 as cb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
worst_c = worst_g = 0.0
nbad = 0
for case in range(n):
    k = int(rng.choice([3, 5, 5, 5, 7]))
    B = int(rng.choice([1, 1, 2, 3]))
    H = int(rng.randint(1, 200))
    W = int(rng.randint(1, 500))
    D = int(rng.choice([0, 1, 4, 17, 32, 64, 100, 128, 192, 200, 256, 320]))
    if D == 0 and W > 220:
        W = 220
    C = D if D > 0 else W
    if B * H * W * C > 50e6:
        H = max(1, int(50e6 / (B * W * C)))
    shape = (B, H, W) if B > 1 else (H, W)
    kind = int(rng.randint(4))
    if kind == 0:
        cam, proj = rng.rand(*shape), rng.rand(*shape)
    elif kind == 1:
        cam, proj = 0.3 + 0.4 * rng.rand(*shape), 0.5 + 0.1 * rng.rand(*shape)
    elif kind == 2:
        xx = np.broadcast_to(np.arange(W, dtype=np.float64), shape)
        cam = 0.5 + 0.3 * np.sin(xx * 0.02) + 0.02 * rng.rand(*shape)
        proj = 0.5 + 0.3 * np.sin((xx + 17) * 0.02) + 0.02 * rng.rand(*shape)
    else:   # half the frame black: some tiles flagged, most not
        cam, proj = rng.rand(*shape), rng.rand(*shape)
        cam[..., : W // 3] = 0.0
    cam = torch.from_numpy(np.ascontiguousarray(cam, np.float32)).cuda()
    proj = torch.from_numpy(np.ascontiguousarray(proj, np.float32)).cuda()
    tag = f"case {case}: B={B} H={H} W={W} D={D} k={k} kind={kind}"
    c0, b0, i0 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=cb.FLAG_DIRECT)
    c1, b1, i1 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True)
    ec = float(((c1 - c0).abs() / c0.abs().clamp(min=1.0)).max())
    if D > 0:
        tb, ti = torch.flip(c1, dims=[-1]).max(dim=-1)
        ok_wta = bool(torch.equal(b1, tb) and torch.equal(i1.long(), (D - 1) - ti))
    else:
        tb, ti = c1.max(dim=-1)
        ok_wta = bool(torch.equal(b1, tb) and torch.equal(i1.long(), ti))
    g = torch.from_numpy(rng.randn(*(shape + (C,))).astype(np.float32)).cuda()
    g0 = cb.backward(g, cam, proj, k, D, flags=cb.FLAG_DIRECT)
    g1 = cb.backward(g, cam, proj, k, D)
    scale = float(g0.abs().max()) + 1e-30
    eg = float((g1 - g0).abs().max()) / scale
    if eg > worst_g:
        print(f'     new worst gradient error {eg:.2e} at {tag}', flush=True)
    if ec > worst_c:
        print(f'     new worst cost error {ec:.2e} at {tag}', flush=True)
    worst_c, worst_g = max(worst_c, ec), max(worst_g, eg)
    bad = ec > 1e-5 or eg > 3e-5 or not ok_wta or not torch.isfinite(g1).all()
    nbad += bad
    if bad or case % 25 == 0:
        print(("BAD  " if bad else "ok   ") + tag + f"  cost err {ec:.2e}  grad err {eg:.2e} of {scale:.2e}  wta {'ok' if ok_wta else 'MISMATCH'}", flush=True)
print(f"{n} cases, {nbad} bad: worst cost error {worst_c:.2e}, worst gradient error {worst_g:.2e} of scale")
