#!/usr/bin/env python
"""One-off fuzz of the tensor-core kernels (CUSTMA_FLAG_TENSOR) against the direct kernels over random shapes:
python tools/fuzz_tensor.py [cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import custereomatching_b200 as cb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
worst_c = worst_g = 0.0
for case in range(n):
    k = int(rng.choice([3, 5, 5]))
    B = int(rng.choice([1, 1, 2, 3]))
    H = int(rng.randint(1, 90))
    W = int(rng.randint(1, 700))
    D = 4 * int(rng.randint(1, 136))                      # 4 .. 540
    if B * H * W * D > 60e6:
        H = max(1, int(60e6 / (B * W * D)))
    shape = (B, H, W) if B > 1 else (H, W)
    kind = int(rng.randint(3))
    if kind == 0:
        cam, proj = rng.rand(*shape), rng.rand(*shape)
    elif kind == 1:
        cam, proj = 0.3 + 0.4 * rng.rand(*shape), 0.5 + 0.1 * rng.rand(*shape)
    else:
        xx = np.broadcast_to(np.arange(W, dtype=np.float64), shape)
        cam = 0.5 + 0.3 * np.sin(xx * 0.02) + 0.02 * rng.rand(*shape)
        proj = 0.5 + 0.3 * np.sin((xx + 17) * 0.02) + 0.02 * rng.rand(*shape)
    cam = torch.from_numpy(np.ascontiguousarray(cam, np.float32)).cuda()
    proj = torch.from_numpy(np.ascontiguousarray(proj, np.float32)).cuda()
    tag = f"case {case}: B={B} H={H} W={W} D={D} k={k} kind={kind}"
    c0, b0, i0 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=cb.FLAG_DIRECT)
    c1, b1, i1 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
    ec = float((c1 - c0).abs().max())
    tb, ti = torch.flip(c1, dims=[-1]).max(dim=-1)
    ok_wta = bool(torch.equal(b1, tb) and torch.equal(i1.long(), (D - 1) - ti))
    g = torch.from_numpy(rng.randn(*(shape + (D,))).astype(np.float32)).cuda()
    g0 = cb.backward(g, cam, proj, k, D, flags=cb.FLAG_DIRECT)
    g1 = cb.backward(g, cam, proj, k, D, flags=cb.FLAG_TENSOR)
    scale = float(g0.abs().max()) + 1e-30
    eg = float((g1 - g0).abs().max()) / scale
    worst_c, worst_g = max(worst_c, ec), max(worst_g, eg)
    bad = ec > 1e-5 or eg > 3e-5 or not ok_wta or not torch.isfinite(g1).all()
    if eg > 3e-5 and B * H * W * D < 3e6:
        # ill-conditioned gradients: compare both kernels with the fp64 oracle (tests do the same, conftest.py)
        from oracle import zncc_oracle as zo
        cs, ps, gs = (t.cpu().numpy().reshape((-1, H, W) + t.shape[len(shape):]) for t in (cam, proj, g))
        truth = np.stack([zo.camera_grad_banded_autograd(cs[b], ps[b], gs[b], D, k).numpy() for b in range(cs.shape[0])]).reshape(g0.shape)
        truth = torch.from_numpy(truth).cuda()
        e_tc, e_dir = float((g1.double() - truth).abs().max()) / scale, float((g0.double() - truth).abs().max()) / scale
        print(f"     vs fp64: tensor-core {e_tc:.2e}, direct {e_dir:.2e} of scale", flush=True)
        bad = bad and not (e_tc <= e_dir + 1e-5)
    if bad or case % 20 == 0:
        print(("BAD  " if bad else "ok   ") + tag + f"  cost err {ec:.2e}  grad err {eg:.2e} of {scale:.2e}  wta {'ok' if ok_wta else 'MISMATCH'}", flush=True)
print(f"{n} cases: worst cost error {worst_c:.2e}, worst gradient error {worst_g:.2e} of scale")
