#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference CUDA extension.

The reference (lzhnb/CuStereoMatching) ships no tests, golden vectors or CPU kernel (SURVEY.md section 4 and
section 8c), so the pins for the oracle are produced here: seeded inputs -> reference `custma.stereo_matching`
forward (custma/stereo_matching_wrapper.py:9-21) and backward (:24-33) on a B200 -> small .npz files.

Run on a GPU box (the driver-prebuilt reference extension lives in baseline/_ref, git-ignored, travels with gpurun):

    gpurun -- python tests/golden/make_golden_from_reference.py

It writes gpurun_out/golden/*.npz; those files are then copied verbatim into tests/golden/ and committed.
Nothing in tests/ or bench.py imports the reference at run time - only these fixtures.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import torch  # noqa: E402
import custma as ref_custma  # noqa: E402  (the reference package, NOT this repo's drop-in)

assert "baseline" in os.path.abspath(ref_custma.__file__), ref_custma.__file__
OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def make_inputs(kind, H, W, seed):
    rng = np.random.RandomState(seed)
    if kind == "rand":
        cam = rng.rand(H, W).astype(np.float32)
        proj = rng.rand(H, W).astype(np.float32)
    elif kind == "smooth":  # low contrast: stresses cancellation in one-pass formulations
        xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
        cam = (0.8 + 0.01 * np.sin(xx / 7) + 0.002 * rng.rand(H, W)).astype(np.float32)
        proj = (0.8 + 0.01 * np.sin((xx + 3) / 7) + 0.002 * rng.rand(H, W)).astype(np.float32)
    elif kind == "shifted":  # camera = projector shifted by a known disparity + noise: WTA has a ground truth
        proj = rng.rand(H, W).astype(np.float32)
        cam = np.zeros_like(proj)
        d0 = 3
        cam[:, d0:] = proj[:, :-d0]
        cam += (0.01 * rng.randn(H, W)).astype(np.float32)
    elif kind == "const":  # analytic KAT: constant camera image -> cost == eps/sqrt(eps) == 1e-4
        cam = np.full((H, W), 0.5, np.float32)
        proj = rng.rand(H, W).astype(np.float32)
    elif kind == "edges":  # flat regions + a step edge + texture
        cam = np.where(np.arange(W)[None, :] < W // 2, 0.2, 0.8).astype(np.float32) * np.ones((H, 1), np.float32)
        cam = (cam + 0.002 * rng.rand(H, W)).astype(np.float32)
        proj = rng.rand(H, W).astype(np.float32)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(cam), np.ascontiguousarray(proj)


def run_ref(cam, proj, k, seed_g):
    camg = torch.from_numpy(cam).cuda().requires_grad_(True)
    projg = torch.from_numpy(proj).cuda()
    cv = ref_custma.stereo_matching(camg, projg, 64, k)
    torch.cuda.synchronize()
    g = np.random.RandomState(seed_g).randn(*cv.shape).astype(np.float32)
    cv.backward(torch.from_numpy(g).cuda())
    torch.cuda.synchronize()
    return cv.detach().cpu().numpy(), g, camg.grad.detach().cpu().numpy()


SMALL = [  # name, kind, H, W, k
    ("rand_k5_12x20", "rand", 12, 20, 5),
    ("rand_k3_9x33", "rand", 9, 33, 3),
    ("rand_k4_10x16", "rand", 10, 16, 4),      # even k: only the CUDA kernel convention is defined
    ("rand_k7_13x17", "rand", 13, 17, 7),
    ("rand_k15_20x24", "rand", 20, 24, 15),    # verify.py's kernel size; window mostly padding
    ("rand_k1_6x8", "rand", 6, 8, 1),
    ("smooth_k5_12x40", "smooth", 12, 40, 5),
    ("smooth_k15_18x40", "smooth", 18, 40, 15),
    ("shifted_k5_16x32", "shifted", 16, 32, 5),
    ("const_k5_8x12", "const", 8, 12, 5),
    ("edges_k5_12x32", "edges", 12, 32, 5),
]

manifest = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cases": []}
for name, kind, H, W, k in SMALL:
    cam, proj = make_inputs(kind, H, W, seed=1234)
    cv, g, grad = run_ref(cam, proj, k, seed_g=4321)
    cv2, _, grad2 = run_ref(cam, proj, k, seed_g=4321)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), camera=cam, projector=proj, kernel_size=np.int32(k),
                        cost_volume=cv, cost_volume_grad=g, camera_grad=grad)
    manifest["cases"].append({"name": name, "kind": kind, "H": H, "W": W, "k": k,
                              "fwd_run2run_bitexact": bool(np.array_equal(cv, cv2)),
                              "bwd_run2run_maxabs": float(np.abs(grad - grad2).max()),
                              "cv_min": float(cv.min()), "cv_max": float(cv.max()),
                              "grad_absmax": float(np.abs(grad).max())})
    print(json.dumps(manifest["cases"][-1]), flush=True)

# cfg1 (BASELINE.json configs[0]): 320x240, 5x5. The full [240,320,320] volume is 98 MB, so the fixture keeps the
# WTA (max / first argmax over the projector axis), the camera gradient and a strided subsample of the volume.
H, W, k = 240, 320, 5
cam, proj = make_inputs("rand", H, W, seed=7)
camg = torch.from_numpy(cam).cuda().requires_grad_(True)
projg = torch.from_numpy(proj).cuda()
cv = ref_custma.stereo_matching(camg, projg, 64, k)
gen = torch.Generator(device="cpu").manual_seed(11)
g_small = torch.randn(H, W, 1, generator=gen)  # upstream gradient g[h,w,d] = u[h,w]*v[d] + 0.25: cheap to regenerate
v_small = torch.randn(1, 1, W, generator=gen)
g = (g_small * v_small + 0.25).contiguous()
cv.backward(g.cuda())
torch.cuda.synchronize()
best, arg = cv.detach().max(dim=-1)
np.savez_compressed(os.path.join(OUT, "cfg1_rand_k5_240x320.npz"), camera=cam, projector=proj, kernel_size=np.int32(k),
                    grad_u=g_small.numpy(), grad_v=v_small.numpy(), grad_bias=np.float32(0.25),
                    best=best.cpu().numpy(), argmax=arg.cpu().numpy().astype(np.int16),
                    camera_grad=camg.grad.cpu().numpy(),
                    cost_sub=cv.detach()[::16, ::16, :].cpu().numpy())
top2 = cv.detach().topk(2, dim=-1).values
manifest["cfg1"] = {"min_top2_gap": float((top2[..., 0] - top2[..., 1]).min()), "grad_absmax": float(camg.grad.abs().max())}
print(json.dumps(manifest["cfg1"]), flush=True)
json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
print("GOLDEN DONE")
