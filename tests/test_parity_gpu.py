"""GPU parity tests: the CUDA path (through the C ABI: custma -> custereomatching_b200.binding -> libcustma_b200.so)
against the golden vectors of the reference extension and against the CPU oracle on the same seeded inputs.

Tolerances (conftest.py): costs |new-ref| <= 1e-5 * max(1,|ref|); gradients |new-ref| <= 1e-5 * max|ref grad|;
WTA indices bit-exact wherever the best cost beats the runner-up by more than 1e-5 (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest
import torch

import custma
import custereomatching_b200 as cb
from conftest import (COST_TOL, GRAD_TOL, assert_cost_close, assert_grad_close, assert_grad_close_or_nearer_truth,
                      golden_manifest, golden_small_names, load_golden)
from oracle import ref_port
from oracle import zncc_oracle as zo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = golden_small_names()
FLAGS = [0, cb.FLAG_DIRECT]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_pair(H, W, seed, B=None):
    rng = np.random.RandomState(seed)
    shape = (H, W) if B is None else (B, H, W)
    return rng.rand(*shape).astype(np.float32), rng.rand(*shape).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# reference-shaped [H,W,W] drop-in path against the reference extension's own outputs
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SMALL)
def test_drop_in_forward_backward_vs_reference_golden(name):
    g = load_golden(name)
    k = int(g["kernel_size"])
    cam = dev(g["camera"]).requires_grad_(True)
    proj = dev(g["projector"])
    cv = custma.stereo_matching(cam, proj, 64, k)       # D is accepted and ignored, as in the reference
    assert cv.shape == g["cost_volume"].shape and cv.dtype == torch.float32
    assert_cost_close(cv.detach().cpu().numpy(), g["cost_volume"])
    cv.backward(dev(g["cost_volume_grad"]))
    assert cam.grad.shape == cam.shape
    truth = zo.camera_grad_full_kernel_formula(g["camera"], g["projector"], g["cost_volume_grad"], k).numpy()
    assert_grad_close_or_nearer_truth(cam.grad.cpu().numpy(), g["camera_grad"], truth)


@pytest.mark.parametrize("name", SMALL)
def test_direct_kernels_forward_bit_exact_with_reference(name):
    g = load_golden(name)
    k = int(g["kernel_size"])
    cost, best, idx = cb.forward(dev(g["camera"]), dev(g["projector"]), 0, k, want_cost=True, want_wta=True,
                                 flags=cb.FLAG_DIRECT)
    assert np.array_equal(cost.cpu().numpy(), g["cost_volume"])
    tb, ti = torch.from_numpy(g["cost_volume"]).max(dim=-1)
    assert np.array_equal(best.cpu().numpy(), tb.numpy())
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ti.numpy())
    grad = cb.backward(dev(g["cost_volume_grad"]), dev(g["camera"]), dev(g["projector"]), k, 0, flags=cb.FLAG_DIRECT)
    truth = zo.camera_grad_full_kernel_formula(g["camera"], g["projector"], g["cost_volume_grad"], k).numpy()
    assert_grad_close_or_nearer_truth(grad.cpu().numpy(), g["camera_grad"], truth)


@pytest.mark.parametrize("flags", FLAGS)
def test_cfg1_reference_fixture(flags):
    """BASELINE.json configs[0]: 320x240, 5x5 window, forward + WTA + backward."""
    g = load_golden("cfg1_rand_k5_240x320")
    k = int(g["kernel_size"])
    cam, proj = dev(g["camera"]), dev(g["projector"])
    cost, best, idx = cb.forward(cam, proj, 0, k, want_cost=True, want_wta=True, flags=flags)
    assert_cost_close(cost[::16, ::16, :].cpu().numpy(), g["cost_sub"])
    assert_cost_close(best.cpu().numpy(), g["best"])
    ref_cv = ref_port.forward_full(g["camera"], g["projector"], k)
    gap = zo.top2_gap(torch.from_numpy(ref_cv)).numpy()
    sel = gap > COST_TOL
    assert sel.mean() > 0.99
    assert np.array_equal(idx.cpu().numpy()[sel], g["argmax"].astype(np.int32)[sel])
    upstream = dev((g["grad_u"] * g["grad_v"] + g["grad_bias"]).astype(np.float32))
    grad = cb.backward(upstream, cam, proj, k, 0, flags=flags)
    assert_grad_close(grad.cpu().numpy(), g["camera_grad"])


# ---------------------------------------------------------------------------------------------------------------
# banded / batched extension against the CPU oracle (C restatement of the reference kernels)
# ---------------------------------------------------------------------------------------------------------------
BANDED_CASES = [  # H, W, D, k
    (24, 40, 16, 5), (17, 33, 7, 3), (9, 21, 32, 5), (12, 50, 64, 7), (30, 47, 1, 5), (16, 64, 64, 4),
    (5, 96, 48, 15), (3, 10, 4, 5), (40, 131, 96, 5), (2, 2, 2, 1), (33, 260, 192, 5), (64, 300, 256, 5),
    # fast-path instances for the other window sizes, with interior / masked / scalar tiles and several row bands
    (150, 300, 64, 3), (70, 420, 192, 3), (90, 300, 64, 7), (40, 280, 128, 7), (200, 520, 192, 5), (37, 700, 100, 5),
]


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("H,W,D,k", BANDED_CASES)
def test_banded_forward_wta_backward_vs_oracle(H, W, D, k, flags):
    cam, proj = rand_pair(H, W, seed=H * 1000 + W)
    ref = ref_port.forward_banded(cam, proj, D, k)
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True, flags=flags)
    cost = cost.cpu().numpy()
    assert cost.shape == (H, W, D)
    assert_cost_close(cost, ref)
    invalid = np.arange(W)[:, None] - np.arange(D)[None, :] < 0
    assert (cost[:, invalid] == cb.INVALID_COST).all()
    # WTA: identical to the WTA of the kernel's own volume (bit-exact), and to the oracle's outside near-ties
    ob, od = ref_port.wta_banded(cost)
    assert np.array_equal(best.cpu().numpy(), ob)
    assert np.array_equal(disp.cpu().numpy(), od)
    rb, rd = ref_port.wta_banded(ref)
    top = np.sort(np.where(invalid[None], -np.inf, ref), axis=-1)
    gap = top[..., -1] - (top[..., -2] if D > 1 else -np.inf)
    sel = gap > COST_TOL
    assert np.array_equal(disp.cpu().numpy()[sel], rd[sel])
    assert_cost_close(best.cpu().numpy(), rb)
    # WTA-only call (no volume) gives the same answer
    b2, d2 = cb.wta(dev(cam), dev(proj), D, k, flags=flags)
    assert torch.equal(b2, best) and torch.equal(d2, disp)
    # backward
    gb = np.random.RandomState(7).randn(H, W, D).astype(np.float32)
    gref = ref_port.backward_banded(gb, cam, proj, k)
    grad = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=flags)
    assert_grad_close(grad.cpu().numpy(), gref)


@pytest.mark.parametrize("flags", FLAGS)
def test_invalid_cells_carry_no_gradient(flags):
    H, W, D, k = 12, 20, 16, 5
    cam, proj = rand_pair(H, W, 3)
    gb = np.random.RandomState(1).randn(H, W, D).astype(np.float32)
    gb2 = gb.copy()
    invalid = np.arange(W)[:, None] - np.arange(D)[None, :] < 0
    gb2[:, invalid] = 1e6                                         # garbage on invalid cells must not leak
    a = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=flags)
    b = cb.backward(dev(gb2), dev(cam), dev(proj), k, D, flags=flags)
    assert torch.equal(a, b)


@pytest.mark.parametrize("flags", FLAGS)
def test_batched_equals_per_pair(flags):
    B, H, W, D, k = 3, 20, 70, 32, 5
    cam, proj = rand_pair(H, W, 11, B=B)
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True, flags=flags)
    gb = np.random.RandomState(2).randn(B, H, W, D).astype(np.float32)
    grad = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=flags)
    assert cost.shape == (B, H, W, D) and grad.shape == (B, H, W)
    for b in range(B):
        c1, b1, d1 = cb.forward(dev(cam[b]), dev(proj[b]), D, k, want_cost=True, want_wta=True, flags=flags)
        assert torch.equal(c1, cost[b]) and torch.equal(b1, best[b]) and torch.equal(d1, disp[b])
        g1 = cb.backward(dev(gb[b]), dev(cam[b]), dev(proj[b]), k, D, flags=flags)
        assert torch.equal(g1, grad[b])
        assert_cost_close(c1.cpu().numpy(), ref_port.forward_banded(cam[b], proj[b], D, k))
        assert_grad_close(g1.cpu().numpy(), ref_port.backward_banded(gb[b], cam[b], proj[b], k))


@pytest.mark.parametrize("kind", ["smooth", "edges", "const"])
def test_low_contrast_inputs_keep_parity(kind):
    """Ill-conditioned inputs: the sliding-window path must hand such tiles to the two-pass
    arithmetic, so costs stay within 1e-5 of the reference-order result."""
    H, W, D, k = 24, 96, 48, 5
    rng = np.random.RandomState(5)
    xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    if kind == "smooth":
        cam = (0.8 + 0.01 * np.sin(xx / 7) + 0.002 * rng.rand(H, W)).astype(np.float32)
        proj = (0.8 + 0.01 * np.sin((xx + 3) / 7) + 0.002 * rng.rand(H, W)).astype(np.float32)
    elif kind == "edges":
        cam = (np.where(xx < W // 2, 0.2, 0.8) + 0.002 * rng.rand(H, W)).astype(np.float32)
        proj = rng.rand(H, W).astype(np.float32)
    else:
        cam = np.full((H, W), 0.5, np.float32)
        proj = rng.rand(H, W).astype(np.float32)
    ref = ref_port.forward_banded(cam, proj, D, k)
    cost, _, _ = cb.forward(dev(cam), dev(proj), D, k)
    assert_cost_close(cost.cpu().numpy(), ref)
    gb = rng.randn(H, W, D).astype(np.float32)
    gref = ref_port.backward_banded(gb, cam, proj, k)
    grad = cb.backward(dev(gb), dev(cam), dev(proj), k, D)
    # here the fp32 restatement of the reference is itself several 1e-5 of scale away from the exact gradient
    # (every term is divided by a tiny den); accept 1e-5 of the fp32 reference arithmetic, or a result that is at
    # least as close to the fp64 truth as that arithmetic is
    truth = zo.camera_grad_banded_autograd(cam, proj, gb, D, k).numpy()
    assert_grad_close_or_nearer_truth(grad.cpu().numpy(), gref, truth)


def test_backward_is_deterministic():
    H, W, D, k = 64, 200, 64, 5
    cam, proj = rand_pair(H, W, 21)
    gb = dev(np.random.RandomState(3).randn(H, W, D).astype(np.float32))
    a = cb.backward(gb, dev(cam), dev(proj), k, D)
    b = cb.backward(gb, dev(cam), dev(proj), k, D)
    assert torch.equal(a, b)                                      # the reference's atomics differ run to run


def test_autograd_matches_oracle_autograd():
    H, W, D, k = 20, 48, 24, 5
    cam, proj = rand_pair(H, W, 31)
    camt = dev(cam).requires_grad_(True)
    band = custma.stereo_matching_banded(camt, dev(proj), D, k)
    w = dev(np.random.RandomState(4).randn(H, W, D).astype(np.float32))
    valid = torch.from_numpy(np.arange(W)[:, None] - np.arange(D)[None, :] >= 0).cuda()
    loss = (band * w * valid).sum()
    loss.backward()
    auto = zo.camera_grad_banded_autograd(cam, proj, (w * valid).cpu().numpy(), D, k).numpy()
    assert_grad_close(camt.grad.cpu().numpy(), auto)
    # the Function returns a gradient for the camera only (custma/stereo_matching_wrapper.py:33)
    projt = dev(proj).requires_grad_(True)
    camt2 = dev(cam).requires_grad_(True)
    cv = custma.stereo_matching(camt2, projt, 0, k)
    # .sum().backward() would hand over an expanded (non-contiguous) gradient, which the reference rejects
    # (CHECK_INPUT at custma/src/stereo_matching.cpp:52) and so does this implementation (test_error_behaviour)
    cv.backward(torch.ones_like(cv))
    assert projt.grad is None and camt2.grad is not None


def test_runs_on_the_current_stream():
    H, W, D, k = 64, 256, 64, 5
    cam, proj = rand_pair(H, W, 41)
    ref, _, _ = cb.forward(dev(cam), dev(proj), D, k)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        c = dev(cam)
        p = dev(proj)
        out, _, _ = cb.forward(c, p, D, k)
    s.synchronize()
    assert torch.equal(out, ref)


# ---------------------------------------------------------------------------------------------------------------
# error behaviour (custma/include/stereo_matching.hpp:20-29, custma/src/stereo_matching.cpp:23-24,52)
# ---------------------------------------------------------------------------------------------------------------
def test_host_submit_and_wait_stream_batches():
    """custma_host_submit / custma_host_wait: two batches in flight give the same results as the device path."""
    from custereomatching_b200 import binding
    B, H, W, D, k = 3, 40, 200, 64, 5
    sets = []
    for i in range(2):
        cam, proj = rand_pair(H, W, seed=50 + i, B=B)
        g = np.random.RandomState(60 + i).randn(B, H, W, D).astype(np.float32)
        sets.append((torch.from_numpy(cam).pin_memory(), torch.from_numpy(proj).pin_memory(), dev(g),
                     torch.empty(B, H, W).pin_memory(), torch.empty(B, H, W, dtype=torch.int32).pin_memory(),
                     torch.empty(B, H, W).pin_memory()))
    tickets = [binding.host_submit(c.data_ptr(), p.data_ptr(), hb.data_ptr(), hi.data_ptr(), hg.data_ptr(), 0, g.data_ptr(),
                                   B, H, W, D, k) for c, p, g, hb, hi, hg in sets]
    assert tickets[1] == tickets[0] + 1
    for t, (c, p, g, hb, hi, hg) in zip(tickets, sets):
        binding.host_wait(t)
        _, best, disp = cb.forward(c.cuda(), p.cuda(), D, k, want_cost=False, want_wta=True)
        grad = cb.backward(g, c.cuda(), p.cuda(), k, D)
        assert torch.equal(hb.cuda(), best) and torch.equal(hi.cuda(), disp) and torch.equal(hg.cuda(), grad)
    binding.host_wait(0)
    binding.host_release()


def test_host_entry_points_keep_two_shapes_alive():
    """Alternating between two shapes must not invalidate tickets of the other shape (the host
    context used to be a single global that a second shape tore down)."""
    from custereomatching_b200 import binding
    shapes = [(2, 30, 120, 32, 5), (1, 44, 90, 16, 3)]
    work = []
    for i, (B, H, W, D, k) in enumerate(shapes * 2):
        cam, proj = rand_pair(H, W, seed=70 + i, B=B)
        c, p_ = torch.from_numpy(cam).pin_memory(), torch.from_numpy(proj).pin_memory()
        hb, hi = torch.empty(B, H, W).pin_memory(), torch.empty(B, H, W, dtype=torch.int32).pin_memory()
        t = binding.host_submit(c.data_ptr(), p_.data_ptr(), hb.data_ptr(), hi.data_ptr(), 0, 0, 0, B, H, W, D, k)
        work.append((t, c, p_, hb, hi, D, k))
    for t, c, p_, hb, hi, D, k in work:
        binding.host_wait(t)
        best, disp = cb.wta(c.cuda(), p_.cuda(), D, k)
        assert torch.equal(hb.cuda(), best) and torch.equal(hi.cuda(), disp)
    binding.host_wait(0)
    binding.host_release()


@pytest.mark.parametrize("shape", ["single", "slots"])
def test_host_pipeline_shapes_agree_with_the_device_path(shape, monkeypatch):
    """Both shapes of the host pipeline (one compute stream for every chunk | a kernel stream per buffer slot;
    host_pipeline.cu picks by chunk size, CUSTMA_HOST_PIPELINE forces one) deliver the device path's bits, streamed
    (three tickets in flight over two slots) and synchronous (custma_host_step cuts the batch into two chunks)."""
    from custereomatching_b200 import binding
    binding.host_release()
    monkeypatch.setenv("CUSTMA_HOST_PIPELINE", shape)
    B, H, W, D, k = 4, 36, 150, 32, 5
    sets = []
    for i in range(3):
        cam, proj = rand_pair(H, W, seed=80 + i, B=B)
        g = np.random.RandomState(90 + i).randn(B, H, W, D).astype(np.float32)
        sets.append((torch.from_numpy(cam).pin_memory(), torch.from_numpy(proj).pin_memory(), dev(g),
                     torch.empty(B, H, W).pin_memory(), torch.empty(B, H, W, dtype=torch.int32).pin_memory(),
                     torch.empty(B, H, W).pin_memory()))
    torch.cuda.synchronize()
    tickets = [binding.host_submit(c.data_ptr(), p.data_ptr(), hb.data_ptr(), hi.data_ptr(), hg.data_ptr(), 0, g.data_ptr(),
                                   B, H, W, D, k) for c, p, g, hb, hi, hg in sets]
    binding.host_wait(tickets[-1])
    binding.host_wait(0)
    for c, p, g, hb, hi, hg in sets:
        _, best, disp = cb.forward(c.cuda(), p.cuda(), D, k, want_cost=False, want_wta=True)
        grad = cb.backward(g, c.cuda(), p.cuda(), k, D)
        assert torch.equal(hb.cuda(), best) and torch.equal(hi.cuda(), disp) and torch.equal(hg.cuda(), grad)
    # synchronous entry point on the same context
    c, p, g, hb, hi, hg = sets[0]
    hb.zero_(); hi.zero_(); hg.zero_()
    binding.host_step(c.data_ptr(), p.data_ptr(), hb.data_ptr(), hi.data_ptr(), hg.data_ptr(), 0, g.data_ptr(), B, H, W, D, k)
    _, best, disp = cb.forward(c.cuda(), p.cuda(), D, k, want_cost=False, want_wta=True)
    assert torch.equal(hb.cuda(), best) and torch.equal(hi.cuda(), disp)
    assert torch.equal(hg.cuda(), cb.backward(g, c.cuda(), p.cuda(), k, D))
    binding.host_release()


def test_plain_launches_give_the_same_bits_as_chained_launches():
    """CUSTMA_NO_PDL=1 (read once per process) launches every kernel in plain stream order; programmatic dependent
    launch only overlaps launch latencies, so forward, WTA and backward must not change by a bit."""
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, hashlib\n"
        "from custereomatching_b200 import functional as cb\n"
        "rng = np.random.RandomState(5)\n"
        "cam = torch.from_numpy(rng.rand(2, 40, 170).astype(np.float32)).cuda()\n"
        "proj = torch.from_numpy(rng.rand(2, 40, 170).astype(np.float32)).cuda()\n"
        "g = torch.from_numpy(rng.randn(2, 40, 170, 64).astype(np.float32)).cuda()\n"
        "cost, best, disp = cb.forward(cam, proj, 64, 5, want_cost=True, want_wta=True)\n"
        "grad = cb.backward(g, cam, proj, 5, 64)\n"
        "h = hashlib.sha256()\n"
        "for t in (cost, best, disp, grad): h.update(t.cpu().numpy().tobytes())\n"
        "print('DIGEST', h.hexdigest())\n")
    digests = []
    for no_pdl in (False, True):
        env = dict(os.environ)
        env.pop("CUSTMA_NO_PDL", None)
        if no_pdl:
            env["CUSTMA_NO_PDL"] = "1"
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]


def test_error_behaviour():
    cam = torch.rand(16, 16, device="cuda")
    with pytest.raises(RuntimeError, match="camera must be contiguous"):
        custma.stereo_matching(cam.t(), cam, 4, 5)
    with pytest.raises(RuntimeError, match="projector must be a CUDA tensor"):
        custma.stereo_matching(cam, cam.cpu(), 4, 5)
    with pytest.raises(RuntimeError, match="float32"):
        custma.stereo_matching(cam.double(), cam.double(), 4, 5)
    with pytest.raises(RuntimeError, match="same shape"):
        custma.stereo_matching(cam, cam[:8].contiguous(), 4, 5)
    with pytest.raises(RuntimeError, match="kernel_size"):
        custma.stereo_matching(cam, cam, 4, 0)
    with pytest.raises(RuntimeError, match="kernel_size"):
        custma.stereo_matching(cam, cam, 4, 33)
    camg = cam.clone().requires_grad_(True)
    cv = custma.stereo_matching(camg, cam, 4, 5)
    with pytest.raises(RuntimeError, match="cost_volume_grad must be contiguous"):
        cv.backward(torch.ones(16, 16, 1, device="cuda").expand(16, 16, 16))   # reference: CHECK_INPUT at cpp:52
    with pytest.raises(RuntimeError, match="shape"):
        custma.src.stereo_matching_backward(torch.ones(16, 16, 8, device="cuda"), cam, cam, 5)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + row-cropped oracle comparison
# ---------------------------------------------------------------------------------------------------------------
def _row_crop_check(cam, proj, cost, grad_in, D, k, rows, H, best=None, disp=None):
    """Oracle on a crop of image rows [h0-r, h1+r): rows h0..h1 of the volume depend on nothing else.  Costs against the
    oracle; with best / disp given, also the fused WTA against the ORACLE's winner-take-all: best within tolerance
    everywhere, disparities bit-exact wherever the oracle's best beats its runner-up by more than the tolerance
    (BASELINE.json north_star)."""
    r = k // 2
    h0, h1 = rows
    lo, hi = max(0, h0 - r), min(H, h1 + (k - 1 - r))
    ref = ref_port.forward_banded(cam[lo:hi], proj[lo:hi], D, k)[h0 - lo:h1 - lo]
    assert_cost_close(cost[h0:h1].cpu().numpy(), ref)
    if best is not None:
        W = ref.shape[1]
        rb, rd = ref_port.wta_banded(ref)
        invalid = np.arange(W)[:, None] - np.arange(D)[None, :] < 0
        top = np.sort(np.where(invalid[None], -np.inf, ref), axis=-1)
        gap = top[..., -1] - (top[..., -2] if D > 1 else -np.inf)
        sel = gap > COST_TOL
        assert sel.mean() > 0.9
        assert_cost_close(best[h0:h1].cpu().numpy(), rb, what="best vs oracle")
        assert np.array_equal(disp[h0:h1].cpu().numpy()[sel], rd[sel]), "disparity differs from the oracle outside near-ties"


def _row_crop_backward_check(cam, proj, camd, projd, D, k, rows, H, seed, flags=0):
    """Backward of an upstream gradient that lives only in volume rows [h0,h1), against the oracle on the cropped
    image rows; everything outside the rows the window can reach must be exactly zero."""
    h0, h1 = rows
    r = k // 2
    W = cam.shape[-1]
    lo, hi = max(0, h0 - r), min(H, h1 + (k - 1 - r))
    gen = torch.Generator(device="cuda").manual_seed(seed)
    g = torch.zeros(H, W, D, device="cuda")
    g[h0:h1] = torch.randn(h1 - h0, W, D, device="cuda", generator=gen)
    gc = cb.backward(g, camd, projd, k, D, flags=flags).cpu().numpy()
    # the oracle sees the crop as a whole image: its volume rows are the crop's rows, the gradient sits on rows
    # [h0-lo, h1-lo) of it (rows whose windows would read beyond the crop carry no gradient, so the crop's zero
    # padding is never consulted)
    gsub = np.zeros((hi - lo, W, D), np.float32)
    gsub[h0 - lo:h1 - lo] = g[h0:h1].cpu().numpy()
    gref = ref_port.backward_banded(gsub, cam[lo:hi], proj[lo:hi], k)
    assert_grad_close(gc[lo:hi], gref)
    assert np.abs(gc[:lo]).max(initial=0) == 0 and np.abs(gc[hi:]).max(initial=0) == 0


def test_cfg2_kitti_full_size_forward_wta():
    """BASELINE.json configs[1]: 1242x375, 192 disparities, forward + WTA."""
    H, W, D, k = 375, 1242, 192, 5
    cam, proj = rand_pair(H, W, 2)
    camd, projd = dev(cam), dev(proj)
    cost, best, disp = cb.forward(camd, projd, D, k, want_cost=True, want_wta=True)
    # WTA of the fused kernel == WTA of its own volume, ties to the largest disparity (lowest projector column)
    flipped = torch.flip(cost, dims=[-1])
    tb, ti = flipped.max(dim=-1)
    assert torch.equal(best, tb)
    assert torch.equal(disp.long(), (D - 1) - ti)
    b2, d2 = cb.wta(camd, projd, D, k)
    assert torch.equal(b2, best) and torch.equal(d2, disp)
    for rows in [(0, 6), (180, 186), (369, 375)]:
        _row_crop_check(cam, proj, cost, None, D, k, rows, H, best, disp)
    # KITTI-size backward against the oracle (interior rows and both image borders)
    for i, rows in enumerate([(0, 3), (186, 190), (372, 375)]):
        _row_crop_backward_check(cam, proj, camd, projd, D, k, rows, H, seed=20 + i)
    # known-answer property: shifting the projector by s0 columns makes disparity s0 the winner with cost ~ 1
    s0 = 37
    cam2 = np.zeros_like(proj)
    cam2[:, s0:] = proj[:, :-s0]
    b3, d3 = cb.wta(dev(cam2), projd, D, k)
    inner = (slice(k, H - k), slice(s0 + k, W - k))
    assert (d3[inner] == s0).all()
    assert (b3[inner] > 0.9999).all()


def test_cfg3_middlebury_full_size_properties():
    """BASELINE.json configs[2]: 2880x1988, 256 disparities, forward + backward (5.9 GB volume)."""
    H, W, D, k = 1988, 2880, 256, 5
    cam, proj = rand_pair(H, W, 3)
    camd, projd = dev(cam), dev(proj)
    cost, best, disp = cb.forward(camd, projd, D, k, want_cost=True, want_wta=True)
    assert cost.shape == (H, W, D)
    valid = (torch.arange(W, device="cuda")[:, None] - torch.arange(D, device="cuda")[None, :]) >= 0
    assert bool((cost[:, ~valid] == cb.INVALID_COST).all())
    assert float(cost[:, valid].abs().max()) <= 1.0 + 1e-5          # ZNCC is a correlation coefficient
    for rows in [(0, 4), (1000, 1004), (1984, 1988)]:
        _row_crop_check(cam, proj, cost, None, D, k, rows, H, best, disp)
    tb = torch.flip(cost, dims=[-1]).max(dim=-1).values
    assert torch.equal(best, tb)
    del tb
    # backward: linear in the upstream gradient, and equal to the oracle on a cropped band of rows
    gen = torch.Generator(device="cuda").manual_seed(1)
    g1 = torch.randn(H, W, D, device="cuda", generator=gen)
    ga = cb.backward(g1, camd, projd, k, D)
    g1.mul_(-2.0)
    gb = cb.backward(g1, camd, projd, k, D)
    assert_grad_close(gb.cpu().numpy(), (-2.0 * ga).cpu().numpy(), 1e-6)
    # gradient only of rows [h0,h1): zero the rest, compare with the oracle on the cropped rows
    h0, h1 = 700, 704
    g1.zero_()
    g1[h0:h1] = torch.randn(h1 - h0, W, D, device="cuda", generator=gen)
    gc = cb.backward(g1, camd, projd, k, D).cpu().numpy()
    r = k // 2
    lo, hi = h0 - r, h1 + (k - 1 - r)
    gsub = np.zeros((hi - lo, W, D), np.float32)
    gsub[h0 - lo:h1 - lo] = g1[h0:h1].cpu().numpy()
    gref = ref_port.backward_banded(gsub, cam[lo:hi], proj[lo:hi], k)
    assert_grad_close(gc[lo:hi], gref)
    assert np.abs(gc[:lo]).max() == 0 and np.abs(gc[hi:]).max() == 0


def test_cfg4_kitti_batch_of_8_parity():
    """BASELINE.json configs[3] as one rank sees it (8 of the 64 KITTI-size pairs): the batched call equals the
    per-pair calls bit for bit on two sampled pairs (forward, WTA, backward), and the sampled pairs agree with the
    oracle on row crops (costs, best, disparities outside near-ties, camera gradient)."""
    B, H, W, D, k = 8, 375, 1242, 192, 5
    cam, proj = rand_pair(H, W, 404, B=B)
    camd, projd = dev(cam), dev(proj)
    cost, best, disp = cb.forward(camd, projd, D, k, want_cost=True, want_wta=True)
    gen = torch.Generator(device="cuda").manual_seed(9)
    g = torch.randn(B, H, W, D, device="cuda", generator=gen)
    grad = cb.backward(g, camd, projd, k, D)
    for b in (2, 7):
        c1, b1, d1 = cb.forward(camd[b].contiguous(), projd[b].contiguous(), D, k, want_cost=True, want_wta=True)
        assert torch.equal(c1, cost[b]) and torch.equal(b1, best[b]) and torch.equal(d1, disp[b])
        g1 = cb.backward(g[b].contiguous(), camd[b].contiguous(), projd[b].contiguous(), k, D)
        assert torch.equal(g1, grad[b])
        del c1, g1
        for rows in [(0, 4), (201, 205), (371, 375)]:
            _row_crop_check(cam[b], proj[b], cost[b], None, D, k, rows, H, best[b], disp[b])
        _row_crop_backward_check(cam[b], proj[b], camd[b].contiguous(), projd[b].contiguous(), D, k, (120, 124), H, seed=b)


# ---------------------------------------------------------------------------------------------------------------
# partitioning on the GPU: the per-rank work of the two shardings, executed rank by rank on one device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_row_band_ranks_reproduce_the_single_gpu_result(world):
    """Each emulated rank computes volume rows [h0,h1) from its haloed crop (custereomatching_b200/sharding.py);
    bands stitched together equal the unsharded result (costs / WTA bit-exact, gradient to rounding)."""
    from custereomatching_b200 import sharding as sh
    H, W, D, k = 83, 300, 64, 5
    cam, proj = rand_pair(H, W, 51)
    camd, projd = dev(cam), dev(proj)
    gfull = dev(np.random.RandomState(6).randn(H, W, D).astype(np.float32))
    cost, best, disp = cb.forward(camd, projd, D, k, want_cost=True, want_wta=True)
    grad = cb.backward(gfull, camd, projd, k, D)
    grad_sum = torch.zeros_like(grad)
    for rank in range(world):
        band = sh.row_band(H, k, rank, world)
        cam_c, proj_c = sh.crop_rows_with_halo(camd, band), sh.crop_rows_with_halo(projd, band)
        c, b, d = cb.forward(cam_c, proj_c, D, k, want_cost=True, want_wta=True)
        assert_cost_close(sh.band_rows_of(c, band).cpu().numpy(), cost[band.h0:band.h1].cpu().numpy(), 2e-6)
        sel = (torch.topk(cost[band.h0:band.h1], 2, dim=-1).values.diff(dim=-1).abs() > COST_TOL)[..., 0]
        assert torch.equal(sh.band_rows_of(d, band)[sel], disp[band.h0:band.h1][sel])
        grad_sum[band.lo:band.hi] += cb.backward(gfull[band.h0:band.h1].contiguous(), cam_c, proj_c, k, D,
                                                 rows=sh.band_gradient_mask_rows(band))
    assert_grad_close(grad_sum.cpu().numpy(), grad.cpu().numpy())


def test_batch_slices_reproduce_the_batched_result():
    from custereomatching_b200 import sharding as sh
    B, H, W, D, k = 5, 40, 200, 64, 5
    cam, proj = rand_pair(H, W, 61, B=B)
    g = dev(np.random.RandomState(8).randn(B, H, W, D).astype(np.float32))
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True)
    grad = cb.backward(g, dev(cam), dev(proj), k, D)
    for rank in range(2):
        b0, b1 = sh.batch_slice(B, rank, 2)
        c, b, d = cb.forward(dev(cam[b0:b1]), dev(proj[b0:b1]), D, k, want_cost=True, want_wta=True)
        assert torch.equal(c, cost[b0:b1]) and torch.equal(b, best[b0:b1]) and torch.equal(d, disp[b0:b1])
        assert torch.equal(cb.backward(g[b0:b1].contiguous(), dev(cam[b0:b1]), dev(proj[b0:b1]), k, D), grad[b0:b1])


def test_more_than_2_31_cells():
    """One row band of BASELINE.json configs[4] (7680 wide, 512 disparities): 2.4e9 cells - past the int32 element
    count the reference overflows at (custma/src/stereo_matching_kernel.cu:194).  Size-independent checks."""
    H, W, D, k = 600, 7680, 512, 5
    assert H * W * D > 2 ** 31
    gen = torch.Generator(device="cuda").manual_seed(5)
    cam = torch.rand(H, W, device="cuda", generator=gen)
    proj = torch.rand(H, W, device="cuda", generator=gen)
    cost, best, disp = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True)
    for rows in [(0, 3), (297, 300), (597, 600)]:       # first, middle and last rows (the last ones live past 2^31)
        _row_crop_check(cam.cpu().numpy(), proj.cpu().numpy(), cost, None, D, k, rows, H, best, disp)
        h0, h1 = rows
        tb, ti = torch.flip(cost[h0:h1], dims=[-1]).max(dim=-1)
        assert torch.equal(best[h0:h1], tb) and torch.equal(disp[h0:h1].long(), (D - 1) - ti)
    # backward of a gradient that lives only in the last rows, against the oracle on the cropped rows
    g = torch.zeros_like(cost)
    del cost
    h0, h1 = 596, 600
    g[h0:h1] = torch.randn(h1 - h0, W, D, device="cuda", generator=gen)
    gc = cb.backward(g, cam, proj, k, D).cpu().numpy()
    r = k // 2
    lo = h0 - r
    gsub = np.zeros((H - lo, W, D), np.float32)
    gsub[h0 - lo:] = g[h0:h1].cpu().numpy()
    gref = ref_port.backward_banded(gsub, cam[lo:].cpu().numpy(), proj[lo:].cpu().numpy(), k)
    assert_grad_close(gc[lo:], gref)
    assert np.abs(gc[:lo]).max() == 0


@pytest.mark.parametrize("kind", ["pedestal", "weak_pedestal", "ramp", "ramp_weak", "blocks", "shifted"])
def test_image_families_keep_parity(kind):
    """Inputs that are neither uniform-random nor flat: whichever path the conditioning verdict picks per tile (fast
    window sums or direct arithmetic), costs and gradients stay within tolerance of the reference arithmetic."""
    H, W, D, k = 100, 420, 192, 5
    rng = np.random.RandomState(11)
    xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    yy = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    a, b = rng.rand(H, W).astype(np.float32), rng.rand(H, W).astype(np.float32)
    if kind == "pedestal":
        cam, proj = 0.25 + 0.5 * a, 0.25 + 0.5 * b
    elif kind == "weak_pedestal":
        cam, proj = 0.4 + 0.2 * a, 0.35 + 0.3 * b
    elif kind == "ramp":
        cam, proj = 0.2 + 0.6 * xx / W + 0.2 * (a - 0.5), 0.2 + 0.6 * xx / W + 0.2 * (b - 0.5)
    elif kind == "ramp_weak":
        cam, proj = 0.2 + 0.6 * yy / H + 0.05 * (a - 0.5), 0.8 - 0.6 * xx / W + 0.05 * (b - 0.5)
    elif kind == "blocks":
        cam = np.where((xx // 37 + yy // 23) % 2 == 0, 0.15, 0.85) + 0.1 * (a - 0.5)
        proj = np.where((xx // 29) % 2 == 0, 0.3, 0.7) + 0.2 * (b - 0.5)
    else:
        proj = b
        cam = np.zeros_like(b)
        cam[:, 40:] = b[:, :-40]
        cam = cam + 0.01 * rng.randn(H, W)
    cam, proj = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(proj, np.float32)
    ref = ref_port.forward_banded(cam, proj, D, k)
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True)
    assert_cost_close(cost.cpu().numpy(), ref)
    ob, od = ref_port.wta_banded(cost.cpu().numpy())
    assert np.array_equal(best.cpu().numpy(), ob) and np.array_equal(disp.cpu().numpy(), od)
    gb = rng.randn(H, W, D).astype(np.float32)
    gref = ref_port.backward_banded(gb, cam, proj, k)
    truth = zo.camera_grad_banded_autograd(cam, proj, gb, D, k).numpy()
    grad = cb.backward(dev(gb), dev(cam), dev(proj), k, D)
    assert_grad_close_or_nearer_truth(grad.cpu().numpy(), gref, truth)


def test_fuzz_fast_path_against_direct_path():
    """Random shapes (batch, odd sizes, every fast-path window size, banded and reference-shaped): the sliding-window
    kernels against the direct two-pass kernels (which are themselves pinned to the reference), same tolerances."""
    rng = np.random.RandomState(2026)
    for case in range(48):
        k = int(rng.choice([3, 5, 5, 5, 7]))
        B = int(rng.choice([1, 1, 2, 3]))
        H = int(rng.randint(1, 140))
        W = int(rng.randint(1, 330))
        D = int(rng.choice([0, 1, 4, 17, 32, 64, 100, 128, 192, 200, 256]))
        if D == 0 and W > 200:
            W = 200                                    # keep the [H,W,W] volumes small
        shape = (B, H, W) if B > 1 else (H, W)
        pedestal, contrast = float(rng.choice([0.0, 0.3])), float(rng.choice([1.0, 0.4]))
        cam = dev((pedestal + contrast * rng.rand(*shape)).astype(np.float32))
        proj = dev((pedestal + contrast * rng.rand(*shape)).astype(np.float32))
        tag = f"case {case}: B={B} H={H} W={W} D={D} k={k} pedestal={pedestal} contrast={contrast}"
        c0, b0, i0 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=cb.FLAG_DIRECT)
        c1, b1, i1 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True)
        assert_cost_close(c1.cpu().numpy(), c0.cpu().numpy(), what=tag + " cost")
        if D > 0 and D % 4 == 0 and k in (3, 5):     # the tensor-core forward on the same input
            c2, b2, i2 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
            assert_cost_close(c2.cpu().numpy(), c0.cpu().numpy(), what=tag + " tensor-core cost")
            tb, ti = torch.flip(c2, dims=[-1]).max(dim=-1)
            assert torch.equal(b2, tb) and torch.equal(i2.long(), (D - 1) - ti), tag + " tensor-core wta"
        # the fused WTA is the arg-max of the kernel's own volume (ties to the lowest projector column)
        if D > 0:
            tb, ti = torch.flip(c1, dims=[-1]).max(dim=-1)
            assert torch.equal(b1, tb) and torch.equal(i1.long(), (D - 1) - ti), tag + " wta"
        else:
            tb, ti = c1.max(dim=-1)
            assert torch.equal(b1, tb) and torch.equal(i1.long(), ti), tag + " wta"
        C = D if D > 0 else W
        g = torch.from_numpy(rng.randn(*(shape + (C,))).astype(np.float32)).cuda()
        g0 = cb.backward(g, cam, proj, k, D, flags=cb.FLAG_DIRECT)
        g1 = cb.backward(g, cam, proj, k, D)
        assert_grad_close(g1.cpu().numpy(), g0.cpu().numpy(), what=tag + " grad")          # GRAD_TOL = 1e-5
        if D > 0 and D % 4 == 0 and k in (3, 5):
            g2 = cb.backward(g, cam, proj, k, D, flags=cb.FLAG_TENSOR)
            assert_grad_close(g2.cpu().numpy(), g0.cpu().numpy(), what=tag + " tensor-core grad")


# ---------------------------------------------------------------------------------------------------------------
# tensor-core forward (tc_forward.cu): forced with FLAG_TENSOR, and picked by the verdict for low-texture inputs
# ---------------------------------------------------------------------------------------------------------------
TENSOR_CASES = [  # H, W, D, k  (banded, D % 4 == 0, k = 3 or 5)
    (24, 40, 16, 5), (9, 21, 32, 5), (3, 10, 4, 5), (40, 131, 96, 5), (33, 260, 192, 5), (64, 300, 256, 5),
    (150, 300, 64, 3), (70, 420, 192, 3), (200, 520, 192, 5), (37, 700, 100, 5), (12, 140, 572, 5), (1, 1, 4, 3),
    (45, 128, 44, 5), (45, 129, 48, 5), (7, 257, 220, 3),
]


@pytest.mark.parametrize("H,W,D,k", TENSOR_CASES)
def test_tensor_core_forward_vs_oracle(H, W, D, k):
    cam, proj = rand_pair(H, W, seed=H * 1000 + W + 1)
    ref = ref_port.forward_banded(cam, proj, D, k)
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
    cost = cost.cpu().numpy()
    assert_cost_close(cost, ref)
    invalid = np.arange(W)[:, None] - np.arange(D)[None, :] < 0
    assert (cost[:, invalid] == cb.INVALID_COST).all()
    ob, od = ref_port.wta_banded(cost)
    assert np.array_equal(best.cpu().numpy(), ob)
    assert np.array_equal(disp.cpu().numpy(), od)
    rb, rd = ref_port.wta_banded(ref)
    top = np.sort(np.where(invalid[None], -np.inf, ref), axis=-1)
    gap = top[..., -1] - (top[..., -2] if D > 1 else -np.inf)
    sel = gap > COST_TOL
    assert np.array_equal(disp.cpu().numpy()[sel], rd[sel])
    b2, d2 = cb.wta(dev(cam), dev(proj), D, k, flags=cb.FLAG_TENSOR)
    assert torch.equal(b2, best) and torch.equal(d2, disp)
    c3 = cb.cost_volume(dev(cam), dev(proj), D, k, flags=cb.FLAG_TENSOR)
    assert np.array_equal(c3.cpu().numpy(), cost)


def test_tensor_core_forward_batched_equals_per_pair():
    B, H, W, D, k = 3, 50, 300, 128, 5
    cam, proj = rand_pair(H, W, seed=77, B=B)
    cost, best, disp = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
    for b in range(B):
        c1, b1, d1 = cb.forward(dev(cam[b]), dev(proj[b]), D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
        assert torch.equal(cost[b], c1) and torch.equal(best[b], b1) and torch.equal(disp[b], d1)


def test_tensor_core_flag_is_refused_where_unsupported():
    cam, proj = rand_pair(16, 40, seed=3)
    for D, k in ((0, 5), (6, 5), (32, 7), (576, 5)):
        with pytest.raises(RuntimeError):
            cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=False, flags=cb.FLAG_TENSOR)
    with pytest.raises(RuntimeError):
        cb.forward(dev(cam), dev(proj), 32, 5, want_cost=True, want_wta=False, flags=cb.FLAG_TENSOR | cb.FLAG_DIRECT)


def test_verdict_hands_low_texture_inputs_to_the_tensor_core_kernel():
    """A smooth scene with 1 % noise flags (nearly) every tile of the sliding-window path; the default call must then
    give exactly the tensor-core kernel's bits - and stay within tolerance of the reference arithmetic.  A textured
    input must keep running the sliding-window kernels (different bits, same tolerance)."""
    H, W, D, k = 96, 600, 192, 5
    rng = np.random.RandomState(5)
    xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    yy = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    scene = lambda sh: 0.5 + 0.4 * np.sin((xx + sh) * 0.01) * np.cos(yy * 0.02)
    cam = np.ascontiguousarray(scene(0) + 0.01 * (rng.rand(H, W) - 0.5), np.float32)
    proj = np.ascontiguousarray(scene(40) + 0.01 * (rng.rand(H, W) - 0.5), np.float32)
    ref = ref_port.forward_banded(cam, proj, D, k)
    c_def, b_def, d_def = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True)
    c_tc, b_tc, d_tc = cb.forward(dev(cam), dev(proj), D, k, want_cost=True, want_wta=True, flags=cb.FLAG_TENSOR)
    assert torch.equal(c_def, c_tc) and torch.equal(b_def, b_tc) and torch.equal(d_def, d_tc)
    assert_cost_close(c_def.cpu().numpy(), ref)
    cam, proj = rand_pair(H, W, seed=9)
    c_def = cb.cost_volume(dev(cam), dev(proj), D, k)
    c_tc = cb.cost_volume(dev(cam), dev(proj), D, k, flags=cb.FLAG_TENSOR)
    assert not torch.equal(c_def, c_tc)
    assert_cost_close(c_def.cpu().numpy(), c_tc.cpu().numpy())


# (the 1 x 1 image is left out: its gradient is mathematically zero, so there is no scale to compare rounding noise with)
@pytest.mark.parametrize("H,W,D,k", [c for c in TENSOR_CASES if c[2] <= 540 and c[0] * c[1] > 1])
def test_tensor_core_backward_vs_oracle(H, W, D, k):
    cam, proj = rand_pair(H, W, seed=H * 1000 + W + 2)
    gb = np.random.RandomState(H + W).randn(H, W, D).astype(np.float32)
    gref = ref_port.backward_banded(gb, cam, proj, k)
    grad = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=cb.FLAG_TENSOR)
    assert_grad_close(grad.cpu().numpy(), gref)
    again = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=cb.FLAG_TENSOR)
    assert torch.equal(grad, again)                       # deterministic: no atomics anywhere


def test_tensor_core_backward_batched_equals_per_pair_and_autograd():
    B, H, W, D, k = 3, 50, 300, 128, 5
    cam, proj = rand_pair(H, W, seed=78, B=B)
    g = torch.from_numpy(np.random.RandomState(1).randn(B, H, W, D).astype(np.float32)).cuda()
    grad = cb.backward(g, dev(cam), dev(proj), k, D, flags=cb.FLAG_TENSOR)
    for b in range(B):
        g1 = cb.backward(g[b].contiguous(), dev(cam[b]), dev(proj[b]), k, D, flags=cb.FLAG_TENSOR)
        assert torch.equal(grad[b], g1)
    c = dev(cam).requires_grad_(True)
    cb.cost_volume(c, dev(proj), D, k, flags=cb.FLAG_TENSOR).backward(g)
    assert torch.equal(c.grad, grad)
    with pytest.raises(RuntimeError):
        cb.backward(torch.zeros(16, 40, 544, device="cuda"), dev(cam[0][:16, :40].copy()), dev(proj[0][:16, :40].copy()), k, 544,
                    flags=cb.FLAG_TENSOR)


def test_verdict_hands_low_texture_backward_to_the_tensor_core_kernel():
    H, W, D, k = 64, 420, 192, 5
    rng = np.random.RandomState(6)
    xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    yy = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    scene = lambda sh: 0.5 + 0.4 * np.sin((xx + sh) * 0.01) * np.cos(yy * 0.02)
    cam = np.ascontiguousarray(scene(0) + 0.02 * (rng.rand(H, W) - 0.5), np.float32)
    proj = np.ascontiguousarray(scene(40) + 0.02 * (rng.rand(H, W) - 0.5), np.float32)
    gb = rng.randn(H, W, D).astype(np.float32)
    g_def = cb.backward(dev(gb), dev(cam), dev(proj), k, D)
    g_tc = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=cb.FLAG_TENSOR)
    assert torch.equal(g_def, g_tc)
    gref = ref_port.backward_banded(gb, cam, proj, k)
    truth = zo.camera_grad_banded_autograd(cam, proj, gb, D, k).numpy()
    assert_grad_close_or_nearer_truth(g_def.cpu().numpy(), gref, truth)
    cam, proj = rand_pair(H, W, seed=10)                  # textured: the sliding-window kernels keep the call
    g_def = cb.backward(dev(gb), dev(cam), dev(proj), k, D)
    g_tc = cb.backward(dev(gb), dev(cam), dev(proj), k, D, flags=cb.FLAG_TENSOR)
    assert not torch.equal(g_def, g_tc)
    assert_grad_close(g_def.cpu().numpy(), g_tc.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# example-level callers fused into the path: confidence mask / masked disparity (examples/verify.py:72-74,
# examples/test.py:78-86), gradient row window (row-band shards), uint8 ingestion (examples/verify.py:138-142)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("D", [0, 48])
def test_fused_confidence_mask_and_masked_disparity(D, flags):
    H, W, k, s0 = 40, 150, 5, 11
    rng = np.random.RandomState(77)
    proj = rng.rand(H, W).astype(np.float32)
    cam = rng.rand(H, W).astype(np.float32)
    cam[:, 60:] = proj[:, 60 - s0:W - s0]                    # right part: a confident match at disparity s0; left: noise
    best, index, mask, mdisp = cb.wta_masked(dev(cam), dev(proj), D, k, threshold=0.6, flags=flags)
    b2, i2 = cb.wta(dev(cam), dev(proj), D, k, flags=flags)
    assert torch.equal(best, b2) and torch.equal(index, i2)
    # oracle: verify.py:72-74 on the oracle's volume; test.py:83-84 for the masked disparity
    ref = ref_port.forward_banded(cam, proj, D, k) if D else ref_port.forward_full(cam, proj, k)
    rbest = torch.from_numpy(np.where(ref == cb.INVALID_COST, -np.inf, ref)).max(dim=-1).values
    rmask = zo.confidence_mask(rbest, 0.6).numpy()
    near = np.abs(rbest.numpy() - 0.6) <= COST_TOL            # the threshold decision may differ within tolerance
    assert np.array_equal(mask.cpu().numpy()[~near], rmask[~near])
    assert 0.3 < mask.mean().item() < 0.9
    disp = index.float() if D else (torch.arange(W, device="cuda")[None, :] - index).float()
    assert torch.equal(mask, (best > 0.6).float())
    assert torch.equal(mdisp, disp * mask)
    inner = (slice(k, H - k), slice(60 + k, W - k))
    assert (mdisp[inner] == s0).all()
    # other thresholds: nothing / everything confident
    _, _, m_hi, d_hi = cb.wta_masked(dev(cam), dev(proj), D, k, threshold=2.0, flags=flags)
    _, _, m_lo, _ = cb.wta_masked(dev(cam), dev(proj), D, k, threshold=-3.0, flags=flags)
    assert m_hi.sum().item() == 0 and d_hi.abs().sum().item() == 0 and m_lo.sum().item() == H * W


@pytest.mark.parametrize("flags", FLAGS + [cb.FLAG_TENSOR])
def test_backward_rows_equals_zero_padded_gradient(flags):
    """custma_backward_rows: a gradient that exists on volume rows [r0,r1) only, without the zero padding."""
    B, H, W, D, k = 2, 70, 260, 64, 5
    cam, proj = rand_pair(H, W, 88, B=B)
    rng = np.random.RandomState(5)
    for r0, r1 in [(0, 3), (31, 52), (66, 70), (0, 70)]:
        g = rng.randn(B, r1 - r0, W, D).astype(np.float32)
        full = np.zeros((B, H, W, D), np.float32)
        full[:, r0:r1] = g
        a = cb.backward(dev(full), dev(cam), dev(proj), k, D, flags=flags)
        b = cb.backward(dev(g), dev(cam), dev(proj), k, D, flags=flags, rows=(r0, r1))
        assert torch.equal(a, b), (r0, r1)
    with pytest.raises(RuntimeError):
        cb.backward(dev(g), dev(cam), dev(proj), k, D, rows=(5, 5))
    with pytest.raises(RuntimeError):
        cb.backward(dev(g), dev(cam), dev(proj), k, D, rows=(3, 80))


@pytest.mark.parametrize("flags", FLAGS)
def test_backward_prepared_ahead_of_time_gives_the_same_bits(flags):
    """custma_backward_prepare on a second stream and workspace while the forward runs, then custma_backward with
    CUSTMA_FLAG_PREPARED: bit-identical to the one-call backward (sliding-window and direct kernels, and the row window)."""
    from custereomatching_b200 import binding
    B, H, W, D, k = 2, 50, 210, 64, 5
    cam, proj = rand_pair(H, W, seed=31, B=B)
    cam, proj = dev(cam), dev(proj)
    g = dev(np.random.RandomState(32).randn(B, H, W, D).astype(np.float32))
    want = cb.backward(g, cam, proj, k, D, flags=flags)
    wsb = binding.backward_workspace_bytes(B, H, W, D, k, flags)
    assert wsb == binding.backward_workspace_bytes(B, H, W, D, k, flags | binding.FLAG_PREPARED)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    side, main = torch.cuda.Stream(), torch.cuda.current_stream()
    side.wait_stream(main)
    binding.backward_prepare(cam.data_ptr(), proj.data_ptr(), B, H, W, D, k, flags, ws.data_ptr(), wsb, side.cuda_stream)
    cost, _, _ = cb.forward(cam, proj, D, k, want_cost=True, want_wta=False, flags=flags)   # meanwhile, on the main stream
    main.wait_stream(side)
    got = torch.empty(B, H, W, device="cuda")
    binding.backward(g.data_ptr(), cam.data_ptr(), proj.data_ptr(), got.data_ptr(), B, H, W, D, k, flags | binding.FLAG_PREPARED,
                     ws.data_ptr(), wsb, main.cuda_stream)
    assert torch.equal(got, want)
    # the same preparation serves a second backward (another gradient, a row window)
    g2 = dev(np.random.RandomState(33).randn(B, 20, W, D).astype(np.float32))
    got2 = torch.empty(B, H, W, device="cuda")
    binding.backward_rows(g2.data_ptr(), cam.data_ptr(), proj.data_ptr(), got2.data_ptr(), B, H, W, D, k, 10, 30,
                          flags | binding.FLAG_PREPARED, ws.data_ptr(), wsb, main.cuda_stream)
    assert torch.equal(got2, cb.backward(g2, cam, proj, k, D, flags=flags, rows=(10, 30)))
    # the forward's counterpart: custma_forward_prepare + CUSTMA_FLAG_PREPARED, cost and WTA
    cost0, best0, disp0 = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True, flags=flags)
    wsf_b = binding.forward_workspace_bytes(B, H, W, D, k, flags)
    wsf = torch.empty(wsf_b, dtype=torch.uint8, device="cuda")
    side.wait_stream(main)
    binding.forward_prepare(cam.data_ptr(), proj.data_ptr(), B, H, W, D, k, flags, wsf.data_ptr(), wsf_b, side.cuda_stream)
    main.wait_stream(side)
    cost1, best1, disp1 = torch.empty_like(cost0), torch.empty_like(best0), torch.empty_like(disp0)
    binding.forward(cam.data_ptr(), proj.data_ptr(), cost1.data_ptr(), best1.data_ptr(), disp1.data_ptr(), B, H, W, D, k,
                    flags | binding.FLAG_PREPARED, wsf.data_ptr(), wsf_b, main.cuda_stream)
    assert torch.equal(cost1, cost0) and torch.equal(best1, best0) and torch.equal(disp1, disp0)
    # the host layer: explicit handle, and the autograd function (which prepares during its forward)
    prep = cb.prepare_backward(cam, proj, k, D, flags=flags)
    assert torch.equal(cb.backward(g, cam, proj, k, D, flags=flags, prepared=prep), want)
    with pytest.raises(RuntimeError, match="prepared belongs to"):
        cb.backward(g, proj, cam, k, D, flags=flags, prepared=prep)
    cam_req = cam.clone().requires_grad_(True)
    (cb.cost_volume(cam_req, proj, D, k, flags) * g).sum().backward()
    assert torch.equal(cam_req.grad, cb.backward(g, cam_req.detach(), proj, k, D, flags=flags))


def test_autograd_prepares_the_backward_of_large_volumes():
    """From 32 Mcell on the autograd function starts the backward's preparation during its forward; same bits."""
    H, W, D, k = 160, 640, 320, 5
    cam, proj = rand_pair(H, W, seed=41)
    cam, proj = dev(cam), dev(proj)
    cam_req = cam.clone().requires_grad_(True)
    cost = cb.cost_volume(cam_req, proj, D, k)
    assert cost.grad_fn.prepared is not None
    g = torch.randn(H, W, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    (cost * g).sum().backward()
    assert torch.equal(cam_req.grad, cb.backward(g, cam, proj, k, D))


def test_uint8_ingestion():
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(2, 37, 53, 3)).astype(np.uint8)
    for ch in range(3):
        out = cb.ingest_u8(dev(img), channel=ch)
        assert out.dtype == torch.float32 and out.shape == (2, 37, 53)
        assert np.array_equal(out.cpu().numpy(), img[..., ch].astype(np.float32) * np.float32(1.0 / 255.0))
    one = cb.ingest_u8(dev(img[0]), channel=2, scale=1.0)
    assert np.array_equal(one.cpu().numpy(), img[0, ..., 2].astype(np.float32))
    gray = cb.ingest_u8(dev(img[0, ..., 0].copy()))
    assert np.array_equal(gray.cpu().numpy(), img[0, ..., 0].astype(np.float32) * np.float32(1.0 / 255.0))
    # the ingested planes feed the kernels directly (verify.py:149 takes channel 0 of both images)
    cam, proj = cb.ingest_u8(dev(img[0]), 0), cb.ingest_u8(dev(img[1]), 0)
    c = cb.cost_volume(cam, proj, 16, 5)
    assert_cost_close(c.cpu().numpy(), ref_port.forward_banded(cam.cpu().numpy(), proj.cpu().numpy(), 16, 5))
    with pytest.raises(RuntimeError):
        cb.ingest_u8(dev(img).float())
    with pytest.raises(RuntimeError):
        cb.ingest_u8(dev(img), channel=3)


def test_misaligned_gradient_view_is_copied_not_faulted():
    """a contiguous view at an odd storage offset reaches the 128-bit kernels only through a copy; the C ABI
    itself refuses the raw pointer with an error code instead of faulting."""
    from custereomatching_b200 import binding
    H, W, D, k = 20, 64, 32, 5
    cam, proj = rand_pair(H, W, 17)
    gflat = torch.randn(H * W * D + 1, device="cuda")
    gview = gflat[1:].view(H, W, D)                       # contiguous, data_ptr % 16 == 4
    assert gview.is_contiguous() and gview.data_ptr() % 16 != 0
    a = cb.backward(gview, dev(cam), dev(proj), k, D)
    b = cb.backward(gview.clone(), dev(cam), dev(proj), k, D)
    assert torch.equal(a, b)
    nbytes = binding.backward_workspace_bytes(1, H, W, D, k, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(H, W, device="cuda")
    with pytest.raises(RuntimeError, match="aligned"):
        binding.backward(gview.data_ptr(), dev(cam).data_ptr(), dev(proj).data_ptr(), out.data_ptr(), 1, H, W, D, k, 0,
                         ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------------------
# fused differentiable disparity head: soft-argmax (examples/verify.py:31-39, beta = 50) times the
# confidence mask (:72-74), forward and backward without a volume
# ---------------------------------------------------------------------------------------------------------------
HEAD_BETA = 50.0
# The head amplifies cost differences by beta: a cost error dc moves a softmax weight by beta*dc (relative) and the soft
# disparity by at most beta * dc * MAD, MAD = the weights' mean absolute deviation of s (first-order perturbation of
# sum_s w_s s).  With the cost tolerance 1e-5 that is 5e-4 * MAD pixels; 1e-4 px of slack for the exp2 approximation.
def assert_soft_close(soft, ref, mad, what="soft disparity"):
    err = np.abs(np.asarray(soft, np.float64) - ref)
    bound = HEAD_BETA * COST_TOL * mad + 1e-4
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} pixels out of bound, worst {err[bad].max():.3e} (bound {bound[bad].min():.3e})"


HEAD_CASES = [  # H, W, D, k
    (24, 90, 48, 5), (31, 260, 192, 5), (17, 77, 20, 3), (12, 50, 0, 5), (40, 300, 100, 5), (9, 140, 256, 5),
]


def _matched_pair(H, W, seed, s_true=9, noise=0.02):
    rng = np.random.RandomState(seed)
    proj = rng.rand(H, W).astype(np.float32)
    cam = rng.rand(H, W).astype(np.float32)
    cam[:, W // 3:] = proj[:, W // 3 - s_true:W - s_true] + noise * rng.randn(H, W - W // 3).astype(np.float32)
    return cam.astype(np.float32), proj


@pytest.mark.parametrize("H,W,D,k", HEAD_CASES)
def test_fused_head_forward_backward_vs_oracle(H, W, D, k):
    cam, proj = _matched_pair(H, W, seed=H * 100 + W)
    gd = np.random.RandomState(3).randn(H, W).astype(np.float32)
    ref_soft, ref_best, ref_mask, mad, ref_grad = zo.soft_disparity_head(cam, proj, D, k, HEAD_BETA, 0.6, soft_grad=gd)
    camt = dev(cam).requires_grad_(True)
    soft, best, index, mask = cb.soft_disparity(camt, dev(proj), D, k, beta=HEAD_BETA, threshold=0.6)
    b2, i2, m2, _ = cb.wta_masked(dev(cam), dev(proj), D, k, threshold=0.6)
    assert torch.equal(best, b2) and torch.equal(index, i2) and torch.equal(mask, m2)       # same WTA as the plain path
    near = np.abs(ref_best.numpy() - 0.6) <= COST_TOL
    assert np.array_equal(mask.cpu().numpy()[~near], ref_mask.numpy()[~near])
    ok = ~near
    assert_soft_close(soft.detach().cpu().numpy()[ok], ref_soft.numpy()[ok], mad.numpy()[ok])
    assert 0.2 < mask.mean().item() < 0.95
    soft.backward(dev(gd))
    # gradient: the same beta amplification (weights to ~beta * 1e-6 relative in fp32) -> 1e-4 of the gradient scale
    g = camt.grad.cpu().numpy()
    scale = np.abs(ref_grad.numpy()).max()
    assert np.isfinite(g).all()
    assert np.abs(g - ref_grad.numpy()).max() <= 1e-4 * scale, np.abs(g - ref_grad.numpy()).max() / scale


def test_fused_head_equals_unfused_gpu_path():
    """Fusion check in fp32 on the device: soft-argmax + mask in torch on the materialised volume, autograd through
    custma's own backward, against the fused head (which never writes the volume)."""
    H, W, D, k = 60, 420, 192, 5
    cam, proj = _matched_pair(H, W, seed=5, s_true=23)
    gd = torch.randn(H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    c1 = dev(cam).requires_grad_(True)
    vol = cb.cost_volume(c1, dev(proj), D, k)
    valid = (torch.arange(W, device="cuda")[:, None] - torch.arange(D, device="cuda")[None, :]) >= 0
    logits = torch.where(valid[None], vol * HEAD_BETA, torch.full_like(vol, float("-inf")))
    w = torch.softmax(logits, dim=-1)
    soft_ref = (w * torch.arange(D, device="cuda", dtype=torch.float32)).sum(-1)
    best_ref = torch.where(valid[None], vol, torch.full_like(vol, float("-inf"))).max(dim=-1).values.detach()
    mask_ref = (best_ref > 0.6).float()
    (soft_ref * mask_ref * gd).sum().backward()
    c2 = dev(cam).requires_grad_(True)
    soft, best, index, mask = cb.soft_disparity(c2, dev(proj), D, k, beta=HEAD_BETA, threshold=0.6)
    assert torch.equal(mask, mask_ref) and torch.equal(best, best_ref)
    assert float((soft - soft_ref * mask_ref).abs().max()) <= 2e-3
    soft.backward(gd)
    scale = float(c1.grad.abs().max())
    assert float((c2.grad - c1.grad).abs().max()) <= 1e-4 * scale
    # deterministic, and batched == per pair
    c3 = dev(cam).requires_grad_(True)
    s3, *_ = cb.soft_disparity(c3, dev(proj), D, k, beta=HEAD_BETA, threshold=0.6)
    s3.backward(gd)
    assert torch.equal(s3, soft) and torch.equal(c3.grad, c2.grad)
    cb2 = torch.stack([dev(cam), dev(cam).flip(0)]).contiguous().requires_grad_(True)
    pb2 = torch.stack([dev(proj), dev(proj).flip(0)]).contiguous()
    sb, *_ = cb.soft_disparity(cb2, pb2, D, k, beta=HEAD_BETA, threshold=0.6)
    sb.backward(torch.stack([gd, gd.flip(0)]))
    assert torch.equal(sb[0], soft) and torch.equal(cb2.grad[0], c2.grad)


def test_fused_head_on_low_texture_input_uses_the_fallback_partials():
    """A smooth scene flags every tile of the sliding-window path: the head's partials then come from the per-cell
    fallback kernels (no tensor-core hand-over in head mode) and must still match the oracle."""
    H, W, D, k = 24, 200, 64, 5
    rng = np.random.RandomState(8)
    xx = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    yy = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    scene = lambda sh: 0.5 + 0.4 * np.sin((xx + sh) * 0.05) * np.cos(yy * 0.1)
    cam = np.ascontiguousarray(scene(0) + 0.01 * (rng.rand(H, W) - 0.5), np.float32)
    proj = np.ascontiguousarray(scene(12) + 0.01 * (rng.rand(H, W) - 0.5), np.float32)
    gd = rng.randn(H, W).astype(np.float32)
    ref_soft, ref_best, ref_mask, mad, ref_grad = zo.soft_disparity_head(cam, proj, D, k, HEAD_BETA, -2.0, soft_grad=gd)
    camt = dev(cam).requires_grad_(True)
    soft, best, index, mask = cb.soft_disparity(camt, dev(proj), D, k, beta=HEAD_BETA, threshold=-2.0)
    assert mask.min().item() == 1.0
    assert_cost_close(best.cpu().numpy(), ref_best.numpy(), what="best")
    assert_soft_close(soft.detach().cpu().numpy(), ref_soft.numpy(), mad.numpy())
    soft.backward(dev(gd))
    truth = ref_grad.numpy()
    # low-texture: the fp32 reference arithmetic itself is ~1e-4 of scale from fp64 here (tiny denominators, times beta)
    assert np.abs(camt.grad.cpu().numpy() - truth).max() <= 5e-4 * np.abs(truth).max()


def test_fused_head_refuses_what_it_cannot_do():
    cam, proj = rand_pair(16, 40, seed=1)
    with pytest.raises(RuntimeError):
        cb.soft_disparity(dev(cam), dev(proj), 16, 7)            # no backward fast path for k = 7
    with pytest.raises(RuntimeError):
        cb.soft_disparity(dev(cam), dev(proj), 16, 5, beta=-1.0)


# ---------------------------------------------------------------------------------------------------------------
# projector gradient: the reference returns None for it (custma/stereo_matching_wrapper.py:33)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,D,k", [(14, 40, 16, 5), (9, 33, 0, 5), (12, 30, 8, 3), (10, 28, 12, 4), (8, 36, 0, 7),
                                     (20, 64, 64, 5), (6, 25, 25, 15)])
def test_projector_gradient_vs_oracle_autograd(H, W, D, k):
    cam, proj = rand_pair(H, W, seed=H * 31 + W)
    C = D if D > 0 else W
    g = np.random.RandomState(4).randn(H, W, C).astype(np.float32)
    ref = zo.projector_grad_autograd(cam, proj, g, D, k).numpy()
    out = cb.backward_projector(dev(g), dev(cam), dev(proj), k, D)
    assert out.shape == (H, W)
    assert_grad_close(out.cpu().numpy(), ref)
    again = cb.backward_projector(dev(g), dev(cam), dev(proj), k, D)
    assert torch.equal(out, again)                                   # deterministic: no atomics
    # both routes: the default (banded, k = 3 / 5: sliding-window kernels on the mirrored problem) and the direct kernels
    direct = cb.backward_projector(dev(g), dev(cam), dev(proj), k, D, flags=cb.FLAG_DIRECT)
    assert_grad_close(direct.cpu().numpy(), ref)
    # symmetry that the kernel exploits: for the reference-shaped volume, the projector gradient is the camera gradient
    # of the transposed problem
    if D == 0:
        gt = np.ascontiguousarray(g.transpose(0, 2, 1))
        sym = cb.backward(dev(gt), dev(proj), dev(cam), k, 0, flags=cb.FLAG_DIRECT)
        assert_grad_close(out.cpu().numpy(), sym.cpu().numpy(), 1e-6)


def test_projector_gradient_fast_path_on_larger_shapes():
    """Several column tiles, row bands and a batch through the mirrored sliding-window route, against the direct kernels
    (themselves checked against fp64 autograd above) and the oracle on one pair."""
    for B, H, W, D, k in [(2, 90, 300, 64, 5), (1, 70, 420, 192, 3), (3, 40, 131, 96, 5)]:
        cam, proj = rand_pair(H, W, seed=H + W, B=B)
        g = torch.from_numpy(np.random.RandomState(2).randn(B, H, W, D).astype(np.float32)).cuda()
        fast = cb.backward_projector(g, dev(cam), dev(proj), k, D)
        direct = cb.backward_projector(g, dev(cam), dev(proj), k, D, flags=cb.FLAG_DIRECT)
        assert_grad_close(fast.cpu().numpy(), direct.cpu().numpy())
        valid = (np.arange(W)[:, None] - np.arange(D)[None, :]) >= 0
        ref = zo.projector_grad_autograd(cam[0], proj[0], g[0].cpu().numpy() * valid, D, k).numpy()
        assert_grad_close(fast[0].cpu().numpy(), ref)
        one = cb.backward_projector(g[0].contiguous(), dev(cam[0]), dev(proj[0]), k, D)
        assert torch.equal(one, fast[0])


def test_projector_gradient_through_autograd_and_batches():
    B, H, W, D, k = 2, 16, 48, 24, 5
    cam, proj = rand_pair(H, W, seed=9, B=B)
    wgt = torch.from_numpy(np.random.RandomState(5).randn(B, H, W, D).astype(np.float32)).cuda()
    c, p = dev(cam).requires_grad_(True), dev(proj).requires_grad_(True)
    vol = cb.cost_volume(c, p, D, k)
    valid = (torch.arange(W, device="cuda")[:, None] - torch.arange(D, device="cuda")[None, :]) >= 0
    (vol * wgt * valid).sum().backward()
    assert c.grad is not None and p.grad is not None
    for b in range(B):
        gm = (wgt[b] * valid).cpu().numpy()
        assert_grad_close(p.grad[b].cpu().numpy(), zo.projector_grad_autograd(cam[b], proj[b], gm, D, k).numpy())
        assert_grad_close(c.grad[b].cpu().numpy(), zo.camera_grad_banded_autograd(cam[b], proj[b], gm, D, k).numpy())
    # the drop-in Function keeps the reference's contract: no projector gradient
    c2, p2 = dev(cam[0]).requires_grad_(True), dev(proj[0]).requires_grad_(True)
    cv = custma.stereo_matching(c2, p2, 0, k)
    cv.backward(torch.ones_like(cv))
    assert p2.grad is None


def test_host_submit_u8_equals_ingest_then_device_path():
    """custma_host_submit_u8: 8-bit interleaved host images in (RGB camera, channel 0; gray projector), same results as
    ingesting on the device and calling the kernels directly."""
    from custereomatching_b200 import binding
    B, H, W, D, k = 2, 40, 200, 64, 5
    rng = np.random.RandomState(12)
    cam_u8 = torch.from_numpy(rng.randint(0, 256, size=(B, H, W, 3)).astype(np.uint8)).pin_memory()
    proj_u8 = torch.from_numpy(rng.randint(0, 256, size=(B, H, W)).astype(np.uint8)).pin_memory()
    g = dev(rng.randn(B, H, W, D).astype(np.float32))
    hb, hi, hg = torch.empty(B, H, W).pin_memory(), torch.empty(B, H, W, dtype=torch.int32).pin_memory(), torch.empty(B, H, W).pin_memory()
    t = binding.host_submit_u8(cam_u8.data_ptr(), 3, 0, proj_u8.data_ptr(), 1, 0, 1.0 / 255.0, hb.data_ptr(), hi.data_ptr(),
                               hg.data_ptr(), 0, g.data_ptr(), B, H, W, D, k)
    binding.host_wait(t)
    cam, proj = cb.ingest_u8(cam_u8.cuda(), 0), cb.ingest_u8(proj_u8.cuda().unsqueeze(-1))   # [B,H,W,1]: a 3-D tensor is [H,W,Ch]
    best, disp = cb.wta(cam, proj, D, k)
    grad = cb.backward(g, cam, proj, k, D)
    assert torch.equal(hb.cuda(), best) and torch.equal(hi.cuda(), disp) and torch.equal(hg.cuda(), grad)
    binding.host_release()


def test_custma_package_exports_the_fused_callers():
    """The drop-in package also carries the example-level callers (mask, soft disparity head, projector gradient)."""
    H, W, D, k = 24, 90, 32, 5
    cam, proj = _matched_pair(H, W, seed=4)
    b, i, m, md = custma.stereo_matching_masked(dev(cam), dev(proj), D, k)
    b2, i2, m2, md2 = cb.wta_masked(dev(cam), dev(proj), D, k)
    assert torch.equal(b, b2) and torch.equal(i, i2) and torch.equal(m, m2) and torch.equal(md, md2)
    c = dev(cam).requires_grad_(True)
    soft, best, idx, mask = custma.stereo_matching_soft_disparity(c, dev(proj), D, k)
    soft.sum().backward()
    assert c.grad is not None and torch.isfinite(c.grad).all() and torch.equal(best, b)
    g = torch.randn(H, W, D, device="cuda")
    pg = custma.stereo_matching_projector_grad(g, dev(cam), dev(proj), D, k)
    assert torch.equal(pg, cb.backward_projector(g, dev(cam), dev(proj), k, D))
