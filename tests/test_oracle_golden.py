"""Pins the oracle (oracle/zncc_oracle.py fp64 torch, oracle/ref_port.c fp32 C) to the golden vectors produced by
the UNMODIFIED reference CUDA extension on a B200 (tests/golden/make_golden_from_reference.py)."""
import numpy as np
import pytest
import torch

from conftest import (COST_TOL, GRAD_TOL, assert_cost_close, assert_grad_close, golden_manifest, golden_small_names,
                      load_golden)
from oracle import ref_port
from oracle import zncc_oracle as zo

SMALL = golden_small_names()
# On the low-contrast fixtures the reference's own fp32 gradient is 2.6e-5 .. 8.7e-5 of scale away from fp64 truth
# (measured), so the fp64 oracle is pinned there with a wider band; the fp32 C port, which
# follows the reference's arithmetic order, stays inside GRAD_TOL everywhere.
FP64_GRAD_TOL = {"smooth_k5_12x40": 5e-5, "smooth_k15_18x40": 2e-4, "edges_k5_12x32": 5e-5}


@pytest.mark.parametrize("name", SMALL)
def test_c_port_forward_bit_exact(name):
    g = load_golden(name)
    cv = ref_port.forward_full(g["camera"], g["projector"], int(g["kernel_size"]))
    assert np.array_equal(cv, g["cost_volume"]), "C restatement differs from the reference extension's bits"


@pytest.mark.parametrize("name", SMALL)
def test_c_port_backward(name):
    g = load_golden(name)
    grad = ref_port.backward_full(g["cost_volume_grad"], g["camera"], g["projector"], int(g["kernel_size"]))
    assert_grad_close(grad, g["camera_grad"], GRAD_TOL)


@pytest.mark.parametrize("name", SMALL)
def test_fp64_oracle_forward(name):
    g = load_golden(name)
    cv = zo.cost_volume_full(g["camera"], g["projector"], int(g["kernel_size"])).numpy()
    assert_cost_close(cv, g["cost_volume"], COST_TOL)


@pytest.mark.parametrize("name", SMALL)
def test_fp64_oracle_backward(name):
    g = load_golden(name)
    k = int(g["kernel_size"])
    formula = zo.camera_grad_full_kernel_formula(g["camera"], g["projector"], g["cost_volume_grad"], k).numpy()
    assert_grad_close(formula, g["camera_grad"], FP64_GRAD_TOL.get(name, GRAD_TOL))
    if k % 2 == 1:  # verify.py's torch path is only defined for odd k
        auto = zo.camera_grad_full_autograd(g["camera"], g["projector"], g["cost_volume_grad"], k).numpy()
        assert_grad_close(auto, formula, 1e-12, what="autograd vs kernel formula")


def test_constant_camera_known_answer():
    # analytic KAT: a patch of constant camera values has ex2 = exy = 0 -> cost = eps / sqrt(eps) = 1e-4
    g = load_golden("const_k5_8x12")
    k = int(g["kernel_size"])
    r = k // 2
    cv = ref_port.forward_full(g["camera"], g["projector"], k)
    interior = cv[r:-r, r:-r, :]
    assert np.allclose(interior, 1e-4, rtol=1e-6, atol=0)
    assert np.allclose(g["cost_volume"][r:-r, r:-r, :], 1e-4, rtol=1e-6, atol=0)


def test_cfg1_fixture():
    """BASELINE.json configs[0]: 320x240, 5x5 window: volume sample, WTA and camera gradient of the reference."""
    g = load_golden("cfg1_rand_k5_240x320")
    man = golden_manifest()["cfg1"]
    k = int(g["kernel_size"])
    cam, proj = g["camera"], g["projector"]
    cv = ref_port.forward_full(cam, proj, k)
    assert np.array_equal(cv[::16, ::16, :], g["cost_sub"])
    best, corr = ref_port.wta_full(cv)
    assert np.array_equal(best, g["best"])
    assert man["min_top2_gap"] > 0  # no exact ties in this fixture -> indices must match everywhere
    assert np.array_equal(corr, g["argmax"].astype(np.int32))
    best64, corr64, disp64 = zo.wta_full(zo.cost_volume_full(cam, proj, k))
    gap = zo.top2_gap(zo.cost_volume_full(cam, proj, k)).numpy()
    sel = gap > COST_TOL
    assert sel.mean() > 0.99
    assert np.array_equal(corr64.numpy()[sel], g["argmax"].astype(np.int64)[sel])
    assert_cost_close(best64.numpy(), g["best"])
    upstream = (g["grad_u"] * g["grad_v"] + g["grad_bias"]).astype(np.float32)
    grad = ref_port.backward_full(upstream, cam, proj, k)
    assert_grad_close(grad, g["camera_grad"], GRAD_TOL)


@pytest.mark.parametrize("name", ["rand_k5_12x20", "rand_k4_10x16", "shifted_k5_16x32"])
@pytest.mark.parametrize("D", [1, 7, 16, 40])
def test_banded_definition(name, D):
    """band[h,w,s] = full[h,w,w-s]; invalid cells hold INVALID_COST; gradients route through the same map."""
    g = load_golden(name)
    k = int(g["kernel_size"])
    cam, proj = g["camera"], g["projector"]
    H, W = cam.shape
    full = torch.from_numpy(g["cost_volume"])
    band_ref = zo.band_from_full(full, D).numpy()
    band_c = ref_port.forward_banded(cam, proj, D, k)
    assert np.array_equal(band_c, band_ref)
    band64 = zo.cost_volume_banded(cam, proj, D, k).numpy()
    assert_cost_close(band64, band_ref)
    assert (band_ref[:, 0, 1:] == zo.INVALID_COST).all()
    # WTA on the band == WTA of the full volume restricted to the band, ties to the lowest projector column
    best_c, disp_c = ref_port.wta_banded(band_c)
    best_t, disp_t = zo.wta_banded(torch.from_numpy(band_ref))
    assert np.array_equal(best_c, best_t.numpy())
    assert np.array_equal(disp_c, disp_t.numpy().astype(np.int32))
    # gradient: banded upstream gradient scattered to the full volume gives the same camera gradient
    rng = np.random.RandomState(5)
    gb = rng.randn(H, W, D).astype(np.float32)
    gfull = zo.full_from_band_grad(torch.from_numpy(gb), W).numpy()
    grad_full = ref_port.backward_full(gfull, cam, proj, k)
    grad_band = ref_port.backward_banded(gb, cam, proj, k)
    assert_grad_close(grad_band, grad_full, 1e-6)
    if k % 2 == 1:
        auto = zo.camera_grad_banded_autograd(cam, proj, gb, D, k).numpy()
        assert_grad_close(grad_band, auto, GRAD_TOL)


def test_wta_tie_break_lowest_projector_column():
    vol = torch.zeros(1, 4, 4)
    vol[0, 2] = torch.tensor([0.5, 0.9, 0.9, 0.1])
    best, corr, _ = zo.wta_full(vol)
    assert corr[0, 2].item() == 1 and best[0, 2].item() == pytest.approx(0.9)
    band = torch.full((1, 4, 4), -2.0)
    band[0, 3] = torch.tensor([0.9, 0.2, 0.9, 0.1])          # at w = 3, s = 0 and s = 2 tie -> largest disparity wins
    band[0, 0, 0] = band[0, 1, 0] = band[0, 2, 0] = 0.3
    _, disp = zo.wta_banded(band)
    assert disp[0, 3].item() == 2
    b, d = ref_port.wta_banded(band.numpy())
    assert d[0, 3] == 2 and d[0, 0] == 0
    b, c = ref_port.wta_full(vol.numpy())
    assert c[0, 2] == 1 and c[0, 0] == 0


def test_mask_and_soft_argmax():
    best = torch.tensor([0.59, 0.6, 0.61])
    assert zo.confidence_mask(best).tolist() == [0.0, 0.0, 1.0]       # strict > 0.6 (examples/verify.py:74)
    vol = torch.zeros(1, 1, 8, dtype=torch.float64)
    vol[0, 0, 5] = 1.0
    assert abs(zo.soft_argmax(vol).item() - 5.0) < 1e-6                 # beta = 50 -> essentially one-hot
