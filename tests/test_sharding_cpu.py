"""Host-side multi-GPU logic on CPU: partition arithmetic, and the world_size-2 gloo path of the gather / halo
overlap-add, with the CPU oracle standing in for the per-rank kernels (tests may use the oracle; the product never
does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from custereomatching_b200 import sharding as sh
from oracle import ref_port


def test_split_even():
    assert sh.split_even(64, 8) == [(i * 8, i * 8 + 8) for i in range(8)]
    assert sh.split_even(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sh.split_even(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for n, p in [(4320, 8), (375, 7), (1, 1)]:
        parts = sh.split_even(n, p)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("H,k,world", [(4320, 5, 8), (375, 15, 4), (40, 4, 3), (9, 7, 2)])
def test_row_band_halo(H, k, world):
    r = k // 2
    covered = np.zeros(H, int)
    for rank in range(world):
        b = sh.row_band(H, k, rank, world)
        covered[b.h0:b.h1] += 1
        assert b.lo == max(0, b.h0 - r) and b.hi == min(H, b.h1 + k - 1 - r)
        assert 0 <= b.top_halo <= r
    assert (covered == 1).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- batch sharding: 5 pairs over 2 ranks (3 + 2), oracle as the per-rank compute ----
        B, H, W, D, k = 5, 10, 24, 8, 5
        rng = np.random.RandomState(0)
        cam = rng.rand(B, H, W).astype(np.float32)
        proj = rng.rand(B, H, W).astype(np.float32)
        g = rng.randn(B, H, W, D).astype(np.float32)
        b0, b1 = sh.batch_slice(B, rank, world)
        best_l, disp_l, grad_l = [], [], []
        for b in range(b0, b1):
            band = ref_port.forward_banded(cam[b], proj[b], D, k)
            bb, dd = ref_port.wta_banded(band)
            best_l.append(bb); disp_l.append(dd)
            grad_l.append(ref_port.backward_banded(g[b], cam[b], proj[b], k))
        best = sh.all_gather_batch(torch.from_numpy(np.stack(best_l)), B)
        disp = sh.all_gather_batch(torch.from_numpy(np.stack(disp_l)), B)
        grad = sh.all_gather_batch(torch.from_numpy(np.stack(grad_l)), B)
        ok_batch = True
        for b in range(B):
            band = ref_port.forward_banded(cam[b], proj[b], D, k)
            bb, dd = ref_port.wta_banded(band)
            ok_batch &= np.array_equal(best[b].numpy(), bb) and np.array_equal(disp[b].numpy(), dd)
            ok_batch &= np.array_equal(grad[b].numpy(), ref_port.backward_banded(g[b], cam[b], proj[b], k))

        # ---- row-band sharding of one pair: rows split over 2 ranks with a window-radius halo ----
        H, W, D, k = 23, 30, 12, 5
        cam1 = rng.rand(H, W).astype(np.float32)
        proj1 = rng.rand(H, W).astype(np.float32)
        g1 = rng.randn(H, W, D).astype(np.float32)
        band_r = sh.row_band(H, k, rank, world)
        cam_c = sh.crop_rows_with_halo(torch.from_numpy(cam1), band_r).numpy()
        proj_c = sh.crop_rows_with_halo(torch.from_numpy(proj1), band_r).numpy()
        cost_c = ref_port.forward_banded(cam_c, proj_c, D, k)
        bb, dd = ref_port.wta_banded(cost_c)
        best_rows = sh.band_rows_of(torch.from_numpy(bb), band_r)
        disp_rows = sh.band_rows_of(torch.from_numpy(dd), band_r)
        g_c = np.zeros_like(cost_c)
        m0, m1 = sh.band_gradient_mask_rows(band_r)
        g_c[m0:m1] = g1[band_r.h0:band_r.h1]
        grad_c = ref_port.backward_banded(g_c, cam_c, proj_c, k)
        best_full = sh.all_gather_row_bands(best_rows, H, k)
        disp_full = sh.all_gather_row_bands(disp_rows, H, k)
        grad_full = sh.assemble_row_band_gradient(torch.from_numpy(grad_c), H, k)
        whole = ref_port.forward_banded(cam1, proj1, D, k)
        wb, wd = ref_port.wta_banded(whole)
        wg = ref_port.backward_banded(g1, cam1, proj1, k)
        ok_rows = np.array_equal(best_full.numpy(), wb) and np.array_equal(disp_full.numpy(), wd)
        err = float(np.abs(grad_full.numpy() - wg).max() / np.abs(wg).max())
        q.put((rank, bool(ok_batch), bool(ok_rows), err))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_batch_and_row_band():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_batch, ok_rows, err in results:
        assert ok_batch, f"rank {rank}: batch-sharded gather differs from the whole-batch result"
        assert ok_rows, f"rank {rank}: row-band WTA differs from the whole-image result"
        assert err <= 1e-6, f"rank {rank}: row-band gradient differs by {err}"
