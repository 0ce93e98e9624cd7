"""Worker of tests/test_nccl_gpu.py: one process per GPU under torch.distributed.run, NCCL backend.

Every rank computes the unsharded result on its own GPU and compares it with what the two partitionings of
custereomatching_b200.sharding deliver after their NCCL gathers (batch sharding: bit-exact; row-band sharding:
costs / WTA bit-exact on the owned rows, camera gradient to rounding because the halo rows are summed in a different
order).  Exit code 0 = every rank agreed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import custereomatching_b200 as cb  # noqa: E402
from custereomatching_b200 import sharding as sh  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    k = 5
    # ---- batch sharding (BASELINE.json configs[3] in small): ragged slices on purpose
    B, H, W, D = 2 * world + 1, 60, 300, 64
    rng = np.random.RandomState(123)
    cam = torch.from_numpy(rng.rand(B, H, W).astype(np.float32)).to(dev)
    proj = torch.from_numpy(rng.rand(B, H, W).astype(np.float32)).to(dev)
    g = torch.from_numpy(rng.randn(B, H, W, D).astype(np.float32)).to(dev)
    _, best, disp = cb.forward(cam, proj, D, k, want_cost=False, want_wta=True)
    grad = cb.backward(g, cam, proj, k, D)
    b0, b1 = sh.batch_slice(B, rank, world)
    gb, gd, gg = sh.batch_sharded_step(cam[b0:b1].contiguous(), proj[b0:b1].contiguous(), D, k,
                                       cost_volume_grad_fn=lambda cost: g[b0:b1].contiguous(), global_batch=B)
    assert torch.equal(gb, best) and torch.equal(gd, disp) and torch.equal(gg, grad), "batch sharding differs"
    # ---- row-band sharding (configs[4] in small)
    H, W, D = 40 * world + 3, 520, 192
    cam = torch.from_numpy(rng.rand(H, W).astype(np.float32)).to(dev)
    proj = torch.from_numpy(rng.rand(H, W).astype(np.float32)).to(dev)
    g = torch.from_numpy(rng.randn(H, W, D).astype(np.float32)).to(dev)
    cost, best, disp = cb.forward(cam, proj, D, k, want_cost=True, want_wta=True)
    grad = cb.backward(g, cam, proj, k, D)
    rb, rd, rg = sh.row_band_sharded_step(cam, proj, D, k, cost_volume_grad_fn=lambda c, band: g[band.h0:band.h1].contiguous())
    # a crop has its own row bands and pivots: costs agree to rounding (not bit for bit), disparities outside near-ties
    assert float((rb - best).abs().max()) <= 2e-6, "row-band best differs"
    top2 = torch.topk(cost, 2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-5
    assert torch.equal(rd[clear], disp[clear]), "row-band disparity differs outside near-ties"
    err = float((rg - grad).abs().max() / grad.abs().max())
    assert err <= 1e-5, f"row-band gradient differs by {err:.2e} of scale"
    # every rank holds the same gathered bits
    check = torch.stack([rb.double().sum(), rd.double().sum(), rg.double().sum()])
    ref = check.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(check, ref), "ranks hold different gathered results"
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_PARITY_OK", flush=True)


if __name__ == "__main__":
    main()
