#!/usr/bin/env python
"""bench.py - throughput of the ZNCC cost-volume hot path (forward + WTA + backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kitti|cfg1|cfg2|cfg3|cfg5|cfg5band|verify15]
                    [--data rand|shifted|natural] [--scaling weak|strong] [--graph] [--impl b200|reference|reference-cuda]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic stereo pairs: forward cost volume + fused
winner-take-all, then the backward camera gradient for an upstream gradient that is already on the device.
Metric (BASELINE.json): Mpix*disp/s = 1e-6 * cost-volume cells / second, fwd+bwd, whole job over all N GPUs.

Default workload ("kitti"): P pairs per GPU of KITTI size 1242x375, 192 disparities, 5x5 window (BASELINE.json
configs[1]'s shape; at N = 8 and P = 8 it is configs[3], the batch of 64 pairs, batch-sharded).  Weak scaling: the
per-GPU batch is fixed (--scaling strong divides 64 pairs over the ranks instead).  No data-path collective; the
[B,H,W] results travel to rank 0 inside the step when N > 1 (peer copies over NVLink, or NCCL gather).  --workload cfg5
is configs[4]: one 8K pair, row-band sharded.

Rank 0 prints ONE JSON line.  `value` is measured with inputs resident in HBM; `e2e` goes through the C-ABI
host-buffer entry point (custma_host_step) with pinned host images in and host results out, copies inside the timed
region.  `roofline` is for the dominant kernel group, timed live with CUDA events on the launching stream.
`cpu_baseline` is the pure-PyTorch CPU restatement of the reference math (oracle/) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, D, k, default pairs per GPU, description)
    "kitti": (375, 1242, 192, 5, 8, "KITTI-size pairs 1242x375, D=192, k=5 (BASELINE configs[1] shape; 8 pairs/GPU, "
                                    "N=8 is configs[3]'s batch of 64)"),
    "cfg2": (375, 1242, 192, 5, 1, "BASELINE configs[1]: one KITTI-size pair 1242x375, D=192, k=5"),
    "cfg3": (1988, 2880, 256, 5, 1, "BASELINE configs[2]: one Middlebury-full-size pair 2880x1988, D=256, k=5"),
    "cfg1": (240, 320, 64, 5, 1, "BASELINE configs[0]: one 320x240 pair, D=64, k=5"),
    "cfg5band": (544, 7680, 512, 5, 1, "BASELINE configs[4] (one 7680x4320 pair, D=512, k=5, row-band over 8 GPUs): the "
                                      "work of ONE rank - 540 volume rows plus the window-radius halo (sharding.row_band)"),
    "cfg5": (4320, 7680, 512, 5, 1, "BASELINE configs[4]: one 7680x4320 pair, D=512, k=5, row-band sharded over the N GPUs "
                                    "with a window-radius halo (strong scaling: the pair is fixed)"),
    "verify15": (330, 422, 0, 15, 1, "the reference's own example constants (examples/verify.py:10-11): 422x330 pair, "
                                     "kernel_size 15, reference-shaped [H,W,W] volume through the drop-in call"),
}
METRIC = "Mpix*disp/s fwd+bwd (cost-volume cells per second, forward+WTA+backward)"
UNIT = "Mpix*disp/s"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


DATA_FAMILIES = {
    "rand": "uniform random camera and projector (textured: the best case of the conditioning verdict)",
    "shifted": "SURVEY 8(d) realistic variant: projector = uniform random speckle, camera = the projector seen under a "
               "smooth disparity field in [4, D/2) + 1 % noise (WTA has a ground truth)",
    "natural": "natural-like rectified pair: both views cut from one 1/f^1.5 scene 40 columns apart, 1 % sensor noise each "
               "(low local texture: flags the sliding-window verdict, runs on the tensor-core kernels)",
}


def make_pairs(kind, P, H, W, D, seed):
    """P synthetic pairs [P,H,W] fp32 on the host (camera, projector) of the given family."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    if kind == "rand":
        return torch.rand(P, H, W, generator=gen), torch.rand(P, H, W, generator=gen)
    Dd = max(D, 8)
    if kind == "shifted":
        proj = torch.rand(P, H, W, generator=gen)
        yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        d0 = 4 + (Dd / 2 - 4) * (0.5 + 0.5 * torch.sin(yy / 61.0) * torch.cos(xx / 97.0))
        src = (xx - d0).round().long()
        fill = torch.rand(P, H, W, generator=gen)                     # left of the pattern: unrelated texture, not black
        cam = torch.where((src >= 0)[None], torch.gather(proj, 2, src.clamp(0, W - 1)[None].expand(P, H, W)), fill)
        return (cam + 0.01 * torch.randn(P, H, W, generator=gen)).contiguous(), proj
    if kind == "natural":
        Wp = W + 40
        spec = torch.fft.rfft2(torch.randn(P, H, Wp, generator=gen))
        fy = torch.fft.fftfreq(H)[:, None]
        fx = torch.fft.rfftfreq(Wp)[None, :]
        f = torch.sqrt(fx * fx + fy * fy)
        f[0, 0] = 1.0
        scene = torch.fft.irfft2(spec / f ** 1.5, s=(H, Wp))
        lo = scene.amin(dim=(1, 2), keepdim=True)
        hi = scene.amax(dim=(1, 2), keepdim=True)
        scene = (scene - lo) / (hi - lo)
        cam = scene[:, :, 40:40 + W] + 0.01 * torch.randn(P, H, W, generator=gen)
        proj = scene[:, :, 0:W] + 0.01 * torch.randn(P, H, W, generator=gen)
        return cam.contiguous(), proj.contiguous()
    raise SystemExit(f"unknown --data {kind}")


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, nvidia-smi fallback)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            handle = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        handle = h
                        break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._nvml = (pynvml, handle)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _reasons(self, pynvml, handle):
        try:
            get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(get(handle))
        except Exception:
            return
        names = {
            0x0000000000000004: "sw_power_cap", 0x0000000000000008: "hw_slowdown",
            0x0000000000000020: "sw_thermal_slowdown", 0x0000000000000040: "hw_thermal_slowdown",
            0x0000000000000080: "hw_power_brake_slowdown", 0x0000000000000002: "applications_clocks_setting",
            0x0000000000000010: "sync_boost", 0x0000000000000100: "display_clock_setting",
        }
        for bit, name in names.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        pynvml, handle = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                self._reasons(pynvml, handle)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._open()
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(2.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def cpu_torch_port_rate(H, W, D, k, budget_s=12.0, max_rows=None, threads=None):
    """fwd + WTA + bwd of the pure-PyTorch CPU restatement (oracle/zncc_oracle.py) on a bounded sample:
    a band of image rows of ONE pair of the workload's shape.  Returns (Mcells/s, description, threads)."""
    import torch
    from oracle import zncc_oracle as zo
    torch.set_num_threads(threads or os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1
    rows = min(H, max_rows or H)
    gen = torch.Generator().manual_seed(0)
    cam = torch.rand(rows, W, generator=gen)
    proj = torch.rand(rows, W, generator=gen)
    g = torch.randn(rows, W, W, generator=torch.Generator().manual_seed(1))
    cells = rows * W * W
    zo.cpu_baseline_step_full(cam[:8], proj[:8], g[:8], k)  # warm
    best, n, t_total = None, 0, 0.0
    while n < 1 or (t_total < budget_s and n < 6):
        t0 = time.perf_counter()
        zo.cpu_baseline_step_full(cam, proj, g, k)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        t_total += dt
        n += 1
    return cells / best / 1e6, (f"1 pair, {rows} of {H} image rows, reference-shaped [rows,{W},{W}] volume "
                                f"(verify.py's bmm form, D ignored as in the reference), k={k}, fp32 torch CPU, "
                                f"fwd+WTA+bwd, best of {n}"), torch.get_num_threads()


def run_reference_cpu(args):
    """--impl reference: the reference's math on the host cores (the reference has no CPU kernel and its CUDA
    extension is not a CPU implementation, so this is the oracle port: kind = "port")."""
    rank, local, world = dist_env()
    if rank != 0:
        return
    H, W, D, k, _, desc = WORKLOADS[args.workload]
    import torch
    from oracle import zncc_oracle as zo
    torch.set_num_threads(os.cpu_count() or 1)               # all host threads; torchrun exports OMP_NUM_THREADS=1
    rows = min(H, args.ref_rows)
    gen = torch.Generator().manual_seed(0)
    cam = torch.rand(rows, W, generator=gen)
    proj = torch.rand(rows, W, generator=gen)
    g = torch.randn(rows, W, W, generator=torch.Generator().manual_seed(1))
    cells = rows * W * W
    for _ in range(args.warmup):
        zo.cpu_baseline_step_full(cam, proj, g, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        zo.cpu_baseline_step_full(cam, proj, g, k)
    dt = time.perf_counter() - t0
    value = cells * args.steps / dt / 1e6
    sample = (f"each step = 1 pair, {rows} of {H} image rows, reference-shaped [rows,{W},{W}] volume (verify.py's bmm "
              f"form; the reference ignores D), k={k}, fwd+WTA+bwd, fp32 torch on CPU")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "H": H, "W": W, "D": D, "kernel_size": k, "sample_rows": rows},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample, "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU kernel (CHECK_CUDA, custma/include/stereo_matching.hpp:20); this arm is "
                "the pure-PyTorch CPU restatement of examples/verify.py:91-123 that BASELINE.json's north_star names",
    }
    print(json.dumps(line), flush=True)


def run_reference_cuda(args):
    """--impl reference-cuda: the UNMODIFIED reference CUDA extension from baseline/_ref on this GPU (full [H,W,W]
    volume, since the reference ignores D).  Extra information, not part of the driver contract."""
    rank, local, world = dist_env()
    if rank != 0:
        return
    H, W, D, k, _, desc = WORKLOADS[args.workload]
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import torch
        import custma as ref_custma
        assert "baseline" in os.path.abspath(ref_custma.__file__)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference-cuda", "unavailable": f"{type(e).__name__}: {e}"[:200]}), flush=True)
        return
    torch.cuda.set_device(0)
    gen = torch.Generator().manual_seed(0)
    cam = torch.rand(H, W, generator=gen).cuda().requires_grad_(True)
    proj = torch.rand(H, W, generator=gen).cuda()
    cells = H * W * W
    if cells >= 2 ** 31:
        print(json.dumps({"impl": "reference-cuda", "unavailable": "H*W*W overflows the reference's int32 element count"}))
        return
    g = torch.randn(H, W, W, device="cuda")

    def step():
        cam.grad = None
        cv = ref_custma.stereo_matching(cam, proj, D, k)
        cv.backward(g)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "reference-cuda", "metric": METRIC, "value": cells / ms / 1e3, "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "f32", "data": "synthetic",
                      "config": {"workload": desc + " - reference computes the full [H,W,W] volume (D ignored)",
                                 "cells_per_step": cells}}), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class ResultGather:
    """Brings the per-rank [.., H, W] results to rank 0 (BASELINE.json: "NCCL ... used only to gather the results").

    Preferred: torch symmetric memory - rank 0 owns the gathered buffer, every rank copies its slice into rank 0's memory
    with one device-to-device copy (cudaMemcpyAsync over NVLink, copy engines: no SM is taken from the kernels).
    Fallback: NCCL dist.gather to rank 0.  Round 1 all-gathered to every rank (358 MB received per rank and step)."""

    def __init__(self, shape, dtype, device, rank, world, mode):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.mode = rank, world, "none"
        self.dist, self.torch = dist, torch
        self.shape = tuple(shape)
        if world == 1:
            return
        if mode in ("auto", "symm"):
            try:
                import torch.distributed._symmetric_memory as symm
                self.buf = symm.empty((world,) + self.shape, dtype=dtype, device=device)
                self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
                self.dst = self.hdl.get_buffer(0, (world,) + self.shape, dtype)[rank]
                self.mode = "symmetric-memory peer copy (copy engines)"
                self.copy_stream = torch.cuda.Stream(device)
                return
            except Exception as e:  # noqa: BLE001
                if mode == "symm":
                    raise
                self.why = f"{type(e).__name__}: {e}"[:120]
        if mode == "allgather":
            self.buf = torch.empty((world,) + self.shape, dtype=dtype, device=device)
            self.mode = "nccl all_gather_into_tensor"
            return
        self.buf = torch.empty((world,) + self.shape, dtype=dtype, device=device) if rank == 0 else None
        self.mode = "nccl gather to rank 0"

    def start(self, local, stream):
        """Asynchronous: returns a handle for wait()."""
        if self.world == 1:
            return None
        if self.mode.startswith("symmetric"):
            ev = self.torch.cuda.Event()
            ev.record(stream)
            with self.torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ev)
                self.dst.copy_(local, non_blocking=True)
                done = self.torch.cuda.Event()
                done.record(self.copy_stream)
            return done
        if self.mode.endswith("all_gather_into_tensor"):
            return self.dist.all_gather_into_tensor(self.buf, local, async_op=True)
        parts = [self.buf[r] for r in range(self.world)] if self.rank == 0 else None
        return self.dist.gather(local, parts, dst=0, async_op=True)

    def wait(self, handle, stream):
        if handle is None:
            return
        if self.mode.startswith("symmetric"):
            stream.wait_event(handle)
        else:
            handle.wait()

    def gathered(self):
        return self.buf


def copy_only_ceiling(P, pix, device, steps, dist, world):
    """The host-copy ceiling of this box for the e2e pattern: the same pinned-buffer traffic per step (2 images in, 3
    results out, per pair) on two streams, no kernels, every rank at once.  Returns ms per step (max over ranks)."""
    import torch
    n_in, n_out = 2 * P * pix, 3 * P * pix
    h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    h_out = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(2)]
    d_in = torch.empty(n_in, dtype=torch.float32, device=device)
    d_out = torch.empty(n_out, dtype=torch.float32, device=device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(n):
        for i in range(n):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[i & 1].copy_(d_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    run(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps * 1e3


def run_b200(args):
    import torch
    import torch.distributed as dist
    from custereomatching_b200 import binding

    rank, local, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    binding.load()                                                   # fail loudly if the extension is missing
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL announces its version on stdout when
    # NCCL_DEBUG is set in the environment) goes to stderr until the result is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_MAX_CTAS", "8")     # whatever NCCL still moves must leave the SMs to the kernels
        dist.init_process_group("nccl", device_id=device)
    try:
        line = bench_cfg5(args, rank, local, world, device) if args.workload == "cfg5" else \
            bench_batch(args, rank, local, world, device)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_batch(args, rank, local, world, device):
    import torch
    import torch.distributed as dist
    from custereomatching_b200 import binding

    H, W, D, k, default_pairs, desc = WORKLOADS[args.workload]
    strong = args.scaling == "strong"
    if strong:
        total = args.global_pairs or 64               # BASELINE configs[3]: the batch of 64, divided over the ranks
        if total % world:
            raise SystemExit(f"--global-pairs {total} is not a multiple of {world} ranks")
        P = total // world
    else:
        P = args.pairs_per_gpu or default_pairs
    flags = binding.FLAG_DIRECT if args.direct else 0
    C = D if D > 0 else W
    pix = H * W
    cells_rank = P * pix * C
    cells_job = cells_rank * world

    # ---- synthetic inputs: host (pinned) images, device-resident upstream gradient -------------------------------
    h_cam, h_proj = make_pairs(args.data, P, H, W, D, 1000 + rank)
    h_cam, h_proj = h_cam.pin_memory(), h_proj.pin_memory()
    cam = h_cam.to(device)
    proj = h_proj.to(device)
    ggen = torch.Generator(device=device).manual_seed(1)
    grad_in = torch.randn(P, H, W, C, device=device, generator=ggen)
    cost = torch.empty(P, H, W, C, device=device)
    # best, disparity and (one step later) the camera gradient travel to rank 0 in one buffer per kind
    results = torch.empty(2, P, H, W, device=device)
    best = results[0]
    disp = results[1].view(torch.int32)
    cam_grad2 = [torch.empty(P, H, W, device=device) for _ in range(2)]   # two buffers: the transfer of one step's gradient
    cam_grad = cam_grad2[0]                                               # runs under the next step's forward
    ws_bytes = max(binding.forward_workspace_bytes(P, H, W, D, k, flags),
                   binding.backward_workspace_bytes(P, H, W, D, k, flags))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    # the backward's image-dependent preparation runs on a second stream beside the forward, into its own workspace
    # (custma_backward_prepare / CUSTMA_FLAG_PREPARED; --no-prepare: the plain two calls)
    prepare = not args.no_prepare and P * H * W * C >= 32e6    # below that the second stream's events cost more than they hide
    ws_bwd_bytes = binding.backward_workspace_bytes(P, H, W, D, k, flags)
    ws_bwd = torch.empty(max(ws_bwd_bytes, 256), dtype=torch.uint8, device=device) if prepare else None
    prep_stream = torch.cuda.Stream(device) if prepare else None
    gather_res = ResultGather((2, P, H, W), torch.float32, device, rank, world, args.gather)
    gather_grad = ResultGather((P, H, W), torch.float32, device, rank, world, args.gather)
    # the step runs on a high-priority stream, so that the backward's preparation (prep_stream, default priority) only
    # takes what the forward leaves idle
    torch.cuda.synchronize(device)
    stream = torch.cuda.Stream(device, priority=-1) if prepare else torch.cuda.current_stream(device)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def fwd():
        binding.forward(cam.data_ptr(), proj.data_ptr(), cost.data_ptr(), best.data_ptr(), disp.data_ptr(),
                        P, H, W, D, k, flags, ws.data_ptr(), ws_bytes, sptr)

    def bwd(out=None):
        binding.backward(grad_in.data_ptr(), cam.data_ptr(), proj.data_ptr(), (cam_grad if out is None else out).data_ptr(),
                         P, H, W, D, k, flags, ws.data_ptr(), ws_bytes, sptr)

    def prepare_bwd(main, side):
        side.wait_stream(main)
        binding.backward_prepare(cam.data_ptr(), proj.data_ptr(), P, H, W, D, k, flags, ws_bwd.data_ptr(), ws_bwd_bytes,
                                 side.cuda_stream)

    def bwd_prepared(out, main, side):
        main.wait_stream(side)
        binding.backward(grad_in.data_ptr(), cam.data_ptr(), proj.data_ptr(), out.data_ptr(), P, H, W, D, k,
                         flags | binding.FLAG_PREPARED, ws_bwd.data_ptr(), ws_bwd_bytes, main.cuda_stream)

    state = {"i": 0, "grad_pending": None}

    def step_eager():
        # results only cross GPUs: 3 * P*H*W*4 bytes per rank; the volume never leaves its GPU.  Both transfers are
        # asynchronous: best / disparity leave under this step's backward, the camera gradient under the NEXT step's
        # forward; finish_steps() waits for the last one inside the timed region.
        if prepare:
            prepare_bwd(stream, prep_stream)
        fwd()
        pending = gather_res.start(results, stream)
        gather_grad.wait(state["grad_pending"], stream)
        out = cam_grad2[state["i"] & 1]
        state["i"] += 1
        if prepare:
            bwd_prepared(out, stream, prep_stream)
        else:
            bwd(out)
        state["grad_pending"] = gather_grad.start(out, stream)
        gather_res.wait(pending, stream)

    def finish_steps():
        gather_grad.wait(state["grad_pending"], stream)
        state["grad_pending"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        step_eager()
    finish_steps()
    barrier()

    # optional: the whole step as one CUDA graph (single GPU; everything the C ABI enqueues is capturable - kernels and
    # one memset, no allocation, no synchronisation)
    graph = None
    step = step_eager
    if args.graph and world == 1:
        side = torch.cuda.Stream(device)
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            sp = side.cuda_stream
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                if prepare:
                    prepare_bwd(side, prep_stream)
                binding.forward(cam.data_ptr(), proj.data_ptr(), cost.data_ptr(), best.data_ptr(), disp.data_ptr(),
                                P, H, W, D, k, flags, ws.data_ptr(), ws_bytes, sp)
                if prepare:
                    bwd_prepared(cam_grad, side, prep_stream)
                else:
                    binding.backward(grad_in.data_ptr(), cam.data_ptr(), proj.data_ptr(), cam_grad.data_ptr(),
                                     P, H, W, D, k, flags, ws.data_ptr(), ws_bytes, sp)
        stream.wait_stream(side)
        step = graph.replay
        for _ in range(3):
            step()
        barrier()

    # ---- timed region 1: whole step, device-resident inputs ------------------------------------------------------
    K = args.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    launches0 = binding.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev[0].record(stream)
        for _ in range(K):
            step()
        finish_steps()
        ev[1].record(stream)
        barrier()
    launches = (binding.launch_count() - launches0) // K if graph is None else None
    ms_total = ev[0].elapsed_time(ev[1])
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K

    # every rank's results must have arrived on rank 0 unchanged: per-rank checksums, exchanged out of band
    gather_ok = None
    if world > 1:
        mine = torch.stack([results.double().sum(), cam_grad2[(state["i"] - 1) & 1].double().sum()])
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            gr, gg = gather_res.gathered(), gather_grad.gathered()
            gather_ok = all(torch.equal(torch.stack([gr[r].double().sum(), gg[r].double().sum()]), sums[r]) for r in range(world))
            gather_ok = bool(gather_ok and torch.equal(gr[0], results))
            if not gather_ok:
                raise SystemExit("bench: the gathered results on rank 0 differ from what the ranks computed")

    # ---- timed region 2: forward and backward kernel groups separately (roofline of the dominant one) ------------
    Kk = min(K, 20)
    e_f = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kk)]
    e_b = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kk)]
    barrier()
    for i in range(Kk):
        e_f[i][0].record(stream); fwd(); e_f[i][1].record(stream)
        e_b[i][0].record(stream); bwd(); e_b[i][1].record(stream)
    barrier()
    ms_f = sum(a.elapsed_time(b) for a, b in e_f) / Kk
    ms_b = sum(a.elapsed_time(b) for a, b in e_b) / Kk
    peak, peak_src = measured_peak_gbs()
    bytes_f = 4 * cells_rank + 8 * P * pix + 8 * P * pix      # write cost once; read 2 images; write best + disparity
    bytes_b = 4 * cells_rank + 8 * P * pix + 4 * P * pix      # read upstream gradient once; read 2 images; write grad
    roof_f = {"kernel": "custma_forward (window stats + cost volume + WTA)", "bound": "hbm",
              "achieved": bytes_f / ms_f / 1e6, "peak": peak, "unit": "GB/s", "frac": bytes_f / ms_f / 1e6 / peak,
              "traffic": None, "ms_per_launch": ms_f, "algorithmic_bytes": bytes_f}
    roof_b = {"kernel": "custma_backward (window stats + camera gradient)", "bound": "hbm",
              "achieved": bytes_b / ms_b / 1e6, "peak": peak, "unit": "GB/s", "frac": bytes_b / ms_b / 1e6 / peak,
              "traffic": None, "ms_per_launch": ms_b, "algorithmic_bytes": bytes_b}
    dominant = dict(roof_b if ms_b >= ms_f else roof_f)
    dominant["peak_source"] = peak_src
    dominant["step"] = {"fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_frac": roof_f["frac"], "bwd_frac": roof_b["frac"],
                        "fwd_bwd_frac": (bytes_f + bytes_b) / (ms_f + ms_b) / 1e6 / peak}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the last ncu capture
    if os.path.exists(traffic_path):
        try:
            tr = json.load(open(traffic_path))
            key = "backward" if ms_b >= ms_f else "forward"
            if tr.get("workload") == args.workload and tr.get("pairs_per_gpu") == P and args.data == tr.get("data", "rand"):
                dominant["traffic"] = tr.get(key)
        except Exception:
            pass
    # how much of the call the conditioning verdict flagged (work items of the per-cell fallback; above 4 % of the
    # capacity the tensor-core kernels take the whole call)
    verdict = None
    try:
        info = binding.debug_layout_info(P, H, W, D, k)
        if info is not None:
            fwd()
            torch.cuda.synchronize(device)
            cnt = int(ws[info["fb_count_offset"]:info["fb_count_offset"] + 4].view(torch.int32).item())
            verdict = {"flagged_items": cnt, "capacity": info["fb_capacity"],
                       "flagged_share": cnt / max(info["fb_capacity"], 1),
                       "path": "tensor-core kernels" if cnt > 0.04 * info["fb_capacity"] and info["tc_supported"] else
                               "sliding-window kernels (+ per-cell fallback for flagged tiles)"}
    except Exception:
        verdict = None

    # ---- timed region 3: end to end through the C-ABI host entry point -------------------------------------------
    e2e = None
    if not args.no_e2e:
        # two sets of host result buffers: step i is submitted, then the results of step i-1 are waited for - the
        # streaming use of the host entry point (copies of one step under the kernels of the other).  The volume of a
        # ticket lives in the library's own per-slot buffer (cost_volume_dev = NULL): two tickets in flight never share one.
        h_best2 = [torch.empty(P, H, W).pin_memory() for _ in range(2)]
        h_disp2 = [torch.empty(P, H, W, dtype=torch.int32).pin_memory() for _ in range(2)]
        h_grad2 = [torch.empty(P, H, W).pin_memory() for _ in range(2)]
        del cost
        torch.cuda.empty_cache()

        def host_submit(i):
            return binding.host_submit(h_cam.data_ptr(), h_proj.data_ptr(), h_best2[i & 1].data_ptr(), h_disp2[i & 1].data_ptr(),
                                       h_grad2[i & 1].data_ptr(), 0, grad_in.data_ptr(), P, H, W, D, k, flags)

        for i in range(3):
            binding.host_wait(host_submit(i))
        Ke = max(30, K)   # enough steps that filling and draining the pipeline (one copy in, one copy out) does not show
        barrier()
        t0 = time.perf_counter()
        prev = None
        for i in range(Ke):
            ticket = host_submit(i)                                  # H2D + forward + backward + D2H of step i enqueued
            if prev is not None:
                binding.host_wait(prev)                              # results of step i-1 are on the host
            prev = ticket
        binding.host_wait(prev)
        barrier()
        dt = time.perf_counter() - t0
        h_best, h_grad = h_best2[(Ke - 1) & 1], h_grad2[(Ke - 1) & 1]
        te = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": cells_job * Ke / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * P * pix * 4,
               "d2h_bytes_per_step": 3 * P * pix * 4, "ms_per_step": dt / Ke * 1e3, "steps": Ke,
               "api": "custma_host_submit / custma_host_wait (include/custma_b200.h; custma_host_step = both): pinned host "
                      "images in, host best/disparity/camera_grad out every step, results of step i-1 awaited after "
                      "step i is submitted; upstream gradient produced on the device; volume in the library's per-slot buffer; "
                      "copies in / kernels / copies out on three streams, the image-dependent preparations of a step "
                      "beside the previous step's kernels"}
        # parity of the two paths on this very data (cheap sanity, outside the timed regions)
        same = bool(torch.equal(h_best.to(device), best) and torch.equal(h_grad.to(device), cam_grad))
        e2e["matches_device_path"] = same
        binding.host_release()
        # the same host traffic with no kernels at all: what the box's pinned-memory / PCIe path allows per step
        ceil_ms = copy_only_ceiling(P, pix, device, Ke, dist, world)
        e2e["copy_only_ms_per_step"] = ceil_ms
        e2e["copy_only_GBps_per_rank"] = 5 * P * pix * 4 / ceil_ms / 1e6
        e2e["frac_of_copy_ceiling"] = min(1.0, ceil_ms / (dt / Ke * 1e3)) if ceil_ms >= ms_step else None
        e2e["bound"] = ("host copies (pinned memory / PCIe): the copy-only loop alone takes longer than the device-resident step"
                        if ceil_ms >= ms_step else "kernels: the copies fit under the device-resident step")

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sample, threads = cpu_torch_port_rate(H, W, D, k, max_rows=args.ref_rows)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "host_cpus": os.cpu_count()}

    return {
        "metric": METRIC, "value": cells_job / ms_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": f"synthetic ({args.data}: {DATA_FAMILIES[args.data]})",
        "config": {"workload": desc, "H": H, "W": W, "D": D, "kernel_size": k, "pairs_per_gpu": P,
                   "global_pairs": P * world, "cells_per_step": cells_job,
                   "parallelism": f"batch-sharded x{world}" if world > 1 else "single GPU",
                   "result_gather": gather_res.mode, "gathered_equals_local": gather_ok,
                   "cuda_graph": graph is not None,
                   "backward_prepared_beside_forward": prepare,
                   "l2": f"inputs larger than L2: {4 * cells_rank / 1e6:.0f} MB volume written and "
                         f"{4 * cells_rank / 1e6:.0f} MB gradient read per step per GPU (L2 = 126 MB)",
                   "kernels": "direct two-pass" if args.direct else "default (sliding-window where available)",
                   "verdict": verdict},
        "roofline": dominant, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches) if launches is not None else "one CUDA graph replay per step",
        "clocks": clocks.summary(),
    }


def bench_cfg5(args, rank, local, world, device):
    """BASELINE configs[4]: ONE 7680x4320 pair, 512 disparities, row-band sharded with a window-radius halo.  Every rank
    holds both images, owns volume rows [h0,h1), computes them from its haloed crop (forward + WTA), and runs the backward
    for the upstream gradient of its own rows only (custma_backward_rows).  Results go to rank 0: the [rows,W] best /
    disparity bands and the haloed camera-gradient band, whose halo rows rank 0 adds in a fixed order."""
    import torch
    import torch.distributed as dist
    from custereomatching_b200 import binding
    from custereomatching_b200 import sharding as sh

    H, W, D, k, _, desc = WORKLOADS["cfg5"]
    if args.rows:                                   # a shorter pair for quick runs
        H = args.rows
    band = sh.row_band(H, k, rank, world)
    Hc, rows = band.hi - band.lo, band.rows
    cam_full, proj_full = make_pairs(args.data, 1, H, W, D, 7)
    cam = cam_full[0, band.lo:band.hi].contiguous().to(device)
    proj = proj_full[0, band.lo:band.hi].contiguous().to(device)
    del cam_full, proj_full
    cells_job = H * W * D
    b0, b1 = sh.band_gradient_mask_rows(band)
    cost = torch.empty(Hc, W, D, device=device)
    gen = torch.Generator(device=device).manual_seed(11 + rank)
    grad_in = torch.randn(rows, W, D, device=device, generator=gen)      # gradient of the OWNED rows only
    max_rows = max(e - b for b, e in sh.split_even(H, world))
    max_hc = max_rows + k - 1
    results = torch.zeros(2, max_hc, W, device=device)                     # best, disparity on the crop's rows (padded)
    best, disp = results[0, :Hc], results[1, :Hc].view(torch.int32)
    cam_grad = torch.zeros(max_hc, W, device=device)
    ws_bytes = max(binding.forward_workspace_bytes(1, Hc, W, D, k, 0), binding.backward_workspace_bytes(1, Hc, W, D, k, 0))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    gather_res = ResultGather((2, max_hc, W), torch.float32, device, rank, world, args.gather)
    gather_grad = ResultGather((max_hc, W), torch.float32, device, rank, world, args.gather)
    # as in bench_batch: the step on a high-priority stream, the backward's preparation beside the forward
    prepare = not args.no_prepare
    torch.cuda.synchronize(device)
    stream = torch.cuda.Stream(device, priority=-1) if prepare else torch.cuda.current_stream(device)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    ws_bwd_bytes = binding.backward_workspace_bytes(1, Hc, W, D, k, 0)
    ws_bwd = torch.empty(ws_bwd_bytes, dtype=torch.uint8, device=device) if prepare else None
    prep_stream = torch.cuda.Stream(device) if prepare else None

    def fwd():
        binding.forward(cam.data_ptr(), proj.data_ptr(), cost.data_ptr(), best.data_ptr(), disp.data_ptr(), 1, Hc, W, D, k, 0,
                        ws.data_ptr(), ws_bytes, sptr)

    def bwd(prepared=False):
        binding.backward_rows(grad_in.data_ptr(), cam.data_ptr(), proj.data_ptr(), cam_grad.data_ptr(), 1, Hc, W, D, k,
                              b0, b1, binding.FLAG_PREPARED if prepared else 0,
                              (ws_bwd if prepared else ws).data_ptr(), ws_bwd_bytes if prepared else ws_bytes, sptr)

    def step():
        if prepare:
            prep_stream.wait_stream(stream)
            binding.backward_prepare(cam.data_ptr(), proj.data_ptr(), 1, Hc, W, D, k, 0, ws_bwd.data_ptr(), ws_bwd_bytes,
                                     prep_stream.cuda_stream)
        fwd()
        p1 = gather_res.start(results, stream)
        if prepare:
            stream.wait_stream(prep_stream)
        bwd(prepare)
        p2 = gather_grad.start(cam_grad, stream)
        gather_res.wait(p1, stream)
        gather_grad.wait(p2, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    K = args.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    launches0 = binding.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev[0].record(stream)
        for _ in range(K):
            step()
        ev[1].record(stream)
        barrier()
    launches = (binding.launch_count() - launches0) // K
    t = torch.tensor([ev[0].elapsed_time(ev[1])], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    # kernels alone (no gather), and the assembly on rank 0
    barrier()
    ev[0].record(stream)
    for _ in range(K):
        fwd()
    ev[1].record(stream)
    for _ in range(K):
        bwd()
    ev[2].record(stream)
    barrier()
    ms_f, ms_b = ev[0].elapsed_time(ev[1]) / K, ev[1].elapsed_time(ev[2]) / K
    tk = torch.tensor([ms_f, ms_b], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    ms_f, ms_b = float(tk[0]), float(tk[1])
    assembled_ok = None
    if rank == 0:
        bands = [sh.row_band(H, k, r, world) for r in range(world)]
        gr = gather_res.gathered() if world > 1 else results[None]
        gg = gather_grad.gathered() if world > 1 else cam_grad[None]
        best_full = torch.cat([gr[r, 0, b.top_halo:b.top_halo + b.rows] for r, b in enumerate(bands)])
        grad_full = torch.zeros(H, W, device=device)
        for r, b in enumerate(bands):
            grad_full[b.lo:b.hi] += gg[r, :b.hi - b.lo]
        # own rows must equal what this rank computed; everything finite
        assembled_ok = bool(torch.equal(best_full[band.h0:band.h1], best[b0:b1]) and torch.isfinite(grad_full).all()
                            and best_full.shape[0] == H)
        if not assembled_ok:
            raise SystemExit("bench cfg5: assembled results are inconsistent")
    peak, peak_src = measured_peak_gbs()
    rank_cells = max_rows * W * D
    bytes_f, bytes_b = 4 * rank_cells, 4 * rank_cells
    dominant = {"kernel": "custma_backward_rows (camera gradient of one row band)", "bound": "hbm",
                "achieved": bytes_b / ms_b / 1e6, "peak": peak, "unit": "GB/s", "frac": bytes_b / ms_b / 1e6 / peak,
                "traffic": None, "ms_per_launch": ms_b, "algorithmic_bytes": bytes_b, "peak_source": peak_src,
                "step": {"fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_frac": bytes_f / ms_f / 1e6 / peak,
                         "bwd_frac": bytes_b / ms_b / 1e6 / peak,
                         "fwd_bwd_frac": (bytes_f + bytes_b) / (ms_f + ms_b) / 1e6 / peak,
                         "gather_and_overlap_ms": ms_step - ms_f - ms_b}}
    return {
        "metric": METRIC, "value": cells_job / ms_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": f"synthetic ({args.data}: {DATA_FAMILIES[args.data]})",
        "config": {"workload": desc, "H": H, "W": W, "D": D, "kernel_size": k, "cells_per_step": cells_job,
                   "parallelism": f"row-band sharded x{world}, halo {k // 2} rows above / {k - 1 - k // 2} below",
                   "rows_per_rank": max_rows, "volume_GB_per_rank": 4 * rank_cells / 1e9,
                   "result_gather": gather_res.mode, "assembled_on_rank0": assembled_ok,
                   "l2": f"{4 * rank_cells / 1e6:.0f} MB volume written and as much gradient read per step per GPU (L2 = 126 MB)"},
        "roofline": dominant, "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks.summary(),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--data", default="rand", choices=sorted(DATA_FAMILIES), help="synthetic image family")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed pairs per GPU; strong: --global-pairs (default 64 = BASELINE configs[3]) divided over the ranks")
    ap.add_argument("--global-pairs", type=int, default=0)
    ap.add_argument("--gather", default="auto", choices=["auto", "symm", "gather", "allgather"],
                    help="how results reach rank 0: symmetric-memory peer copies, NCCL gather, or round 1's all_gather")
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (single GPU)")
    ap.add_argument("--rows", type=int, default=0, help="cfg5 only: a shorter image (quick runs)")
    ap.add_argument("--direct", action="store_true", help="force the direct two-pass kernels (CUSTMA_FLAG_DIRECT)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prepare", action="store_true",
                    help="plain custma_forward + custma_backward (default: custma_backward_prepare beside the forward)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-rows", type=int, default=48, help="image rows of one pair in the CPU sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
