"""Multi-GPU parity over NCCL: needs at least two GPUs in the box (skipped otherwise; the gloo tests in
test_sharding_cpu.py cover the host logic everywhere).  Launches tests/nccl_worker.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_steps_over_nccl_match_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_PARITY_OK" in r.stdout, r.stdout[-4000:]
