"""Multi-GPU (needs >= 2 visible B200s; skipped otherwise): one process per GPU under torchrun,
NCCL communicator attached through the C ABI, one solve sharded over the ranks with a single
all-reduce of the per-unit partial sums, and independent solves sharded with no collective
(profiles/multi_gpu_check.py asserts bit-identical results across ranks and <=1e-12 vs 1 GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_nccl_sharded_solve():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "profiles", "multi_gpu_check.py"), "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "[multi] world=2" in r.stdout and "[multi] weak:" in r.stdout
