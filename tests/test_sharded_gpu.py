"""The N > 1 CUDA + NCCL path against the oracle on hardware (needs >= 2 GPUs; the single-GPU emulation of
the same exchange is tests/test_gpu_parity.py::test_packed_event_exchange_two_shards)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_step_and_recovery_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(REPO, "tests", "sharded_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert p.returncode == 0 and "SHARDED_OK" in p.stdout, (p.stdout[-2000:], p.stderr[-4000:])
