"""The N > 1 CUDA + NCCL path against the oracle on hardware (needs >= 2 GPUs; the single-GPU emulation of
the same exchange is tests/test_gpu_parity.py::test_packed_event_exchange_two_shards)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["auto", "nccl"])
def test_sharded_step_and_recovery_two_ranks(exchange):
    """exchange "auto": the fused peer-memory event exchange where torch can set symmetric memory up
    (the worker prints which one ran), "nccl": pack + all-gather + import."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631" if exchange == "auto" else "29632", os.path.join(REPO, "tests", "sharded_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO, env=dict(os.environ, B200ADSB_EXCHANGE=exchange))
    assert p.returncode == 0 and "SHARDED_OK" in p.stdout, (p.stdout[-2000:], p.stderr[-4000:])
    print(p.stdout[-200:])
