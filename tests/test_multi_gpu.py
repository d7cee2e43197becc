"""The sharded path over NCCL on real GPUs (skipped on boxes with fewer than two): tools/multi_gpu_check.py under
torchrun compares sharded runs with the same batches on one GPU.  The gloo world-size-2 tests of the same reductions
(tests/test_host_logic.py) run everywhere."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_runs_equal_single_gpu_runs_over_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "multi-GPU check: all passed" in res.stdout
    assert "FAIL" not in res.stdout
