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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_can_drive_two_devices():
    """The kernels that opt in to more than 48 KiB of shared memory (play tables, observation rings, pipelined GAE,
    embedding tables) configure themselves once PER DEVICE: the same process gets the same results on cuda:1 after
    having used cuda:0."""
    from g2048 import engine as E

    results = []
    for d in (0, 1):
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            subs = E.chain_advance(E.words_tensor([0, 99], dev), E.RNG_PARTITIONABLE, 1 + 2 * 1024)
            out = E.play(E.POLICY_RANDOM, subs, 40000, 0, 40000, E.RNG_PARTITIONABLE, entry="g2048_play_tables")
            g = torch.Generator(device=dev).manual_seed(5)
            n = (1 << 23) + 77  # the pipelined flat-GAE kernel
            r, v = torch.rand(n, device=dev, generator=g), torch.rand(n, device=dev, generator=g)
            dn = (torch.rand(n, device=dev, generator=g) < 0.01).to(torch.uint8)
            adv, ret, mom = E.gae_flat(r, v, dn, 0.99, 0.95)
            boards = out["final_boards"][:5000].contiguous()
            obs = torch.empty((5000, 16, 31), dtype=torch.float32, device=dev)
            E.N.call("g2048_expand_obs", E.N.ptr(boards), 5000, E.N.OBS_F32, E.N.ptr(obs), 0, 0, E.N.stream_ptr())
            table = torch.randn(31, 1024, device=dev, generator=g)  # 124 KiB of table in shared memory
            emb = E.embed_boards(boards, table)
            torch.cuda.synchronize()
            results.append([t.cpu() for t in (out["final_boards"], out["lengths"], adv, ret, obs, emb)])
    for a, b in zip(*results):
        assert torch.equal(a, b)
