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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_a_runner_on_cuda1_works_while_cuda0_is_current():
    """`device=` is where the runner's kernels RUN, not only where its tensors live: every libg2048 call goes to the
    device of the tensors it is given (and to that device's current stream) whatever the process's current device is
    (g2048/_native.py: ptr / stream_ptr / call)."""
    import g2048
    from g2048.ppo import collect_rollouts

    torch.cuda.set_device(0)
    want = g2048.BatchRunner(init_seed=5, act_fn=g2048.act_randomly, device="cuda:0")
    got = g2048.BatchRunner(init_seed=5, act_fn=g2048.act_randomly, device="cuda:1")
    assert torch.cuda.current_device() == 0
    for batch in (300, 40000):
        a, b = want.run_flat_batch(batch), got.run_flat_batch(batch)
        assert b.boards.device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
        for name in ("boards", "meta", "rewards", "log_probs", "lengths", "final_boards", "scores"):
            assert torch.equal(getattr(a, name).cpu(), getattr(b, name).cpu()), name
        pa, pb = want.run_packed_batch(200), got.run_packed_batch(200)
        assert torch.equal(pa.boards.cpu(), pb.boards.cpu()) and torch.equal(pa.rewards.cpu(), pb.rewards.cpu())
    assert (want.key == got.key).all()
    buf = g2048.RolloutBuffer(31, 16, 4)
    out = collect_rollouts(got, buf, 500, 1)
    packed = buf.get_packed()
    assert packed["boards"].device == torch.device("cuda", 1) and out["total_episodes"] == 500
    batches = g2048.DevicePPOBatches(packed, batch_size=256, epoch_prefetch=True)
    first = next(iter(batches))
    assert first["observations"].device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
    # the same minibatch from the same buffer gathered on cuda:0
    packed0 = {k: v.to("cuda:0") for k, v in packed.items()}
    torch.manual_seed(3)
    ref = next(iter(g2048.DevicePPOBatches(packed0, batch_size=256)))
    torch.manual_seed(3)
    mine = next(iter(g2048.DevicePPOBatches(packed, batch_size=256)))
    for k in ref:
        assert torch.equal(ref[k].cpu(), mine[k].cpu()), k
