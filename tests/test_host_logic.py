"""CPU-only checks: the C-ABI library loads and exports what include/g2048.h declares, the host
logic (validation, statistics merge, sharding, distributed reductions over gloo) is right, there
is no CPU fallback, and the per-env DEVICE functions -- compiled for the host through the
test-only shim in tests/host_shim/ -- agree with the oracle."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import pgx2048_oracle as O

ROOT = Path(__file__).resolve().parent.parent


# ----------------------------------------------------------------------------------------- boundary
def test_library_exports_every_declared_symbol():
    from g2048 import _native as N

    declared = N.declared_symbols()
    assert len(declared) >= 30
    nm = subprocess.run(["nm", "-D", "--defined-only", str(N.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(N._SIGNATURES), set(declared) ^ set(N._SIGNATURES)
    assert N.lib.g2048_version() >= 100


def test_library_is_sm100a_only():
    from g2048 import _native as N

    out = subprocess.run(["cuobjdump", "--list-elf", str(N.LIB_PATH)], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a GPU-less host")
def test_no_cpu_fallback():
    import g2048

    with pytest.raises(RuntimeError, match="CUDA"):
        g2048.BatchRunner(0, g2048.act_randomly)
    with pytest.raises(RuntimeError, match="CUDA"):
        g2048.act_drul(None, np.zeros((4, 4, 31), bool), np.ones(4, bool))
    with pytest.raises(RuntimeError, match="CUDA"):
        g2048.RunningStatsVec().push(np.zeros((2, 3)))
    with pytest.raises(RuntimeError, match="CUDA"):
        g2048.engine.gae_host(np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.uint8), 0.99, 0.95, True)
    # the C entry points themselves report a CUDA error code instead of computing on the host
    from g2048 import _native as N

    stats = np.zeros(N.PLAY_STATS_WORDS, np.uint64)
    rc = N.lib.g2048_play_host(0, 0, None, 4, 0, 4, 1, None, None, None, stats.ctypes.data)
    assert rc > 0 and "CUDA" in N.last_error()


def test_bad_arguments_are_rejected_before_any_launch():
    from g2048 import _native as N

    assert N.lib.g2048_env_init(None, 4, 3, 2, 1, None, None, None) == -1  # env_lo + n > batch
    assert "env_init" in N.last_error()
    assert N.lib.g2048_play(7, None, 100, 4, 0, 4, 1, None, None, None, None, None, None) == -1
    assert N.lib.g2048_normalize(None, 4, None, 2, None) == -1
    assert N.lib.g2048_gae_flat_scratch_bytes(1) == 24 and N.lib.g2048_gae_flat_scratch_bytes(1025) == 32


# ----------------------------------------------------------------------------------------- host logic
def test_rollout_buffer_validation_messages():
    from g2048 import RolloutBuffer

    rb = RolloutBuffer(31, (4, 4), 4)
    ok = rb._validate_and_reshape_observations(np.zeros((2, 3, 16, 31)))
    assert ok.shape == (2, 3, 4, 4, 31)
    assert rb._validate_and_reshape_observations(np.zeros((2, 3, 4, 4, 31))).shape == (2, 3, 4, 4, 31)
    with pytest.raises(ValueError, match="must have at least 2 dimensions"):
        rb._validate_and_reshape_observations(np.zeros(7))
    with pytest.raises(ValueError, match="Failed to reshape observations") as exc:
        rb._validate_and_reshape_observations(np.zeros((2, 3, 5, 31)))
    assert "Cannot reshape observations" in str(exc.value)
    rb = RolloutBuffer(31, 16, 4)
    assert rb._validate_and_reshape_observations(np.zeros((2, 3, 4, 4, 31))).shape == (2, 3, 16, 31)
    assert rb.buffer_size == 0


def test_running_stats_merge_matches_reference_fixture(golden_ppo):
    from g2048 import RunningStatsVec

    rs = RunningStatsVec()
    for i in range(4):
        x = golden_ppo[f"rs_push{i}"]
        rs.merge_triple(np.full((3, 1), x.shape[1], np.int64), x.mean(1, keepdims=True), x.var(1, keepdims=True))
        np.testing.assert_allclose(rs.mean, golden_ppo[f"rs_mean{i}"], rtol=1e-12)
        np.testing.assert_allclose(rs.variance, golden_ppo[f"rs_var{i}"], rtol=1e-12)
        np.testing.assert_array_equal(rs.num_samples, golden_ppo[f"rs_n{i}"])
    other = RunningStatsVec()
    other.merge(rs)
    np.testing.assert_allclose(other.mean, rs.mean)
    rs.clear()
    assert rs.mean == 0.0 and rs.std == 0.0


def test_shard_ranges_tile_the_batch():
    from g2048.dist import shard_range

    for n in (1, 7, 1024, 16_777_216):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_rng_mode_resolution(monkeypatch):
    from g2048 import engine as E

    monkeypatch.delenv("G2048_THREEFRY_PARTITIONABLE", raising=False)
    monkeypatch.delenv("JAX_THREEFRY_PARTITIONABLE", raising=False)
    assert E.resolve_rng_mode(None) == E.RNG_PARTITIONABLE  # jax 0.5.3 default
    assert E.resolve_rng_mode("original") == E.RNG_ORIGINAL and E.resolve_rng_mode(True) == E.RNG_PARTITIONABLE
    monkeypatch.setenv("G2048_THREEFRY_PARTITIONABLE", "0")
    assert E.resolve_rng_mode(None) == E.RNG_ORIGINAL
    # jax.random.key(seed) in jax's default 32-bit mode: the high word is always 0, negative seeds keep their
    # two's-complement bits, anything that does not fit 32 bits raises (golden: jax.random.key(-1) -> [0, 4294967295])
    assert E.key_words(42) == (0, 42) and E.key_words(-1) == (0, 0xFFFFFFFF) and E.key_words(0xFFFFFFFF) == (0, 0xFFFFFFFF)
    assert E.key_words(-(1 << 31)) == (0, 0x80000000)
    for bad in ((7 << 32) | 9, 1 << 32, -(1 << 31) - 1):
        with pytest.raises(OverflowError):
            E.key_words(bad)
    b = np.arange(16)[None, :] % 16
    np.testing.assert_array_equal(E.boards_numpy(torch.from_numpy(E.pack_boards(b))), b)


# ----------------------------------------------------------------------------------------- gloo, world_size 2
_WORKER = r"""
import os, sys
sys.path[:0] = [{root!r}, {pkg!r}]
import numpy as np, torch, torch.distributed as dist
from g2048 import dist as D
from g2048.stats import RunningStatsVec
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["MASTER_PORT"], rank=rank, world_size=world)
rng = np.random.default_rng(0)
data = rng.standard_normal((3, 1000)) * 5 + 2
lo, hi = D.shard_range(1000, rank, world)
# episode statistics: every rank folds its own shard, then all ranks agree on the global triple
rs = RunningStatsVec()
part = data[:, lo:hi]
rs.merge_triple(np.full((3, 1), hi - lo, np.int64), part.mean(1, keepdims=True), part.var(1, keepdims=True))
rs.all_reduce()
assert np.allclose(rs.mean[:, 0], data.mean(1)) and np.allclose(rs.variance[:, 0], data.var(1)), "stats merge"
assert (rs.num_samples == 1000).all()
# play statistics block: sums, except slot 5 which is a max
stats = torch.arange(32, dtype=torch.int64) * (rank + 1)
stats[5] = 100 + rank
out = D.allreduce_play_stats(stats)
want = torch.arange(32, dtype=torch.int64) * sum(r + 1 for r in range(world))
want[5] = 100 + world - 1
assert torch.equal(out, want), "play stats"
# normalisation moments
m = torch.tensor([hi - lo, part[0].sum(), (part[0] ** 2).sum(), 0, 0, 0], dtype=torch.float64)
D.allreduce_sum_(m)
assert m[0] == 1000 and abs(m[1] - data[0].sum()) < 1e-9 and abs(m[2] - (data[0] ** 2).sum()) < 1e-6, "moments"
assert D.allreduce_max_int(10 * (rank + 1), torch.device("cpu")) == 10 * world
assert D.all_ranks_true(True, torch.device("cpu")) and not D.all_ranks_true(rank == 0, torch.device("cpu"))
# gradient averaging
p = torch.nn.Parameter(torch.zeros(5)); p.grad = torch.full((5,), float(rank + 1))
q = torch.nn.Parameter(torch.zeros(2, 3)); q.grad = torch.full((2, 3), float(10 * (rank + 1)))
D.allreduce_gradients([p, q], bucket_bytes=16)
mean = sum(r + 1 for r in range(world)) / world
assert torch.allclose(p.grad, torch.full((5,), mean)) and torch.allclose(q.grad, torch.full((2, 3), 10 * mean)), "grads"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_distributed_reductions_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT), pkg=str(ROOT / "2048-ppo-agent_b200")))
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {rank} ok" in out, out[-2000:]


# ----------------------------------------------------------------------------------------- device code on the host
@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "libg2048_hostshim.so"
    src = ROOT / "tests" / "host_shim" / "harness.cpp"
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-w", "-o", str(out), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pack(b):
    b = np.asarray(b, np.uint64).reshape(-1, 16)
    return (b << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(1).astype(np.uint64)


def _unpack(x):
    return ((x[:, None] >> (np.arange(16, dtype=np.uint64) * np.uint64(4))) & np.uint64(15)).astype(np.uint8)


def test_kth_flag_cell_exhaustively(shim):
    """The multiply-based prefix count behind the spawn position, for all 65 536 empty-cell masks and every k."""
    got = np.zeros((65536, 16), np.int32)
    shim.shim_kth_flag_cell_all(_p(got))
    masks = np.arange(65536, dtype=np.uint32)
    bits = ((masks[:, None] >> np.arange(16)) & 1).astype(np.int32)  # (65536, 16)
    prefix = bits.cumsum(axis=1)
    want = np.full((65536, 16), -1, np.int32)
    for k in range(1, 17):
        hit = (prefix == k) & (bits == 1)  # the k-th set flag
        has = hit.any(axis=1)
        want[has, k - 1] = hit[has].argmax(axis=1)
    np.testing.assert_array_equal(got, want)


def test_device_random_subset_on_host_matches_oracle(shim):
    """feistel_position (g2048_rng.cuh, what random_subset_kernel evaluates per index) against the oracle's restatement."""
    for n, first, m in ((1, 0, 1), (5, 0, 5), (17, 3, 14), (4097, 0, 4097), (100003, 777, 5000), (31_000_000, 0, 20000),
                        ((1 << 40) + 12345, 1 << 39, 3000)):
        key = (0x9E3779B9, n & 0xFFFFFFFF)
        got = np.zeros(m, np.int64)
        shim.shim_random_subset(C.c_uint32(key[0]), C.c_uint32(key[1]), C.c_uint64(n), C.c_uint64(first), C.c_int64(m), _p(got))
        np.testing.assert_array_equal(got, O.random_subset(key, n, first, m))
        assert len(np.unique(got)) == m and got.min() >= 0 and got.max() < n


def test_device_rng_on_host_matches_oracle(shim):
    out = np.zeros(2, np.uint32)
    shim.shim_threefry(C.c_uint32(0x13198A2E), C.c_uint32(0x03707344), C.c_uint32(0x243F6A88), C.c_uint32(0x85A308D3), _p(out))
    assert out.tolist() == [0xC4923A9C, 0x483DF7A0]
    key = np.array([0xDEADBEEF, 0x12345678], np.uint32)
    for mode in (0, 1):
        for n in (1, 2, 3, 4, 7, 100, 1001):
            got = np.zeros((n, 2), np.uint32)
            shim.shim_split(_p(key), C.c_uint32(n), mode, _p(got))
            np.testing.assert_array_equal(got, CO.split(key, n, mode))


@pytest.mark.parametrize("mode", [0, 1])
def test_device_env_logic_on_host_matches_oracle(shim, mode):
    rng = np.random.default_rng(1 + mode)
    n = 60000
    boards = rng.integers(0, 8, (n, 16)) * (rng.random((n, 16)) < 0.7)
    boards[:3000] = rng.integers(1, 6, (3000, 16))
    boards[3000:6000] = rng.integers(0, 3, (3000, 16))
    boards[6000:6050] = rng.integers(9, 14, (50, 16))
    acts = rng.integers(0, 4, n).astype(np.int32)
    # bare move + exact legal mask
    packed = _pack(boards)
    moved = np.zeros(n, np.uint64)
    rew = np.zeros(n, np.uint32)
    lm = np.zeros(n, np.uint8)
    shim.shim_move(_p(packed), _p(acts), C.c_int64(n), _p(moved), _p(rew), _p(lm))
    want_b, want_r = O.move(boards, acts)
    np.testing.assert_array_equal(_unpack(moved), want_b)
    np.testing.assert_array_equal(rew, want_r)
    legal = O.exact_legal(boards)
    np.testing.assert_array_equal(((lm[:, None] >> np.arange(4)) & 1).astype(bool), legal)
    # full step incl. illegal actions, terminal boards and frozen envs
    done = ~legal.any(1)
    masks = np.where(done[:, None], True, legal)
    status = ((masks * (1 << np.arange(4))).sum(1) | np.where(done, 16, 0)).astype(np.uint8)
    sub = np.array([123, 456], np.uint32)
    keys = CO.split(sub, n, mode)
    wb, wm, wd, wr = CO.env_step(boards, masks, done, acts, keys, mode)
    gb, gs, gr = packed.copy(), status.copy(), np.zeros(n, np.float32)
    shim.shim_env_step(_p(gb), _p(gs), _p(acts), _p(sub), C.c_uint32(n), C.c_uint32(0), C.c_int64(n), mode, _p(gr))
    np.testing.assert_array_equal(_unpack(gb), wb)
    np.testing.assert_array_equal(gr, wr)
    np.testing.assert_array_equal((gs >> 4) & 1, wd)
    np.testing.assert_array_equal((gs[:, None] >> np.arange(4)) & 1, wm)
    assert (wr == -1).sum() > 50 and wd.sum() > 50
    # init
    ib, ist = np.zeros(2000, np.uint64), np.zeros(2000, np.uint8)
    shim.shim_env_init(_p(sub), C.c_uint32(2000), C.c_uint32(0), C.c_int64(2000), mode, _p(ib), _p(ist))
    wb, wm = CO.env_init(CO.split(sub, 2000, mode), mode)
    np.testing.assert_array_equal(_unpack(ib), wb)
    np.testing.assert_array_equal((ist[:, None] >> np.arange(4)) & 1, wm)
    # policies
    ga, glp = np.zeros(n, np.int32), np.zeros(n, np.float32)
    shim.shim_act(0, _p(status), _p(sub), C.c_uint32(n), C.c_uint32(0), C.c_int64(n), mode, _p(ga), _p(glp))
    wa, wlp = CO.act(keys, masks, CO.RANDOM, mode)
    np.testing.assert_array_equal(ga, wa)
    np.testing.assert_allclose(glp, wlp, rtol=1e-6)
    shim.shim_act(1, _p(status), _p(sub), C.c_uint32(n), C.c_uint32(0), C.c_int64(n), mode, _p(ga), _p(glp))
    np.testing.assert_array_equal(ga, CO.act(None, masks, CO.DRUL, mode)[0])


def test_device_spawn_on_host_given_draws(shim):
    rng = np.random.default_rng(4)
    n = 50000
    boards = rng.integers(0, 6, (n, 16)) * (rng.random((n, 16)) < 0.6)
    legal = O.exact_legal(boards)
    done = ~legal.any(1)
    masks = np.where(done[:, None], True, legal)
    status = ((masks * (1 << np.arange(4))).sum(1) | np.where(done, 16, 0)).astype(np.uint8)
    acts = rng.integers(0, 4, n).astype(np.int32)
    bp = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    bv = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    bp[:8] = [0, 0xFFFFFFFF, 0x1FF, 0x200, 0x3FF, 0x400, 0x80000000, 0x7FFFFE00]
    bv[:8] = [0, 0xFFFFFFFF, 0x19999800, 0x19999A00, 0x199999FF, 0x19999C00, 0x19999600, 0x1999A000]
    wb, wm, wd, wr = CO.env_step_given(boards, masks, done, acts, O._bits_to_unit_float(bp), O._bits_to_unit_float(bv))
    gb, gs, gr = _pack(boards), status.copy(), np.zeros(n, np.float32)
    shim.shim_env_step_draws(_p(gb), _p(gs), _p(acts), _p(bp), _p(bv), C.c_int64(n), _p(gr))
    np.testing.assert_array_equal(_unpack(gb), wb)
    np.testing.assert_array_equal(gr, wr)
    np.testing.assert_array_equal((gs >> 4) & 1, wd)


@pytest.mark.parametrize("mode", [0, 1])
def test_device_logit_sampling_on_host(shim, mode, golden_ppo):
    rng = np.random.default_rng(6)
    n = 40000
    logits = (rng.standard_normal((n, 4)) * 3).astype(np.float32)
    masks = rng.random((n, 4)) < 0.6
    masks[masks.sum(1) == 0, 3] = True
    status = (masks * (1 << np.arange(4))).sum(1).astype(np.uint8)
    sub = np.array([99, 1234], np.uint32)
    keys = CO.split(sub, n, mode)
    wa, wlp = O.act_from_logits((keys[:, 0], keys[:, 1]), O.mask_logits(logits, masks), mode, sample=True)
    ga, glp, gent = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    shim.shim_sample(_p(logits), _p(status), 1, 1, _p(sub), C.c_uint32(n), C.c_uint32(0), C.c_int64(n), mode, _p(ga), _p(glp), _p(gent))
    agree = ga == wa
    assert agree.mean() > 0.9999 and masks[np.arange(n), ga].all()
    np.testing.assert_allclose(glp[agree], wlp[agree], rtol=1e-5, atol=1e-6)
    dist = torch.distributions.Categorical(logits=torch.from_numpy(O.mask_logits(logits, masks)))
    np.testing.assert_allclose(gent, dist.entropy().numpy(), rtol=1e-5, atol=1e-6)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/g2048.h must compile as C99 and as C++ on its own."""
    import shutil
    import subprocess
    from pathlib import Path

    header = Path(__file__).resolve().parent.parent / "include" / "g2048.h"
    for compiler, flags in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++17", "-x", "c++"])):
        if shutil.which(compiler) is None:
            pytest.skip(f"{compiler} not available")
        res = subprocess.run([compiler, *flags, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", str(header)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_reference_learner_modules_are_found_behind_the_drop_in(tmp_path):
    """G2048_REFERENCE_ROOT: modules this package does not provide (the reference's learner) are imported from a
    reference checkout, while every module on the hot path still resolves to this package."""
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    fake = tmp_path / "checkout"
    (fake / "src" / "ppo").mkdir(parents=True)
    (fake / "src" / "optim").mkdir()
    (fake / "src" / "ppo" / "ppo_agent.py").write_text("from ..env_definitions import OBS_DIM\nfrom .rollout_buffer import RolloutBuffer\nWHO = 'reference learner'\n")
    (fake / "src" / "ppo" / "rollout_buffer.py").write_text("raise RuntimeError('the drop-in must shadow this module')\n")
    (fake / "src" / "optim" / "__init__.py").write_text("NAME = 'reference optim'\n")
    code = ("import src.ppo.ppo_agent as a, src.optim as o, src.ppo.rollout_buffer as r\n"
            "assert a.WHO == 'reference learner' and a.OBS_DIM == 31 and o.NAME == 'reference optim'\n"
            "assert a.RolloutBuffer is r.RolloutBuffer and 'g2048' in r.RolloutBuffer.__module__\n"
            "print('ok')\n")
    env = dict(__import__("os").environ, G2048_REFERENCE_ROOT=str(fake), PYTHONPATH=str(root / "2048-ppo-agent_b200"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stderr[-2000:]


def test_batched_fetch_collates_like_the_default_collate():
    """_LazySamples / _collate (what create_ppo_dataloader's DataLoader uses) against torch's default collation of the
    per-sample dicts, on a stand-in dataset (PPODataset itself needs the GPU for its GAE)."""
    import torch
    from torch.utils.data import DataLoader, Dataset

    from g2048.ppo.data_loader import _collate, _LazySamples

    class Stub(Dataset):
        def __init__(self):
            g = torch.Generator().manual_seed(0)
            self.fields = {"observations": torch.rand(100, 16, 31, generator=g), "actions": torch.rand(100, 4, generator=g),
                           "terminations": torch.rand(100, generator=g) < 0.1, "returns": torch.rand(100, generator=g)}
            self.active_indices = torch.randperm(100, generator=g)[:40]

        def __len__(self):
            return 40

        def _item(self, pos):
            return {k: v[pos] for k, v in self.fields.items()}

        def __getitem__(self, i):
            return self._item(self.active_indices[i])

        def __getitems__(self, indices):
            return _LazySamples(self, self.active_indices[torch.as_tensor(indices, dtype=torch.long)])

    ds = Stub()
    fast = DataLoader(ds, batch_size=16, shuffle=False, collate_fn=_collate)
    slow = DataLoader(ds, batch_size=16, shuffle=False)  # default collate iterates the lazy samples one by one
    n = 0
    for a, b in zip(fast, slow):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
        n += a["returns"].shape[0]
    assert n == 40
    lazy = ds.__getitems__([3, 1, 2])
    assert len(lazy) == 3 and torch.equal(lazy[1]["returns"], ds[1]["returns"])
    assert torch.equal(_collate([ds[0], ds[1]])["returns"], torch.stack([ds[0]["returns"], ds[1]["returns"]]))
