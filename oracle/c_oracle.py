"""ctypes binding of oracle/pgx2048_oracle.c.  TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_ref" / "libpgx2048_oracle.so"
_lib = None

ORIGINAL, PARTITIONABLE = 0, 1
RANDOM, DRUL = 0, 1


def build(force: bool = False) -> Path:
    src = _DIR / "pgx2048_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_play.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def threefry2x32(k0, k1, x0, x1):
    out = np.zeros(2, np.uint32)
    lib().orc_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(x0), C.c_uint32(x1), _p(out))
    return int(out[0]), int(out[1])


def split(key, n, mode):
    key = np.asarray(key, np.uint32)
    out = np.zeros((n, 2), np.uint32)
    lib().orc_split(_p(key), C.c_uint32(n), C.c_int(mode), _p(out))
    return out


def chain(key, mode, n_sub):
    """Returns (advanced key, subs (n_sub, 2))."""
    key = np.array(key, np.uint32)
    subs = np.zeros((n_sub, 2), np.uint32)
    lib().orc_chain(_p(key), C.c_int(mode), C.c_int64(n_sub), _p(subs))
    return key, subs


def env_init(keys, mode):
    keys = np.ascontiguousarray(keys, np.uint32)
    n = keys.shape[0]
    boards = np.zeros((n, 16), np.uint8)
    masks = np.zeros((n, 4), np.uint8)
    lib().orc_env_init(_p(keys), C.c_int64(n), C.c_int(mode), _p(boards), _p(masks))
    return boards, masks


def spawn_draws(keys, mode):
    keys = np.ascontiguousarray(keys, np.uint32)
    n = keys.shape[0]
    u_pos = np.zeros(n, np.float32)
    u_val = np.zeros(n, np.float32)
    lib().orc_spawn_draws(_p(keys), C.c_int64(n), C.c_int(mode), _p(u_pos), _p(u_val))
    return u_pos, u_val


def env_step_given(boards, masks, done, actions, u_pos, u_val):
    """In place on copies; returns (boards, masks, done, rewards)."""
    boards = np.array(boards, np.uint8)
    masks = np.array(masks, np.uint8)
    done = np.array(done, np.uint8)
    actions = np.ascontiguousarray(actions, np.int32)
    u_pos = np.ascontiguousarray(u_pos, np.float32)
    u_val = np.ascontiguousarray(u_val, np.float32)
    n = boards.shape[0]
    rewards = np.zeros(n, np.float32)
    lib().orc_env_step_given(_p(boards), _p(masks), _p(done), _p(actions), _p(u_pos), _p(u_val), C.c_int64(n), _p(rewards))
    return boards, masks, done, rewards


def env_step(boards, masks, done, actions, keys, mode):
    boards = np.array(boards, np.uint8)
    masks = np.array(masks, np.uint8)
    done = np.array(done, np.uint8)
    actions = np.ascontiguousarray(actions, np.int32)
    keys = np.ascontiguousarray(keys, np.uint32)
    n = boards.shape[0]
    rewards = np.zeros(n, np.float32)
    lib().orc_env_step(_p(boards), _p(masks), _p(done), _p(actions), _p(keys), C.c_int64(n), C.c_int(mode), _p(rewards))
    return boards, masks, done, rewards


def act(keys, masks, policy, mode):
    masks = np.ascontiguousarray(masks, np.uint8)
    n = masks.shape[0]
    keys = np.ascontiguousarray(keys if keys is not None else np.zeros((n, 2)), np.uint32)
    actions = np.zeros(n, np.int32)
    log_probs = np.zeros(n, np.float32)
    lib().orc_act(_p(keys), _p(masks), C.c_int64(n), C.c_int(policy), C.c_int(mode), _p(actions), _p(log_probs))
    return actions, log_probs


def play(seed, batch, policy, mode, env_lo=0, env_hi=None, max_steps=4096, first_actions=False, key=None):
    """Play envs [env_lo, env_hi) of `BatchRunner(seed).run_*(batch)` to termination.

    Returns dict(final_boards (n,16) u8, lengths, scores, loop_steps, key (advanced chain key)).
    """
    env_hi = batch if env_hi is None else env_hi
    n = env_hi - env_lo
    if key is None:
        key = np.array([(int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF], np.uint32)
    _, subs = chain(key, mode, 1 + 2 * max_steps)
    boards = np.zeros((n, 16), np.uint8)
    lengths = np.zeros(n, np.int32)
    scores = np.zeros(n, np.int64)
    fa = np.full((n, 16), -1, np.int32) if first_actions else None
    longest = lib().orc_play(
        _p(subs), C.c_int64(max_steps), C.c_int64(batch), C.c_int64(env_lo), C.c_int64(env_hi),
        C.c_int(policy), C.c_int(mode), _p(boards), _p(lengths), _p(scores), _p(fa) if fa is not None else None,
    )
    if longest < 0:
        raise RuntimeError("oracle play: max_steps too small")
    return dict(final_boards=boards, lengths=lengths, scores=scores, first_actions=fa, longest=int(longest))


def gae(rewards, values, dones, gamma=0.99, lambda_gae=0.95):
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    d = np.ascontiguousarray(dones, np.uint8)
    adv = np.zeros_like(r)
    ret = np.zeros_like(r)
    lib().orc_gae(_p(r), _p(v), _p(d), C.c_int64(r.shape[0]), C.c_double(gamma), C.c_double(lambda_gae), _p(adv), _p(ret))
    return adv, ret


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))
