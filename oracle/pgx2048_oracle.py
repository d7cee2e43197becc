"""CPU restatement (numpy) of the reference's rollout hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path (the ``g2048``
package and ``libg2048.so``) never does.

Parity status: PINNED.  The restatement reproduces, bit for bit, the artefacts the real
Pgx/JAX path left in the reference repo (decoded into ``tests/golden/``):
  * ``assets/2048_drul_actions.svg`` / ``assets/2048_random_actions.svg`` (1 632 boards)
    under the original Threefry counter layout (``jax_threefry_partitionable=False``);
  * the 1000-episode max-tile histograms of ``assets/{random,drul}_strategy_statistics.png``
    under the partitionable layout (default of the pinned ``jax==0.5.3``).
Rows that no artefact pins (per-step reward values, terminal mask all-True, the
illegal-action path, zero reward on a finished env) follow the published Pgx 2.6.0
algorithm as recalled in SURVEY.md Appendix A and are labelled "recalled" in DESIGN.md.

What is restated, and the reference lines each piece follows (relative to the
reference repo root):
  * ``jax.random`` (third party, jax==0.5.3, uv.lock:701-702): key/split/bits/uniform/
    choice/categorical on Threefry-2x32 -- call sites src/runs/batch_runner.py:32,105-106,
    118-119,126-127; src/actions/act_randomly.py:48; src/ppo/torch_action_wrapper.py:91.
  * Pgx "2048" env (third party, pgx==2.6.0, uv.lock:1564-1565): init/step/observe/
    legal mask/terminal rules -- call sites src/runs/batch_runner.py:33-35,107,128.
  * act_randomly   src/actions/act_randomly.py:40-56
  * act_drul       src/actions/act_drul.py:40-49
  * runner loops   src/runs/batch_runner.py:105-154,176-195; src/runs/run_actions_batch.py:41-57
  * max-tile run   src/runs/run_actions_max_tile.py:42-71
  * RolloutBuffer  src/ppo/rollout_buffer.py:164-206
  * GAE + normalisation  src/ppo/data_loader.py:61-67,103-130
  * RunningStatsVec      src/stats/running_stats_vec.py:55-87
  * policy-logit sampling / log-prob  src/ppo/torch_action_wrapper.py:84-102,
    src/ppo/ppo_agent.py:117-121
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

U32 = np.uint32
ORIGINAL = 0  # jax_threefry_partitionable=False (jax < 0.5 default) -- matches the SVGs
PARTITIONABLE = 1  # jax_threefry_partitionable=True (jax 0.5.3 default) -- matches the PNGs

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_FLT_TINY = np.float32(np.finfo(np.float32).tiny)
_FLT_LOWEST = np.float32(np.finfo(np.float32).min)


# --------------------------------------------------------------------------- Threefry
def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Random123).  All arguments broadcast as uint32."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
        return x0, x1


def key_from_seed(seed: int):
    """jax.random.key(seed) -> (hi, lo) words."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return U32(seed >> 32), U32(seed & 0xFFFFFFFF)


def split(key, n: int, mode: int):
    """jax.random.split(key, n) for a batch of keys.

    key: (k0, k1) arrays of shape S.  Returns (k0', k1') of shape S + (n,).
    """
    k0 = np.asarray(key[0], dtype=U32)[..., None]
    k1 = np.asarray(key[1], dtype=U32)[..., None]
    if mode == PARTITIONABLE:
        return threefry2x32(k0, k1, U32(0), np.arange(n, dtype=U32))
    c = np.arange(2 * n, dtype=U32)
    y0, y1 = threefry2x32(k0, k1, c[:n], c[n:])
    flat = np.concatenate([y0, y1], axis=-1)  # S + (2n,)
    return flat[..., 0::2], flat[..., 1::2]


def random_bits(key, m: int, mode: int):
    """32-bit draws of shape (m,) (m = 1 stands for the scalar shape ()) per key."""
    k0 = np.asarray(key[0], dtype=U32)[..., None]
    k1 = np.asarray(key[1], dtype=U32)[..., None]
    if mode == PARTITIONABLE:
        y0, y1 = threefry2x32(k0, k1, U32(0), np.arange(m, dtype=U32))
        return y0 ^ y1
    c = np.arange(m + (m & 1), dtype=U32)
    if m & 1:
        c[-1] = 0
    h = c.size // 2
    y0, y1 = threefry2x32(k0, k1, c[:h], c[h:])
    return np.concatenate([y0, y1], axis=-1)[..., :m]


def _bits_to_unit_float(bits):
    return ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)


def uniform(key, m: int, mode: int, lo=np.float32(0.0), hi=np.float32(1.0)):
    f = _bits_to_unit_float(np.ascontiguousarray(random_bits(key, m, mode)))
    lo = np.float32(lo)
    hi = np.float32(hi)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def categorical(key, logits, mode: int):
    """jax.random.categorical(key, logits) over the last axis (4 actions)."""
    logits = np.asarray(logits, dtype=np.float32)
    u = uniform(key, logits.shape[-1], mode, lo=_FLT_TINY, hi=np.float32(1.0))
    with np.errstate(divide="ignore"):
        g = -np.log(-np.log(u))
    return np.argmax(g.astype(np.float32) + logits, axis=-1).astype(np.int32), u


# --------------------------------------------------------------------------- Pgx 2048
def _row_left(rows):
    """rows: (N, 4) int32 exponents.  Returns (new_rows, reward)."""
    n = rows.shape[0]
    out = np.zeros_like(rows)
    reward = np.zeros(n, dtype=np.int64)
    for i in range(n):
        tiles = [int(v) for v in rows[i] if v]
        merged = []
        j = 0
        while j < len(tiles):
            if j + 1 < len(tiles) and tiles[j] == tiles[j + 1]:
                merged.append(tiles[j] + 1)
                reward[i] += 1 << (tiles[j] + 1)
                j += 2
            else:
                merged.append(tiles[j])
                j += 1
        out[i, : len(merged)] = merged
    return out, reward


_ROW_CACHE: dict = {}


def _row_tables():
    """Row-left result for every row of exponents 0..17 (base-18 index), built once."""
    if "t" not in _ROW_CACHE:
        base = 18
        idx = np.arange(base**4)
        rows = np.stack([(idx // base**k) % base for k in range(4)], axis=1).astype(np.int32)
        new, rew = _row_left(rows)
        _ROW_CACHE["t"] = (base, new.astype(np.int32), rew.astype(np.int64))
    return _ROW_CACHE["t"]


def move(boards, actions):
    """Slide/merge every board toward its action.  0=Left 1=Up 2=Right 3=Down.

    boards: (B, 16) int exponents, actions: (B,).  Returns (new_boards, reward int64).
    Pgx: rot90(board, k=action) -> row-left -> rot90(., -action).
    """
    boards = np.asarray(boards, dtype=np.int32).reshape(-1, 4, 4)
    actions = np.asarray(actions).reshape(-1)
    base, table, rtable = _row_tables()
    assert boards.max(initial=0) < base - 1, "oracle row table covers exponents < 17"
    out = np.empty_like(boards)
    reward = np.zeros(boards.shape[0], dtype=np.int64)
    for a in range(4):
        sel = np.nonzero(actions == a)[0]
        if sel.size == 0:
            continue
        rot = np.rot90(boards[sel], k=a, axes=(1, 2))
        rows = rot.reshape(-1, 4)
        idx = rows[:, 0] + base * (rows[:, 1] + base * (rows[:, 2] + base * rows[:, 3]))
        new = table[idx].reshape(-1, 4, 4)
        reward[sel] = rtable[idx].reshape(-1, 4).sum(axis=1)
        out[sel] = np.rot90(new, k=-a, axes=(1, 2))
    return out.reshape(-1, 16), reward


def exact_legal(boards):
    """legal[a] = (move(board, a) != board)."""
    boards = np.asarray(boards, dtype=np.int32).reshape(-1, 16)
    legal = np.zeros((boards.shape[0], 4), dtype=bool)
    for a in range(4):
        moved, _ = move(boards, np.full(boards.shape[0], a))
        legal[:, a] = (moved != boards).any(axis=1)
    return legal


def spawn_draws(key, mode: int):
    """The two uniforms _add_random_num consumes: k1,k2 = split(key); u_pos, u_val."""
    s0, s1 = split(key, 2, mode)
    u_pos = uniform((s0[..., 0], s1[..., 0]), 1, mode)[..., 0]
    u_val = uniform((s0[..., 1], s1[..., 1]), 1, mode)[..., 0]
    return u_pos, u_val


def add_random_given(boards, u_pos, u_val):
    """Pgx _add_random_num with the uniforms already drawn.

    pos = choice(arange(16), p=(board==0)):  cum = cumsum_f32(empty); r = cum[-1]*(1-u);
          searchsorted_left(cum, r).  val = [1,2][choice(p=[0.9,0.1])]: 2 iff 1*(1-u') > 0.9f.
    A full board (cum[-1] == 0) leaves searchsorted at index 0 -> cell 0 is overwritten,
    which Pgx can only reach through the illegal-action path on a full board.
    """
    boards = np.asarray(boards, dtype=np.int32).reshape(-1, 16).copy()
    empty = (boards == 0).astype(np.float32)
    cum = np.cumsum(empty, axis=1, dtype=np.float32)
    one = np.float32(1.0)
    r = cum[:, -1] * (one - np.asarray(u_pos, dtype=np.float32))
    pos = (cum < r[:, None]).sum(axis=1)  # searchsorted side='left'
    pos = np.minimum(pos, 15)
    cum_v = np.cumsum(np.array([0.9, 0.1], dtype=np.float32), dtype=np.float32)
    rv = cum_v[-1] * (one - np.asarray(u_val, dtype=np.float32))
    val = 1 + (cum_v[None, :] < rv[:, None]).sum(axis=1)
    boards[np.arange(boards.shape[0]), pos] = val
    return boards


def add_random(boards, key, mode: int):
    u_pos, u_val = spawn_draws(key, mode)
    return add_random_given(boards, u_pos, u_val)


@dataclass
class State:
    """Field names follow pgx.State (v2 API)."""

    board: np.ndarray  # (B, 16) int32 exponents
    legal_action_mask: np.ndarray  # (B, 4) bool
    rewards: np.ndarray  # (B, 1) float32
    terminated: np.ndarray  # (B,) bool
    truncated: np.ndarray  # (B,) bool

    @property
    def observation(self):
        return observe(self.board)


def observe(boards):
    """One-hot (B, 4, 4, 31) bool; empty cell -> channel 0."""
    boards = np.asarray(boards).reshape(-1, 16)
    obs = np.zeros((boards.shape[0], 16, 31), dtype=bool)
    np.put_along_axis(obs, boards[:, :, None].astype(np.int64), True, axis=2)
    return obs.reshape(-1, 4, 4, 31)


def env_init(key, mode: int) -> State:
    """Pgx 2048 _init: r1, r2 = split(key); two _add_random_num; exact legal mask."""
    s0, s1 = split(key, 2, mode)
    b = np.zeros((np.asarray(key[0]).shape[0], 16), dtype=np.int32)
    b = add_random(b, (s0[..., 0], s1[..., 0]), mode)
    b = add_random(b, (s0[..., 1], s1[..., 1]), mode)
    n = b.shape[0]
    return State(b, exact_legal(b), np.zeros((n, 1), np.float32), np.zeros(n, bool), np.zeros(n, bool))


def env_step_given(state: State, actions, u_pos, u_val) -> State:
    """pgx core.Env.step + 2048 _step with the spawn uniforms given.

    Already-terminated envs are returned frozen with zero reward; an action that is illegal
    under the PRE-step mask gives reward -1 and terminates (after the no-op move + spawn
    was applied); a terminal state's mask is all-True.
    """
    actions = np.asarray(actions).reshape(-1)
    done_before = state.terminated | state.truncated
    moved, rew = move(state.board, actions)
    new = add_random_given(moved, u_pos, u_val)
    legal = exact_legal(new)
    terminated = ~legal.any(axis=1)
    rewards = rew.astype(np.float32)
    illegal = ~state.legal_action_mask[np.arange(actions.size), actions]
    rewards = np.where(illegal, np.float32(-1.0), rewards)
    terminated = terminated | illegal
    legal = np.where(terminated[:, None], True, legal)
    # frozen envs
    new = np.where(done_before[:, None], state.board, new)
    legal = np.where(done_before[:, None], state.legal_action_mask, legal)
    rewards = np.where(done_before, np.float32(0.0), rewards)
    terminated = np.where(done_before, state.terminated, terminated)
    return State(new.astype(np.int32), legal, rewards.reshape(-1, 1).astype(np.float32), terminated, state.truncated.copy())


def env_step(state: State, actions, key, mode: int) -> State:
    u_pos, u_val = spawn_draws(key, mode)
    return env_step_given(state, actions, u_pos, u_val)


# --------------------------------------------------------------------------- policies
def act_randomly(key, mask, mode: int, shortcut: bool = False):
    """src/actions/act_randomly.py:40-56.  Returns (action int32, log_prob float32)."""
    mask = np.asarray(mask, dtype=bool)
    n = mask.sum(axis=1).astype(np.float32)
    probs = np.where(n[:, None] > 0, mask.astype(np.float32) / np.maximum(n, 1)[:, None], np.float32(0.25))
    with np.errstate(divide="ignore"):
        logits = np.maximum(np.log(probs.astype(np.float32)), _FLT_LOWEST)
    if shortcut:  # all legal logits are equal, gumbel is monotone in u
        u = uniform(key, 4, mode, lo=_FLT_TINY, hi=np.float32(1.0))
        allowed = np.where(n[:, None] > 0, mask, True)
        action = np.argmax(np.where(allowed, u, np.float32(-1.0)), axis=1).astype(np.int32)
    else:
        action, _ = categorical(key, logits, mode)
    with np.errstate(divide="ignore"):
        log_prob = np.log(probs[np.arange(mask.shape[0]), action]).astype(np.float32)
    return action, log_prob


def act_drul(mask):
    """src/actions/act_drul.py:40-44: first legal of [3, 2, 1, 0]; none legal -> 3."""
    mask = np.asarray(mask, dtype=bool)
    order = np.array([3, 2, 1, 0], dtype=np.int32)
    return order[mask[:, order].argmax(axis=1)]


def act_from_logits(key, logits, mode: int, sample: bool = True):
    """src/ppo/torch_action_wrapper.py:84-102 given the network's (already masked) logits."""
    logits = np.maximum(np.asarray(logits, dtype=np.float32), _FLT_LOWEST)
    if sample:
        action, _ = categorical(key, logits, mode)
    else:
        action = np.argmax(logits, axis=-1).astype(np.int32)
    m = logits.max(axis=-1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(axis=-1, keepdims=True, dtype=np.float32))).astype(np.float32)
    log_prob = logits[np.arange(logits.shape[0]), action] - lse[:, 0]
    return action, log_prob.astype(np.float32)


def mask_logits(logits, mask):
    """src/ppo/ppo_agent.py:117-121: logits - 1e8 * (1 - mask)."""
    return (np.asarray(logits, np.float32) - np.float32(1e8) * (np.float32(1.0) - np.asarray(mask, np.float32))).astype(np.float32)


# --------------------------------------------------------------------------- runner
class KeyChain:
    """The runner's global key chain: key, sub = split(key) (batch_runner.py:105,118,126)."""

    def __init__(self, seed: int, mode: int):
        self.key = key_from_seed(seed)
        self.mode = mode

    def next_subkey(self):
        k0, k1 = split((np.array([self.key[0]]), np.array([self.key[1]])), 2, self.mode)
        self.key = (k0[0, 0], k1[0, 0])
        return k0[0, 1], k1[0, 1]

    def next_batch_keys(self, batch: int):
        sub = self.next_subkey()
        k0, k1 = split((np.array([sub[0]]), np.array([sub[1]])), batch, self.mode)
        return k0[0], k1[0]


def rollout(chain: KeyChain, batch: int, policy: str, max_steps: int = 100000):
    """One BatchRunner pass (batch_runner.py:105-136).

    Returns dict with init state and per-loop-step lists: pre-step board/mask, action,
    log_prob, and the post-step State.
    """
    mode = chain.mode
    state = env_init(chain.next_batch_keys(batch), mode)
    out = {"init": state, "boards": [], "masks": [], "actions": [], "log_probs": [], "states": []}
    steps = 0
    while not (state.terminated | state.truncated).all():
        act_keys = chain.next_batch_keys(batch)
        if policy == "random":
            action, log_prob = act_randomly(act_keys, state.legal_action_mask, mode)
        elif policy == "drul":
            action, log_prob = act_drul(state.legal_action_mask), None
        else:
            raise ValueError(policy)
        step_keys = chain.next_batch_keys(batch)
        out["boards"].append(state.board)
        out["masks"].append(state.legal_action_mask)
        out["actions"].append(action)
        out["log_probs"].append(log_prob)
        state = env_step(state, action, step_keys, mode)
        out["states"].append(state)
        steps += 1
        if steps >= max_steps:
            raise RuntimeError("rollout did not terminate")
    return out


def episode_summary(out):
    """Per-env length (first done + 1), score (sum of positive rewards) and final max tile."""
    term = np.stack([s.terminated for s in out["states"]], axis=1)
    rew = np.concatenate([s.rewards for s in out["states"]], axis=1)
    length = term.argmax(axis=1) + 1
    score = np.where(rew > 0, rew, 0).sum(axis=1).astype(np.int64)
    final = out["states"][-1].board
    return length, score, (1 << final.max(axis=1)).astype(np.int64)


# --------------------------------------------------------------------------- buffer / GAE / stats
def store_batch_indices(terminations):
    """rollout_buffer.py:164-187: keep steps 0..first_done inclusive per env, env-major.

    Returns (env_idx, step_idx) of the kept steps in buffer order.
    """
    terminations = np.asarray(terminations, dtype=bool)
    envs, steps = [], []
    for b in range(terminations.shape[0]):
        idx = np.nonzero(terminations[b])[0]
        end = idx[0] + 1 if idx.size else 0
        envs.extend([b] * end)
        steps.extend(range(end))
    return np.asarray(envs, dtype=np.int64), np.asarray(steps, dtype=np.int64)


def gae_returns(rewards, values, terminations, gamma=0.99, lambda_gae=0.95):
    """data_loader.py:103-130, fp32 step by step (gamma, gamma*lambda rounded to f32 at the multiply)."""
    r = np.asarray(rewards, dtype=np.float32)
    v = np.asarray(values, dtype=np.float32)
    d = np.asarray(terminations, dtype=bool)
    adv = np.zeros_like(r)
    ret = np.zeros_like(r)
    g = np.float32(gamma)
    gl = np.float32(gamma * lambda_gae)
    last_gae = np.float32(0.0)
    last_value = np.float32(0.0)
    for t in range(r.shape[0] - 1, -1, -1):
        if d[t]:
            last_value = np.float32(0.0)
            last_gae = np.float32(0.0)
        delta = np.float32(np.float32(r[t] + np.float32(g * last_value)) - v[t])
        last_gae = np.float32(delta + np.float32(gl * last_gae))
        adv[t] = last_gae
        ret[t] = np.float32(last_gae + v[t])
        last_value = v[t]
    return adv, ret


def normalize(x):
    """data_loader.py:61-67: (x - mean) / (std_unbiased + 1e-8)."""
    x = np.asarray(x, dtype=np.float32)
    mean = x.astype(np.float64).mean()
    std = x.astype(np.float64).std(ddof=1)
    return ((x - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))).astype(np.float32)


def running_stats_merge(n_a, mean_a, var_a, n_b, mean_b, var_b):
    """running_stats_vec.py:74-87 (Chan merge of count / mean / population variance)."""
    total = n_a + n_b
    delta2 = (mean_b - mean_a) ** 2.0
    mean = (mean_a * n_a + mean_b * n_b) / total
    var = (var_b * n_b + n_a * var_a + delta2 * (n_a * n_b) / total) / total
    return total, mean, var


def random_subset(key, n: int, first: int, m: int):
    """The keyed bijection of [0, n) behind the product's g2048_random_subset, restated (the reference draws its
    subsets with torch.randperm(total_length)[:length], data_loader.py:73-101: same distribution, different stream):
    4-round balanced Feistel network over 2h bits, 4^h >= n, round function = low h bits of the first word of
    Threefry-2x32(key; (right half, round)), cycle walking for images >= n."""
    h = 1
    while h < 31 and (1 << (2 * h)) < n:
        h += 1
    mask = np.uint64((1 << h) - 1)
    x = np.arange(first, first + m, dtype=np.uint64)
    todo = np.ones(m, dtype=bool)
    while todo.any():
        cur = x[todo]
        left = (cur >> np.uint64(h)).astype(np.uint64)
        right = cur & mask
        for r in range(4):
            f = threefry2x32(key[0], key[1], right.astype(U32), np.full(right.shape, r, dtype=U32))[0].astype(np.uint64) & mask
            left, right = right, left ^ f
        cur = (left << np.uint64(h)) | right
        x[todo] = cur
        todo[todo] = cur >= np.uint64(n)
    return x.astype(np.int64)
