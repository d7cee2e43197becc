"""Fixed-horizon rollouts of a persistent, auto-resetting batch of envs (BASELINE config C3).

The reference always plays a fresh batch to termination (src/runs/batch_runner.py:117) and its GAE never
bootstraps (src/ppo/data_loader.py:103-130); SURVEY 7 lists the standard PPO collection mode -- B envs that live
across iterations, T steps per iteration, ``pgx.experimental.auto_reset`` semantics, a bootstrap value for the
unfinished episodes -- as the new mode the fused kernels make cheap.  This is its host side:

    runner = FixedHorizonRunner(init_seed=0, act_fn=TorchActionFunction(agent, use_mask=True, device="cuda"),
                                batch_size=65536)
    rollout = runner.collect(128)                       # (T, n) time-major packed records on the device
    adv, ret, moments = rollout.gae(0.99, 0.95, bootstrap=runner.bootstrap_values())

Per step: the network forward (PyTorch) and ONE ``g2048_policy_step_obs`` launch (mask rule, sampling with
jax-compatible draws, log-prob, env step, auto-reset, record write and the observation of the next forward pass, written
from registers); with ``cuda_graph=True`` the step is captured once and replayed.  The key chain advances like the reference's: one sub key for the initial
``env.init``, then two per step (act, step).  ``state_dict()`` holds the key chain AND the env state (SURVEY 8f
rank 4: the reference checkpoints neither, src/ppo/ppo_trainer.py:511-530).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import engine as E
from .batch_runner import BatchRunner


@dataclass
class FixedRollout:
    """T steps of n envs, time-major: boards int64 (pre-step), meta uint8 (action | mask << 2 | done << 6),
    rewards / log_probs / values float32; final_boards / final_status: the state after the last step."""

    boards: torch.Tensor
    meta: torch.Tensor
    rewards: torch.Tensor
    log_probs: torch.Tensor
    values: torch.Tensor
    final_boards: torch.Tensor
    final_status: torch.Tensor
    t_steps: int
    n_envs: int

    def gae(self, gamma: float = 0.99, lambda_gae: float = 0.95, bootstrap: torch.Tensor | None = None):
        """-> advantages, returns (T, n) float32 and the fp64 moment block; `bootstrap` = V(state after step T)
        for the episodes still running (None: treat them as finished, the reference's behaviour)."""
        return E.gae_time_major(self.rewards, self.values, self.meta, self.t_steps, self.n_envs, bootstrap, gamma,
                                lambda_gae)


class FixedHorizonRunner:
    def __init__(self, init_seed: int, act_fn, batch_size: int, rng_mode=None, device=None,
                 shard: tuple[int, int] | None = None, cuda_graph: bool = False):
        if not hasattr(act_fn, "forward_logits"):
            raise ValueError("FixedHorizonRunner needs a TorchActionFunction (a policy network)")
        self._runner = BatchRunner(init_seed, act_fn, rng_mode=rng_mode, device=device, shard=shard, cuda_graph=cuda_graph)
        self._runner._check(batch_size)
        self.batch_size = batch_size
        self.lo, self.n = self._runner._range(batch_size)
        chain = self._runner.chain
        self.boards, self.status = E.env_init(chain.peek(1)[0], batch_size, self.lo, self.n, self._runner.rng_mode)
        chain.consume(1)
        self._graph = None

    @property
    def act_fn(self):
        return self._runner.act_fn

    @property
    def key(self):
        return self._runner.key

    def _records(self, t_steps: int):
        dev, n = self._runner.device, self.n
        return (torch.empty((t_steps, n), dtype=torch.int64, device=dev), torch.empty((t_steps, n), dtype=torch.uint8, device=dev),
                torch.empty((t_steps, n), dtype=torch.float32, device=dev), torch.empty((t_steps, n), dtype=torch.float32, device=dev),
                torch.empty((t_steps, n), dtype=torch.float32, device=dev))

    def collect(self, t_steps: int) -> FixedRollout:
        if t_steps <= 0:
            raise ValueError("t_steps must be positive")
        r, fn = self._runner, self._runner.act_fn
        mode = r.rng_mode
        subs = r.chain.peek(2 * t_steps)
        if r.cuda_graph:
            g = r._captured_step(self.batch_size, self.lo, self.n, steps=t_steps, auto_reset=True)
            if g is not self._graph:  # first use of this graph: the env state moves into its static buffers
                g["boards"].copy_(self.boards)
                g["status"].copy_(self.status)
                self.boards, self.status, self._graph = g["boards"], g["status"], g
            g["subs"].copy_(subs)
            g["step_index"].zero_()
            g["refresh_obs"]()  # the env state may have been replaced since the last replay (load_state_dict)
            for _ in range(t_steps):
                g["graph"].replay()
            rb, rm, rr, rl, rv = (g[k].clone() for k in ("rb", "rm", "rr", "rl", "rv"))
        else:
            rb, rm, rr, rl, rv = self._records(t_steps)
            obs = E.expand_obs(self.boards, fn.obs_dtype)
            for t in range(t_steps):  # per step: the forward pass and ONE launch (step + record + next observation)
                logits, values = fn.forward_logits(obs)
                E.policy_step_obs(self.boards, self.status, logits, values, fn.use_mask, fn.sample_actions, True, subs[2 * t:],
                                  None, self.batch_size, self.lo, mode, obs, rb[t], rm[t], rr[t], rl[t], rv[t])
        r.chain.consume(2 * t_steps)
        return FixedRollout(rb, rm, rr, rl, rv, self.boards.clone(), self.status.clone(), t_steps, self.n)

    def bootstrap_values(self) -> torch.Tensor:
        """V(current state) from the policy network, (n,) float32: the bootstrap of the running episodes."""
        fn = self._runner.act_fn
        _, values = fn.forward_logits(E.expand_obs(self.boards, fn.obs_dtype))
        return values

    # -- checkpointing: key chain + env state ------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"runner": self._runner.state_dict(), "batch_size": self.batch_size, "lo": self.lo,
                "boards": self.boards.cpu().clone(), "status": self.status.cpu().clone()}

    def load_state_dict(self, state: dict) -> None:
        if state["batch_size"] != self.batch_size or state["lo"] != self.lo or state["boards"].shape[0] != self.n:
            raise ValueError("checkpoint was written for a different batch / shard")
        self._runner.load_state_dict(state["runner"])
        self.boards.copy_(state["boards"].to(self.boards.device))
        self.status.copy_(state["status"].to(self.status.device))
