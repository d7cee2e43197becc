"""BASELINE.json config C4: one full PPO iteration's data path at 262 144 envs, phase by phase, on one GPU.

    rollout (recorded, to termination)  -> BatchRunner.run_packed_batch        g2048_rollout_steps
    store                               -> RolloutBuffer.store_packed          episode_lengths / scan / compact_records
    GAE + normalisation                 -> DevicePPOBatches (compute_gae)      g2048_gae_flat + g2048_normalize
    4 epochs of 2048-sample minibatches -> DevicePPOBatches.__iter__           g2048_gather_minibatch
    (policy update)                     -> a small stand-in MLP in PyTorch: the learner is outside the product path

plus the "bit-exact board check vs Pgx on a 4 096-env sample with actions and spawn draws recorded" that C4 asks
for, against the oracle.  Prints one JSON object.  Usage: python tools/c4_iteration.py [--envs 262144]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

import numpy as np
import torch

import g2048
from g2048 import engine as E


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262144)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatch", type=int, default=2048)
    ap.add_argument("--max-samples", type=int, default=300000)  # configs/trainer/default.yaml max_samples_per_epoch
    args = ap.parse_args()
    report = {"config": f"C4: {args.envs} envs to termination, random policy records, {args.epochs} epochs x minibatch {args.minibatch}"}

    runner = g2048.BatchRunner(init_seed=4, act_fn=g2048.act_randomly)
    warm = runner.run_packed_batch(1024)  # warm-up: table build, allocator, lazy loading of every kernel on the path
    wb = g2048.RolloutBuffer(31, 16, 4)
    wb.store_packed(warm)
    for _ in g2048.DevicePPOBatches(wb.get_packed(), batch_size=256):
        break
    rollout, t_roll = timed(lambda: runner.run_packed_batch(args.envs))
    # A/B on throw-away runners, each warmed up at full size (allocator, lazy loading): every env stepped until the
    # last one ends (the reference's loop) against only the live ones
    ab = {}
    for name, flag in (("all_envs", False), ("live_envs_only", True)):
        r2 = g2048.BatchRunner(init_seed=5, act_fn=g2048.act_randomly, compact_live=flag)
        r2.run_packed_batch(args.envs)
        ro2, ab[name] = timed(lambda: r2.run_packed_batch(args.envs))
        ab[name + "_env_steps"] = ro2.env_steps
        del ro2, r2
    assert ab["all_envs_env_steps"] == ab["live_envs_only_env_steps"]
    t_roll_live = ab["live_envs_only"]
    report["rollout"] = {"seconds_first_call_at_this_size": t_roll, "seconds": ab["all_envs"], "loop_steps": rollout.t_steps,
                         "env_steps": rollout.env_steps, "env_steps_per_sec": rollout.env_steps / ab["all_envs"],
                         "record_bytes": int(rollout.t_steps * args.envs * 17),
                         "seconds_warm_all_envs": ab["all_envs"], "seconds_warm_live_envs_only": t_roll_live,
                         "live_fraction": rollout.env_steps / (rollout.t_steps * args.envs)}

    buf = g2048.RolloutBuffer(31, 16, 4)
    kept, t_store = timed(lambda: buf.store_packed(rollout))
    report["store_packed"] = {"seconds": t_store, "steps_kept": kept, "steps_per_sec": kept / t_store}
    packed = buf.get_packed()
    packed["values"].copy_(torch.randn_like(packed["values"]))  # stand-in critic outputs

    batches, t_gae = timed(lambda: g2048.DevicePPOBatches(packed, 0.99, 0.95, batch_size=args.minibatch,
                                                          max_samples_per_epoch=args.max_samples, shuffle_on_reset=True))
    report["gae_normalise"] = {"seconds": t_gae, "steps_per_sec": kept / t_gae}

    class StandIn(torch.nn.Module):
        """Per-cell input embedding like the reference agent's (Linear(31, d_model, bias=False), ppo_agent.py:60),
        mean over the 16 cells, one hidden layer, 4 logits + 1 value.  Takes observations or bitboards."""

        def __init__(self):
            super().__init__()
            self.input_embedding = torch.nn.Linear(31, 256, bias=False)
            self.head = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 5))

        def forward(self, batch):
            if "boards" in batch:
                emb = g2048.ppo.embed_boards(self.input_embedding.weight, batch["boards"])
            else:
                emb = self.input_embedding(batch["observations"])
            return self.head(emb.mean(dim=1))

    net = StandIn().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)

    def epochs(update: bool, source=None):
        source = batches if source is None else source
        n = 0
        for _ in range(args.epochs):
            source.reset_epoch()
            for b in source:
                n += b["actions"].shape[0]
                if update:
                    out = net(b)
                    logits = out[:, :4] - 1e8 * (1 - b["action_masks"].float())
                    logp = torch.distributions.Categorical(logits=logits).log_prob(b["actions"])
                    ratio = torch.exp(logp - b["log_probs"])
                    loss = -(torch.min(ratio * b["advantages"], ratio.clamp(0.8, 1.2) * b["advantages"])).mean() \
                        + 0.5 * ((out[:, 4] - b["returns"]) ** 2).mean()
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    opt.step()
        return n

    n_samples, t_feed = timed(lambda: epochs(False))
    batches_r = g2048.DevicePPOBatches(packed, 0.99, 0.95, batch_size=args.minibatch, max_samples_per_epoch=args.max_samples,
                                       shuffle_on_reset=True, reuse_buffers=True)
    epochs(False, batches_r)
    _, t_feed_r = timed(lambda: epochs(False, batches_r))
    _, t_update = timed(lambda: epochs(True))
    report["minibatches"] = {"seconds_gather_only": t_feed, "samples": n_samples, "samples_per_sec": n_samples / t_feed,
                             "seconds_with_stand_in_update": t_update, "seconds_gather_only_reused_buffers": t_feed_r}
    # the same epochs with bitboards instead of observations: the embedding becomes a row gather
    batches_b = g2048.DevicePPOBatches(packed, 0.99, 0.95, batch_size=args.minibatch, max_samples_per_epoch=args.max_samples,
                                       shuffle_on_reset=True, obs_dtype=None, reuse_buffers=True)
    epochs(True, batches_b)  # warm-up of the embedding kernels
    _, t_feed_b = timed(lambda: epochs(False, batches_b))
    _, t_update_b = timed(lambda: epochs(True, batches_b))
    report["minibatches_boards"] = {"seconds_gather_only": t_feed_b, "seconds_with_stand_in_update": t_update_b,
                                    "bytes_per_sample": 8 + 8 + 4 + 16, "bytes_per_sample_observations": 1984 + 8 + 4 + 16}

    # bit-exact board check on a 4 096-env sample: replay the recorded actions through the oracle's step, with the
    # oracle's own spawn draws from the same keys, and compare every recorded pre-step board, reward and done flag
    from oracle import c_oracle as CO

    sample = 4096
    t_chk = min(rollout.t_steps, 200)
    # the warm-up batch advanced the key chain: a fresh runner with the same seed replays it to get the chain key
    ref_runner = g2048.BatchRunner(init_seed=4, act_fn=g2048.act_randomly)
    ref_runner.run_packed_batch(1024)
    chain_key = ref_runner.key
    _, subs = CO.chain(chain_key, 1, 1 + 2 * t_chk)
    boards, masks = CO.env_init(CO.split(subs[0], args.envs, 1)[:sample], 1)
    done = np.zeros(sample, np.uint8)
    rec_b = E.boards_numpy(rollout.boards[:t_chk, :sample])
    rec_m = rollout.meta[:t_chk, :sample].cpu().numpy()
    rec_r = rollout.rewards[:t_chk, :sample].cpu().numpy()
    ok = True
    for t in range(t_chk):
        ok &= np.array_equal(rec_b[t], boards)
        actions = (rec_m[t] & 3).astype(np.int32)
        step_keys = CO.split(subs[2 + 2 * t], args.envs, 1)[:sample]
        boards, masks, done, rew = CO.env_step(boards, masks, done, actions, step_keys, 1)
        ok &= np.array_equal(rew, rec_r[t]) and np.array_equal(done, (rec_m[t] >> 6) & 1)
    report["bit_exact_board_check"] = {"envs": sample, "steps": t_chk, "identical": bool(ok)}
    t_roll = ab["all_envs"]  # warm; the first call at a new size also pays cudaMalloc for ~2 GB of record buffers
    total = t_roll + t_store + t_gae + t_feed
    report["product_path_seconds"] = total
    report["share"] = {k: round(v / total, 3) for k, v in (("rollout", t_roll), ("store", t_store), ("gae", t_gae), ("minibatches", t_feed))}
    print(json.dumps(report))
    assert ok, "recorded rollout differs from the oracle"


if __name__ == "__main__":
    main()
