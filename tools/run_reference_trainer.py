"""The reference's UNMODIFIED learner (PPOTrainer, PPOAgent, TransformerEncoder, src/optim) on top of this package's
rollout path: one PPO iteration -- collect_rollouts (BatchRunner + TorchActionFunction + RolloutBuffer of this package)
and update_policy (create_ppo_dataloader / PPODataset of this package, GAE by g2048_gae_host) -- on cuda:0.

The reference's files are never added to this repository: `--stage` copies the learner modules from a checkout into
the git-ignored scratch directory refcheck/reference/ (which travels to the GPU box with the gpurun snapshot); a plain
run points G2048_REFERENCE_ROOT at it, so `src.ppo.ppo_trainer` is the reference's file and every `src.*` module on
the hot path is this package's (2048-ppo-agent_b200/src).

    python tools/run_reference_trainer.py --stage /root/reference
    gpurun -- python tools/run_reference_trainer.py [--envs 256] [--batches 2]
"""
import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SCRATCH = ROOT / "refcheck" / "reference"
FILES = ("src/ppo/ppo_agent.py", "src/ppo/transformer_encoder.py", "src/ppo/ppo_trainer.py", "src/optim/__init__.py",
         "src/optim/configure_optimizers.py", "src/optim/lamb.py")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", metavar="REFERENCE_ROOT")
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--minibatch", type=int, default=2048)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--device", default="cuda:0", help="where the reference's trainer keeps the agent (its default is cpu)")
    ap.add_argument("--default-agent", action="store_true",
                    help="configs/model/transformer_combined.yaml: d_model 256, 8 heads, 4 layers, ff 1024, hidden 512, cls")
    ap.add_argument("--no-mask", action="store_true")
    ap.add_argument("--train-timesteps", type=int, default=0,
                    help="then run PPOIterationLoop.for_trainer(trainer).train(...) -- this package's version of the trainer's "
                         "train loop around the reference's own update_policy -- for this many timesteps")
    args = ap.parse_args()
    if args.stage:
        for rel in FILES:
            dst = SCRATCH / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(Path(args.stage) / rel, dst)
        print(f"staged {len(FILES)} files under {SCRATCH} (git-ignored)")
        return 0
    if not (SCRATCH / "src" / "ppo" / "ppo_trainer.py").exists():
        print("nothing staged: run with --stage <reference checkout> first")
        return 2
    os.environ["G2048_REFERENCE_ROOT"] = str(SCRATCH)
    sys.path.insert(0, str(ROOT / "2048-ppo-agent_b200"))
    import torch

    from src.ppo.ppo_agent import PPOAgent          # the reference's file
    from src.ppo.ppo_trainer import PPOTrainer      # the reference's file
    from src.ppo.rollout_buffer import RolloutBuffer  # this package
    from src.runs.batch_runner import BatchRunner     # this package
    import src.ppo.ppo_trainer as trainer_module
    import src.runs.batch_runner as runner_module

    origin = {"PPOTrainer": trainer_module.__file__, "BatchRunner": runner_module.__file__}
    assert str(SCRATCH) in origin["PPOTrainer"] and "2048-ppo-agent_b200" in origin["BatchRunner"], origin

    torch.manual_seed(0)
    dev = torch.device(args.device)
    if args.default_agent:
        agent = PPOAgent(hidden_dim=512, d_model=256, nhead=8, num_layers=4, dim_feedforward=1024, dropout=0.1, reduction="cls")
    else:
        agent = PPOAgent(hidden_dim=128, d_model=64, nhead=4, num_layers=2, dim_feedforward=128, dropout=0.0, reduction="cls")
    optim = dict(opt_name="adamw", max_lr=4e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, warmup_steps_ratio=0.025,
                 scheduler_names=["constant", "constant"], blacklist_weight_modules=["norm", "embedding"])
    os.chdir(ROOT / "refcheck")  # the trainer writes tensorboard logs to ./logs
    trainer = PPOTrainer(agent=agent, batch_runner=BatchRunner(init_seed=0), rollout_buffer=RolloutBuffer(31, 16, 4),
                         optimizer_param_dict=optim, max_steps=1000, use_action_mask=not args.no_mask, device=dev,
                         mixed_precision="bfloat16", max_samples_per_epoch=300000, shuffle_on_reset=True)
    report = {"modules": origin, "envs": args.envs, "batches": args.batches, "device": str(dev), "default_agent": args.default_agent,
              "use_action_mask": not args.no_mask}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        trainer.collect_rollouts(args.envs, args.batches)
        report["collect_rollouts"] = "completed"
    except TypeError as exc:
        # ppo_trainer.py:237-239 slices a deque: the rollouts are in the buffer by then
        report["collect_rollouts"] = f"buffer filled; the reference's own statistics code then raised {exc!r}"
    torch.cuda.synchronize()
    report["collect_seconds"] = time.perf_counter() - t0
    report["buffer_size"] = int(trainer.rollout_buffer.buffer_size)
    report["episodes_tracked"] = len(trainer.episode_rewards)
    report["mean_episode_length"] = float(sum(trainer.episode_lengths) / max(1, len(trainer.episode_lengths)))
    before = [p.detach().clone() for p in agent.parameters()]
    t0 = time.perf_counter()
    metrics = trainer.update_policy(batch_size=args.minibatch, n_epochs=args.epochs)
    torch.cuda.synchronize()
    report["update_seconds"] = time.perf_counter() - t0
    report["update_metrics"] = {k: float(v) for k, v in metrics.items()}
    report["parameters_changed"] = bool(any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.parameters())))
    assert report["buffer_size"] > 0 and report["parameters_changed"] and report["episodes_tracked"] == args.envs * args.batches
    assert all(v == v for v in report["update_metrics"].values()), "NaN in the update metrics"
    if args.train_timesteps > 0:
        # the reference's train() cannot get past its first collect_rollouts (the deque slice above); the same loop from
        # this package, driving the SAME trainer object: its agent, optimizer, update_policy, writer and checkpoints
        from g2048.ppo import PPOIterationLoop

        loop = PPOIterationLoop.for_trainer(trainer)
        start = loop.total_timesteps
        t0 = time.perf_counter()
        records = loop.train(total_timesteps=args.train_timesteps, rollout_batch_size=args.envs, rollout_batches=args.batches,
                             update_epochs=args.epochs, train_batch_size=args.minibatch, save_freq=args.train_timesteps // 2)
        torch.cuda.synchronize()
        report["train_loop"] = {
            "seconds": time.perf_counter() - t0, "iterations": len(records), "timesteps": loop.total_timesteps - start,
            "trainer_total_timesteps": int(trainer.total_timesteps), "trainer_total_update_steps": int(trainer.total_update_steps),
            "checkpoints": sorted(p.name for p in Path(".").glob("*.pt")),
            "per_iteration": [{"timesteps": r["rollout"]["timesteps"], "mean_episode_length": r["rollout"]["mean_episode_length"],
                               "mean_max_episode_reward": r["rollout"]["mean_max_episode_reward"],
                               "policy_loss": float(r["update"]["policy_loss"]), "value_loss": float(r["update"]["value_loss"]),
                               "kl_divergence": float(r["update"]["kl_divergence"]), "n_updates": int(r["update"]["n_updates"])}
                              for r in records]}
        assert trainer.total_timesteps == loop.total_timesteps >= start + args.train_timesteps
        assert len(trainer.episode_rewards) == min(trainer.episode_rewards.maxlen, args.envs * args.batches * (1 + len(records)))
        assert "final_model.pt" in report["train_loop"]["checkpoints"]
    print(json.dumps(report, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
