#!/usr/bin/env python
"""``python train.py <env>:<variant> [--override KEY=VALUE]... [--max-env-steps N]`` (reference: train.py:30-148).

Multi-GPU: ``python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 train.py <env>:<variant>``; each
rank owns n_envs/N environments; the gradient mean of every minibatch runs over NVLink peer memory inside the step's tail
kernel (``grad_allreduce=nccl`` selects the torch.distributed path).  ``--checkpoint-dir`` / ``--resume`` save and continue a run
bit for bit (weights, Adam state, device env snapshot per rank)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch


def _parse_value(v: str):
    low = v.lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("none", "null"):
        return None
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    try:
        return json.loads(v)
    except ValueError:
        return v


def parse_overrides(items) -> dict:
    """``KEY=VALUE`` strings -> dict (reference utils/train_launcher.py:22-53: bools, ints, floats, else strings; here also
    ``none`` / scientific notation / JSON lists).  ``ValueError`` for an item without ``=``."""
    out = {}
    for item in items or ():
        if "=" not in item:
            raise ValueError(f"Invalid override format: {item}. Expected KEY=VALUE")
        k, v = item.split("=", 1)
        out[k.strip()] = _parse_value(v.strip())
    return out


def apply_overrides(config, overrides: dict):
    """reference utils/train_launcher.py:81-98: ``setattr`` after validation, ``ValueError`` for a key that is not a Config field."""
    from dataclasses import fields

    valid = {f.name for f in fields(config)}
    for k, v in overrides.items():
        if k not in valid:
            raise ValueError(f"Invalid config field: {k}. Not a valid Config attribute.")
        setattr(config, k, v)
    return config


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Train an agent on the b200 engine.")
    ap.add_argument("config_id", nargs="?", default="CartPole-v1:ppo", help="<env>:<variant>, e.g. CartPole-v1:ppo")
    ap.add_argument("--config_id", dest="config_id_opt", default=None)
    ap.add_argument("--override", action="append", default=[], metavar="KEY=VALUE")
    ap.add_argument("--max-env-steps", type=int, default=None)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--resume", default=None, help="checkpoint directory to resume from")
    ap.add_argument("--run-dir", default=None, help="write config.json here and re-read it at every epoch start: edit policy_lr / clip_range / "
                                                     "ent_coef / vf_coef / n_epochs while training runs (reference agents/hyperparameter_mixin.py:37-64)")
    args = ap.parse_args(argv)
    spec = args.config_id_opt or args.config_id
    if ":" not in spec:
        ap.error("config must be <env>:<variant>")
    env_id, variant = spec.split(":", 1)

    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config
    from gymnasium_solver_b200.utils.random import set_random_seed

    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    config = load_config(env_id, variant)
    apply_overrides(config, parse_overrides(args.override))
    if args.max_env_steps is not None:
        config.max_env_steps = args.max_env_steps
    set_random_seed(config.seed)
    agent = build_agent(config)
    if args.resume:
        agent.load_checkpoint(args.resume)
    if args.run_dir:                                # rank 0 reads the file at every epoch start and broadcasts its reading
        from gymnasium_solver_b200.agents.hyperparameter_mixin import RunConfigFile
        agent.run = RunConfigFile(args.run_dir)
        if int(os.environ.get("RANK", 0)) == 0 or not os.path.exists(agent.run.path):
            agent.run.save_config(config)

    def log(row):
        if config.quiet:
            return
        keys = ("epoch", "train/cnt/total_env_steps", "train/roll/ep_rew/mean", "train/roll/fps", "train/opt/loss/total",
                "train/opt/ppo/approx_kl", "val/roll/ep_rew/mean")
        print("  ".join(f"{k.split('/')[-1]}={row[k]:.4g}" if isinstance(row.get(k), float) else f"{k.split('/')[-1]}={row.get(k)}"
                        for k in keys if k in row), flush=True)

    result = agent.learn(log_fn=log)
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"Stopped: {result['stop_reason']}  epochs={result['epochs']}  env_steps={result['total_env_steps']}  "
              f"elapsed={result['elapsed_s']:.1f}s  best_eval={result['best_eval_reward']}")
    if args.checkpoint_dir:                       # every rank: rank 0 writes the shared files, each rank its own env shard
        agent.save_checkpoint(args.checkpoint_dir)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
