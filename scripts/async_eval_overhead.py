#!/usr/bin/env python
"""What an evaluation every epoch costs the training loop at the headline configuration (CartPole-v1:ppo_b200: 65,536 envs x 128 steps,
64x64 MLP, 10 passes x 8 minibatches), three ways: no evaluation, synchronous evaluation (the val collector runs on the training stream
between two epochs), asynchronous evaluation (``eval_async``: weight snapshot + background thread + its own CUDA stream).  One
evaluation = 65,536 episodes (one per val env).  Wall-clock per epoch over ``--epochs`` epochs after two warm-up epochs, the final
background evaluation included (learn() joins it).  Prints a markdown table."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gymnasium_solver_b200.agents import build_agent
from gymnasium_solver_b200.utils.config import load_config


def run(mode: str, epochs: int):
    cfg = load_config("CartPole-v1", "ppo_b200")
    cfg.early_stop_on_eval_threshold = cfg.early_stop_on_train_threshold = False
    cfg.eval_freq_epochs = None if mode == "none" else 1
    cfg.eval_async = mode == "async"
    cfg.validate()
    agent = build_agent(cfg, rank=0, world_size=1)
    agent.learn(max_epochs=2)                       # warm-up: allocations, first evaluation
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = agent.learn(max_epochs=2 + epochs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    evals = [r for r in out["history"] if "val/roll/ep_rew/mean" in r]
    last = agent.wait_async_eval() if mode == "async" else (evals[-1] if evals else {})
    n_eval = len({r.get("val/eval/model_epoch", r["epoch"]) for r in evals}) if mode != "none" else 0
    steps = int(cfg.n_envs) * int(cfg.n_steps) * epochs
    return dict(mode=mode, ms=dt / epochs * 1e3, rate=steps / dt, n_eval=n_eval,
                eval_mean=last.get("val/roll/ep_rew/mean", last.get("roll/ep_rew/mean")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10)
    args = ap.parse_args()
    rows = [run(m, args.epochs) for m in ("none", "sync", "async")]
    base = rows[0]["ms"]
    print("| evaluation every epoch | ms per epoch (wall clock) | training env-steps/s | added per epoch | distinct evaluations reported | last eval mean return |")
    print("|---|---|---|---|---|---|")
    for r in rows:
        em = "-" if r["eval_mean"] is None else f"{r['eval_mean']:.1f}"
        print(f"| {r['mode']} | {r['ms']:.2f} | {r['rate'] / 1e6:.1f} M | {r['ms'] - base:+.2f} ms | {r['n_eval']} | {em} |", flush=True)


if __name__ == "__main__":
    main()
