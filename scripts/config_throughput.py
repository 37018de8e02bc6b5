#!/usr/bin/env python
"""Throughput of BASELINE.json's other configurations on ONE GPU at full size: C3 (CartPole-v1 REINFORCE, 262,144 envs),
C4 (Acrobot-v1 PPO 128x128 and MountainCar-v0 PPO 256x256 + StateCountBonus, 1,048,576 envs), each as whole training
iterations (collect + targets + every pass) timed with CUDA events after one warm-up iteration.  Prints a markdown table.
These are parity-test configurations (tests/test_gpu_fullsize.py), not bench lines; the numbers document where they stand."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gymnasium_solver_b200.agents import build_agent
from gymnasium_solver_b200.utils.config import load_config

CASES = [("C3 CartPole-v1:reinforce_b200", "CartPole-v1", "reinforce_b200", 3),
         ("C4 Acrobot-v1:ppo_b200", "Acrobot-v1", "ppo_b200", 1),
         ("C4 MountainCar-v0:ppo_b200", "MountainCar-v0", "ppo_b200", 1)]
print("| config | envs x steps | network | passes x minibatches | ms per iteration | env-steps/s | update path |\n|---|---|---|---|---|---|---|")
only = os.environ.get("GS_CASES")
for name, env, variant, iters in CASES:
    if only and not any(o in name for o in only.split(",")):
        continue
    cfg = load_config(env, variant)
    cfg.eval_freq_epochs = None
    cfg.validate()
    agent = build_agent(cfg, rank=0, world_size=1)
    agent.train_one_rollout()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        agent.train_one_rollout()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    steps = int(cfg.n_envs) * int(cfg.n_steps)
    n_mb = steps // int(cfg.batch_size)
    hd = tuple(cfg.hidden_dims)
    path = "tcgen05 (fp16x3)" if hd in ((64, 64), (128, 128), (256, 256)) else "fp32 FMA pipe"
    print(f"| {name} | {int(cfg.n_envs):,} x {int(cfg.n_steps)} | {hd} | {int(agent.n_epochs)} x {n_mb} | {ms:.1f} | {steps / ms * 1e3 / 1e6:.1f} M | {path} |", flush=True)
    del agent
    torch.cuda.empty_cache()
