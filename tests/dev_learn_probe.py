import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from gymnasium_solver_b200.agents import build_agent
from gymnasium_solver_b200.utils.random import set_random_seed
from gymnasium_solver_b200.utils.config import load_config
for rep in range(3):
    cfg = load_config("CartPole-v1", "ppo"); cfg.validate()
    set_random_seed(cfg.seed)
    agent = build_agent(cfg, rank=0, world_size=1)
    out = agent.learn()
    hist = out["history"]
    tc = [r["train/roll/ep_rew/mean"] for r in hist if "train/roll/ep_rew/mean" in r]
    print(rep, "train peak", max(tc), "best eval", out["best_eval_reward"], "steps", out["total_env_steps"], flush=True)
