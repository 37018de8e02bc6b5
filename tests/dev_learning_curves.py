#!/usr/bin/env python
"""(dev tool; lives under tests/ because its ``--arm port`` leg runs the oracle, which only tests / smoke / the bench baseline may do)

Learning curves of BASELINE.json's C1 configuration (CartPole-v1:ppo as shipped: 8 envs x 32 steps, 256x256 MLP, 20 passes of one
256-sample minibatch, gamma .98, lambda .8, clip .1, Adam 1e-3, 1e5 env steps) on the two arms:

  --arm engine   the CUDA engine (needs a GPU): build_agent(load_config("CartPole-v1", "ppo")).learn() without early stopping
  --arm port     the CPU port of the reference's loop (oracle/: fp64 C env step + NEXT_STEP autoreset, torch CPU policy / PPO loss with
                 autograd, numpy GAE, global-norm clip, torch Adam) -- the reference itself cannot run its collect loop here
                 (gymnasium / pytorch_lightning are not installed)

Each arm prints one JSON line per seed: the 100-episode training mean (the reference's ``train/roll/ep_rew/mean``, window 100) after every
rollout, against env steps.  ``--table a.jsonl b.jsonl`` renders the comparison (mean over seeds at fixed env-step marks, env steps to
first reach a training mean of 195 and of 400).  The two arms draw different random streams (Philox on the device; numpy / torch
generators on the CPU), so curves agree in distribution, not sample by sample.  Like the reference's README claim, this is a statement
about learning behaviour, not a parity test; the arithmetic parity tests are in tests/."""
import argparse
import json
import os
import sys
from collections import deque

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MARKS = [10_000, 20_000, 30_000, 40_000, 50_000, 60_000, 70_000, 80_000, 90_000, 99_840]          # ppo: 1e5-step budget
MARKS_REINFORCE = [20_480, 40_960, 61_440, 81_920, 102_400, 122_880, 143_360, 163_840, 184_320, 196_608]      # 2e5-step budget, 4,096 per rollout


def c1_config(seed, variant="ppo"):
    from gymnasium_solver_b200.utils.config import load_config

    cfg = load_config("CartPole-v1", variant)
    cfg.seed, cfg.seed_train, cfg.seed_val = seed, seed, 1000 + seed
    cfg.eval_freq_epochs, cfg.early_stop_on_eval_threshold, cfg.early_stop_on_train_threshold = None, False, False
    cfg.validate()
    return cfg


def run_engine(seed, variant="ppo"):
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.random import set_random_seed

    cfg = c1_config(seed, variant)
    set_random_seed(seed)
    agent = build_agent(cfg, rank=0, world_size=1)
    out = agent.learn()
    return [(int(r["train/cnt/total_env_steps"]), float(r["train/roll/ep_rew/mean"])) for r in out["history"] if "train/roll/ep_rew/mean" in r]


def run_port(seed, variant="ppo"):
    import torch

    from oracle import envs as OE
    from oracle import policy as P
    from oracle import returns as R

    cfg = c1_config(seed, variant)
    n, T, B = int(cfg.n_envs), int(cfg.n_steps), int(cfg.batch_size)
    torch.manual_seed(seed)
    torch.set_num_threads(int(os.environ.get("GS_PORT_THREADS", "1")))      # 256-sample minibatches: more threads only spin
    env = OE.OracleVecEnv("CartPole-v1", n, seed=seed)
    obs, _ = env.reset()
    params = {k: v.requires_grad_(True) for k, v in P.init_params(4, tuple(cfg.hidden_dims), 2, seed=seed).items()}
    opt = torch.optim.Adam(list(params.values()), lr=float(cfg.policy_lr))
    gen = torch.Generator().manual_seed(seed)
    ppo = variant == "ppo"
    if ppo:
        hp = dict(clip_range=float(cfg.clip_range), clip_range_vf=float(cfg.clip_range_vf), vf_coef=float(cfg.vf_coef), ent_coef=float(cfg.ent_coef),
                  normalize_adv=cfg.normalize_advantages == "batch")
    else:       # REINFORCE as shipped: Monte-Carlo reward-to-go (timeouts count as terminals), policy_targets=returns, one full-rollout batch
        assert cfg.returns_type == "mc:rtg" and cfg.policy_targets == "returns" and B == n * T and int(cfg.n_epochs) == 1
        hp = dict(ent_coef=float(cfg.ent_coef), policy_targets="returns", normalize_returns=bool(cfg.normalize_returns),
                  normalize_adv=cfg.normalize_advantages == "batch")
        baseline = R.RunningStats()
    window, curve, steps = deque(maxlen=100), [], 0
    while steps + n * T <= cfg.max_env_steps:
        bufs = dict(obs=np.zeros((T, n, 4), np.float32), act=np.zeros((T, n), np.int64), logp=np.zeros((T, n), np.float32),
                    val=np.zeros((T, n), np.float32), rew=np.zeros((T, n), np.float32), done=np.zeros((T, n), bool), to=np.zeros((T, n), bool))
        with torch.no_grad():
            for t in range(T):
                a, lp, v, _ = P.act(params, torch.from_numpy(obs), uniforms=torch.rand(n, generator=gen))
                bufs["obs"][t], bufs["act"][t], bufs["logp"][t], bufs["val"][t] = obs, a.numpy(), lp.numpy(), v.numpy()
                obs, r, term, trunc, info = env.step(a.numpy().astype(np.int32))
                bufs["rew"][t], bufs["done"][t], bufs["to"][t] = r, term | trunc, trunc
                if "_episode" in info:
                    window.extend(np.asarray(info["episode"]["r"])[np.asarray(info["_episode"])].tolist())
            _, last_v = P.forward(params, torch.from_numpy(obs))
        flat = lambda x: torch.from_numpy(np.ascontiguousarray(x.swapaxes(0, 1).reshape(n * T, *x.shape[2:])))
        remap = None
        if ppo:
            adv, ret = R.gae(bufs["val"], bufs["rew"], bufs["done"], bufs["to"], last_v.numpy(), np.zeros_like(bufs["val"]),
                             float(cfg.gamma), float(cfg.gae_lambda))
            data = [flat(bufs["obs"]), flat(bufs["act"]), flat(bufs["logp"]), flat(bufs["val"]), flat(adv), flat(ret)]
        else:   # rollout_collector.py:394-425: returns, valid mask / index map of complete episodes, running-mean baseline
            zeros = np.zeros_like(bufs["to"])
            ret = R.mc_returns(bufs["rew"], bufs["done"], zeros, float(cfg.gamma))
            valid, idx_map = R.valid_mask_and_index_map(bufs["done"], zeros)
            if valid is not None:
                baseline.update(ret.swapaxes(0, 1).reshape(-1)[valid])
                remap = torch.from_numpy(idx_map)
            adv = ret - np.float32(baseline.mean())
            data = [flat(bufs["obs"]), flat(bufs["act"]), flat(bufs["logp"]), flat(adv), flat(ret)]
        for _ in range(int(cfg.n_epochs)):
            order = torch.argsort(torch.rand(n * T, generator=gen))
            for k in range(n * T // B):
                idx = order[k * B:(k + 1) * B]
                if remap is not None:                    # slice_trajectories: samples of unfinished episodes -> nearest valid one
                    idx = remap[idx]
                opt.zero_grad()
                loss, _ = (P.ppo_loss if ppo else P.reinforce_loss)(params, *(d[idx] for d in data), **hp)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(list(params.values()), float(cfg.max_grad_norm))
                opt.step()
        steps += n * T
        if window:
            curve.append((steps, float(np.mean(window))))
    return curve


def first_reach(curve, level):
    for s, v in curve:
        if v >= level:
            return s
    return None


def at_mark(curve, mark):
    best = None
    for s, v in curve:
        if s <= mark:
            best = v
    return best


def table(paths):
    arms = {}
    for p in paths:
        for line in open(p):
            line = line.strip()
            if line.startswith("{"):
                d = json.loads(line)
                arms.setdefault(d["arm"], []).append(d)
    names = sorted(arms)
    print("| env steps | " + " | ".join(f"{a}: mean of the 100-episode training return over {len(arms[a])} seeds (min .. max)" for a in names) + " |")
    print("|---|" + "---|" * len(names))
    variant = next(iter(arms.values()))[0].get("variant", "ppo")
    for m in (MARKS if variant == "ppo" else MARKS_REINFORCE):
        cells = []
        for a in names:
            vals = [v for v in (at_mark(r["curve"], m) for r in arms[a]) if v is not None]
            cells.append(f"{np.mean(vals):.1f} ({min(vals):.0f} .. {max(vals):.0f})" if vals else "-")
        print(f"| {m:,} | " + " | ".join(cells) + " |")
    for level in (195.0, 400.0):
        cells = []
        for a in names:
            fr = [first_reach(r["curve"], level) for r in arms[a]]
            hit = [f for f in fr if f is not None]
            cells.append(f"{len(hit)}/{len(fr)} seeds" + (f", median {int(np.median(hit)):,} steps" if hit else ""))
        print(f"| first training mean >= {level:.0f} | " + " | ".join(cells) + " |")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["engine", "port"])
    ap.add_argument("--seeds", type=int, nargs="+", default=[42, 43, 44, 45, 46])
    ap.add_argument("--variant", choices=["ppo", "reinforce"], default="ppo")
    ap.add_argument("--table", nargs="+")
    args = ap.parse_args()
    if args.table:
        return table(args.table)
    for seed in args.seeds:
        curve = run_engine(seed, args.variant) if args.arm == "engine" else run_port(seed, args.variant)
        print(json.dumps({"arm": args.arm, "variant": args.variant, "seed": seed, "curve": curve}), flush=True)


if __name__ == "__main__":
    main()
