"""BASELINE.json's larger configurations at FULL size on one GPU, checked through size-independent properties (an oracle run
at these sizes would take minutes to hours on the CPU):

  C3  CartPole-v1:reinforce, 262,144 envs x 128 steps, Monte-Carlo reward-to-go + running-mean baseline
  C4  Acrobot-v1:ppo, 1,048,576 envs (128x128 MLP: the FMA-pipe update kernel) and MountainCar-v0:ppo with the per-env
      50x50 StateCountBonus tables (10 GB of counters)

Each runs one full training iteration (collect -> targets -> every minibatch of one pass -> optimizer steps)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agent(env, variant, **over):
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config

    cfg = load_config(env, variant)
    for k, v in over.items():
        setattr(cfg, k, v)
    cfg.validate()
    return build_agent(cfg, rank=0, world_size=1), cfg


def _need_memory(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * 2 ** 30:
        pytest.skip(f"needs {gb} GB of free device memory")


def _finite_metrics(agent):
    m = agent.pop_epoch_metrics()
    bad = [k for k, v in m.items() if isinstance(v, float) and not math.isfinite(v)]
    assert not bad, bad
    return m


def test_c3_reinforce_262144_envs_full_iteration():
    _need_memory(20)
    agent, cfg = _agent("CartPole-v1", "reinforce_b200")
    T, n = int(cfg.n_steps), int(cfg.n_envs)
    assert (T, n) == (128, 262144)
    w0 = agent.policy_model.flat_params.clone()
    traj = agent.train_one_rollout()
    tm = traj.tm
    r, d = tm["rewards"], tm["dones"].bool()
    # CartPole: reward 1 on every real step, 0 on the autoreset step that follows a done (NEXT_STEP mode), never two dones in a row
    assert bool(((r == 1) | (r == 0)).all())
    assert bool((r[1:][d[:-1]] == 0).all()) and not bool(d[1:][d[:-1]].any())
    assert int(d.sum()) > n            # an untrained policy drops the pole within ~20 steps: more than one episode per env
    # reward-to-go recurrence, bit for bit: R[t] = r[t] + gamma * R[t+1] inside an episode, R[t] = r[t] at its end (timeouts count
    # as terminals in MC mode); the last row continues into nothing
    R = tm["ret"] if (str(cfg.returns_type).endswith("rtg") and str(cfg.normalize_returns) in ("off", "False", "None")) else None
    assert R is not None, (cfg.returns_type, cfg.normalize_returns)
    if R is not None:
        g = torch.tensor(float(cfg.gamma), dtype=torch.float32, device=R.device)
        nxt = torch.where(d[:-1], torch.zeros_like(R[1:]), R[1:])
        assert torch.equal(R[:-1], r[:-1] + g * nxt)
        assert torch.equal(R[-1], r[-1])
    m = _finite_metrics(agent)
    assert m["opt/policy/entropy"] > 0.5           # still close to log 2 after one update
    assert not torch.equal(agent.policy_model.flat_params, w0)
    cm = agent.get_rollout_collector("train").get_metrics()
    assert cm["cnt/total_env_steps"] == T * n and cm["roll/episodes"] == int(d.sum())
    assert 9 < cm["roll/ep_len/mean"] < 60


def test_c4_acrobot_1m_envs_full_iteration():
    _need_memory(40)
    agent, cfg = _agent("Acrobot-v1", "ppo_b200", n_epochs=1)
    T, n = int(cfg.n_steps), int(cfg.n_envs)
    assert (T, n) == (128, 1048576)
    traj = agent.train_one_rollout()
    tm = traj.tm
    obs = tm["obs"]
    # observation manifold: (cos, sin) pairs on the unit circle, velocities inside the clamp box
    c1, s1, c2, s2, v1, v2 = (obs[..., i] for i in range(6))
    assert float(((c1 * c1 + s1 * s1) - 1).abs().max()) < 1e-6 and float(((c2 * c2 + s2 * s2) - 1).abs().max()) < 1e-6
    assert float(v1.abs().max()) <= 4 * math.pi + 1e-5 and float(v2.abs().max()) <= 9 * math.pi + 1e-5
    r, d = tm["rewards"], tm["dones"].bool()
    assert bool(((r == -1) | (r == 0)).all())
    assert not bool(d[1:][d[:-1]].any())
    # GAE identities that hold whatever the data: ret = adv + values exactly; where a step is terminal adv = r - v
    assert torch.equal(tm["ret"], tm["adv"] + tm["values"])
    term = d & ~tm["timeouts"].bool()
    if bool(term.any()):
        assert torch.equal(tm["adv"][term], (r - tm["values"])[term])
    m = _finite_metrics(agent)
    assert 0.9 < m["opt/policy/entropy"] <= math.log(3) + 1e-4
    cm = agent.get_rollout_collector("train").get_metrics()
    assert cm["cnt/total_env_steps"] == T * n
    np.testing.assert_allclose(cm["roll/reward/mean"], float(r.double().mean()), rtol=1e-6)


def test_c4_mountaincar_1m_envs_state_count_bonus_full_iteration():
    _need_memory(60)
    agent, cfg = _agent("MountainCar-v0", "ppo_b200", n_epochs=1)
    T, n = int(cfg.n_steps), int(cfg.n_envs)
    assert (T, n) == (128, 1048576)
    traj = agent.train_one_rollout()
    tm = traj.tm
    obs, r, d = tm["obs"], tm["rewards"], tm["dones"].bool()
    assert float(obs[..., 0].min()) >= -1.2 - 1e-6 and float(obs[..., 0].max()) <= 0.6 + 1e-6
    assert float(obs[..., 1].abs().max()) <= 0.07 + 1e-7
    # reward = -1 + 0.1 / sqrt(max(count, 1)) with count >= 0 before the increment: inside (-1, -0.9]; the autoreset step pays 0
    real = torch.ones_like(d)
    real[1:] = ~d[:-1]
    assert bool((r[real] > -1.0).all()) and bool((r[real] <= -0.9 + 1e-6).all())
    assert bool((r[~real] == 0).all())
    # 128 steps cannot finish a 200-step MountainCar episode unless the goal is reached, which a random policy does not do
    assert int(d.sum()) == 0
    # first visit of the first step pays the full bonus: -0.9 (count 0 -> max(count, 1) = 1)
    np.testing.assert_allclose(r[0].cpu().numpy(), -0.9, rtol=0, atol=1e-6)
    m = _finite_metrics(agent)
    assert 0.9 < m["opt/policy/entropy"] <= math.log(3) + 1e-4
