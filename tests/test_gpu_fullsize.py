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


# ---- exact replays of env SUBSETS of the full-size runs ---------------------------------------------------------------------------------
# Reset noise and action draws are Philox streams keyed by (seed, GLOBAL env id, reset count / vector step), so the CPU oracle can replay
# any handful of the 65,536 .. 1,048,576 envs of a full-size rollout exactly: same initial states, same action uniforms, same policy.
# Four blocks of 64 global ids (first, two interior, last) x all 128 steps per configuration.
TAG_ACTION = 0xAC700000


def _action_uniforms(seed, gid0, n, step0, T):
    from oracle import envs as OE
    u = np.zeros((T, n), np.float32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for t in range(T):
        step = step0 + t
        for i in range(n):
            gid = gid0 + i
            r = OE.philox4x32_10([gid & 0xFFFFFFFF, gid >> 32, step & 0xFFFFFFFF, TAG_ACTION | ((step >> 32) & 0xFFFFF)], key)
            u[t, i] = np.float32(int(r[0]) >> 8) * np.float32(1.0 / 16777216.0)
    return u


def _replay_subsets(agent, cfg, traj, params_before, *, mc: bool):
    """Oracle replay of 4 x 64 envs of the rollout `traj` that `agent` collected with the weights `params_before`."""
    from oracle import envs as OE
    from oracle import policy as P
    from oracle import returns as R

    tm = traj.tm
    T, n = int(cfg.n_steps), int(cfg.n_envs)
    env = agent.get_env("train")
    col = agent.get_rollout_collector("train")
    wrappers = [dict(w) for w in (getattr(cfg, "env_wrappers", None) or ())]
    p = {k: v.detach().cpu().clone() for k, v in params_before.items()}
    act = getattr(cfg, "activation", "relu")
    blocks = [0, (n // 3) & ~63, (2 * n // 3 + 64) & ~63, n - 64]
    compared = flips = 0
    for g0 in blocks:
        sl = slice(g0, g0 + 64)
        oenv = OE.OracleVecEnv(cfg.env_id, 64, seed=env.seed, env_id_offset=env.env_id_offset + g0, max_episode_steps=env.max_episode_steps,
                               env_wrappers=wrappers)
        obs, _ = oenv.reset()
        U = _action_uniforms(col.rng_seed, env.env_id_offset + g0, 64, 0, T)
        alive = np.ones(64, bool)                           # envs whose action sequence still agrees (a draw within 1e-5 of a CDF edge may flip)
        dev = {k: tm[k][:, sl].cpu().numpy() for k in ("obs", "actions", "logprobs", "values", "rewards", "dones", "timeouts")}
        ep_len = col._buffer.ep_length_buf[:T, sl].cpu().numpy()
        for t in range(T):
            a, lp, v, _ = P.act(p, torch.from_numpy(obs), deterministic=False, uniforms=torch.from_numpy(U[t]), activation=act)
            a = a.numpy().astype(np.int32)
            alive &= dev["actions"][t] == a
            np.testing.assert_allclose(dev["obs"][t][alive], obs[alive], atol=1e-6, rtol=0, err_msg=f"obs, block {g0}, step {t}")
            np.testing.assert_allclose(dev["logprobs"][t][alive], lp.numpy()[alive], rtol=1e-5, atol=2e-6)
            np.testing.assert_allclose(dev["values"][t][alive], v.numpy()[alive], rtol=1e-5, atol=2e-6)
            obs, r, term, trunc, info = oenv.step(np.where(alive, a, dev["actions"][t]))
            np.testing.assert_allclose(dev["rewards"][t][alive], r.astype(np.float32)[alive], rtol=1e-6, atol=1e-7)
            np.testing.assert_array_equal(dev["dones"][t].astype(bool)[alive], (term | trunc)[alive])            # flags: exact
            np.testing.assert_array_equal(dev["timeouts"][t].astype(bool)[alive], trunc[alive])
            done = (term | trunc) & alive
            if done.any():
                np.testing.assert_array_equal(ep_len[t][done], info["episode"]["l"][done])                      # episode lengths: exact
        compared += int(alive.sum()); flips += int((~alive).sum())
        # targets of those columns, bit for bit, from the device's own rollout columns (the scans are per env)
        d, to = dev["dones"].astype(bool), dev["timeouts"].astype(bool)
        if mc:
            ret = R.mc_returns(dev["rewards"], d, to if not col.mc_treat_timeouts_as_terminals else None, float(cfg.gamma))
            if str(cfg.returns_type) == "mc:episode":
                ret = R.to_full_episode(ret, d, to if not col.mc_treat_timeouts_as_terminals else None)
            np.testing.assert_array_equal(col._ret[:T, sl].cpu().numpy() if not col.normalize_returns else ret, ret)
        else:
            lv = col._last_values[sl].cpu().numpy()
            adv, ret = R.gae(dev["values"], dev["rewards"], d, to, lv, np.zeros_like(dev["values"]), float(cfg.gamma), float(cfg.gae_lambda))
            if not col.normalize_advantages:
                np.testing.assert_array_equal(col._adv[:T, sl].cpu().numpy(), adv)
            np.testing.assert_array_equal(col._ret[:T, sl].cpu().numpy(), ret)
    assert flips <= 0.05 * (compared + flips), (compared, flips)
    return compared


def _params_of(agent):
    m = agent.policy_model
    sd = m.state_dict()
    out = dict(w1=sd["backbone.0.weight"], b1=sd["backbone.0.bias"], w2=sd["backbone.2.weight"], b2=sd["backbone.2.bias"],
               wp=sd["policy_head.weight"], bp=sd["policy_head.bias"])
    if "value_head.weight" in sd:
        out.update(wv=sd["value_head.weight"], bv=sd["value_head.bias"])
    return {k: v.detach().clone() for k, v in out.items()}


@pytest.mark.parametrize("env_id,variant,mem,mc", [("CartPole-v1", "ppo_b200", 8, False), ("CartPole-v1", "reinforce_b200", 20, True),
                                                   ("Acrobot-v1", "ppo_b200", 40, False), ("MountainCar-v0", "ppo_b200", 60, False)])
def test_full_size_rollout_subset_replays_exactly_on_the_cpu_oracle(env_id, variant, mem, mc):
    """C2 / C3 / both C4 configurations at FULL size: 256 of the rollout's envs (4 blocks of 64 global ids) replayed step by step by the
    CPU oracle -- observations 1e-6, rewards 1e-6, done / timeout flags and episode lengths exact, GAE / MC targets of those columns bit for bit."""
    _need_memory(mem)
    agent, cfg = _agent(env_id, variant, n_epochs=1)
    assert int(cfg.n_steps) == 128 and int(cfg.n_envs) >= 65536
    before = _params_of(agent)
    traj = agent.get_rollout_collector("train").collect()
    n_ok = _replay_subsets(agent, cfg, traj, before, mc=mc)
    assert n_ok >= 0.9 * 256
