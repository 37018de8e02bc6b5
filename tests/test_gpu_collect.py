"""GPU parity: the fused persistent collect kernel (gs_rollout_collect) vs (a) the unfused device sequence
gs_policy_act + gs_env_step, which must agree bit for bit, and (b) the CPU oracle loop (oracle env + oracle policy with
the same counter-based action uniforms)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import envs as OE
from oracle import policy as P

pytestmark = pytest.mark.gpu

TAG_ACTION = 0xAC700000


def _action_uniforms(seed, gid0, n, step0, T):
    u = np.zeros((T, n), np.float32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for t in range(T):
        step = step0 + t
        for i in range(n):
            gid = gid0 + i
            r = OE.philox4x32_10([gid & 0xFFFFFFFF, gid >> 32, step & 0xFFFFFFFF, TAG_ACTION | ((step >> 32) & 0xFFFFF)], key)
            u[t, i] = np.float32(int(r[0]) >> 8) * np.float32(1.0 / 16777216.0)
    return u


def _alloc_rollout(N_, T, n, D, with_next=True):
    from gymnasium_solver_b200 import _native as N

    f = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")
    t = dict(obs=f(T, n, D), next_obs=f(T, n, D) if with_next else None, actions=f(T, n, dt=torch.int32), logprobs=f(T, n),
             values=f(T, n), rewards=f(T, n), dones=f(T, n, dt=torch.uint8), timeouts=f(T, n, dt=torch.uint8), last_obs=f(n, D),
             last_values=f(n), ep_return=f(T, n, dt=torch.float64), ep_length=f(T, n, dt=torch.int32))
    r = N.GsRollout()
    r.T, r.obs_dim, r.N = T, D, n
    for k, v in t.items():
        setattr(r, k, N.ptr(v))
    return r, t


def _collect(env, p_dev, T, cur_obs, seed, step0, deterministic=False, with_next=True):
    from gymnasium_solver_b200 import _native as N

    m = N.mlp_struct_from_params(p_dev)
    r, t = _alloc_rollout(N, T, env.n, env.D, with_next)
    N.check(N.lib().gs_rollout_collect(env.h, C.byref(m), C.byref(r), N.ptr(cur_obs), seed, step0, int(deterministic), N.stream()))
    torch.cuda.synchronize()
    return {k: (None if v is None else v.cpu().numpy()) for k, v in t.items()}


@pytest.mark.parametrize("env_id,hidden", [("CartPole-v1", (64, 64)), ("Acrobot-v1", (128, 128)), ("MountainCar-v0", (256, 256)),
                                           ("CartPole-v1", (64,))])
def test_fused_collect_equals_unfused_device_sequence(env_id, hidden):
    import engine_api as E

    n, T, seed = 200, 24, 11
    kind = OE.ENV_KINDS[env_id]
    D, A = OE.OBS_DIM[kind], OE.N_ACTIONS[kind]
    p_dev = E.dev_params(P.random_params(D, hidden, A, seed=2))
    fused_env = E.DevEnv(env_id, n, seed=seed, max_episode_steps=9)
    step_env = E.DevEnv(env_id, n, seed=seed, max_episode_steps=9)
    cur = E.cu(fused_env.reset())
    obs = step_env.reset()
    out = _collect(fused_env, p_dev, T, cur, seed=77, step0=1000)
    for t in range(T):
        a, lp, v, _ = E.policy_act(p_dev, obs, seed=77, offset=1000 + t)
        np.testing.assert_array_equal(out["obs"][t], obs)
        np.testing.assert_array_equal(out["actions"][t], a)
        np.testing.assert_array_equal(out["logprobs"][t], lp)
        np.testing.assert_array_equal(out["values"][t], v)
        obs, r, term, trunc, epr, epl = step_env.step(a)
        np.testing.assert_array_equal(out["next_obs"][t], obs)
        np.testing.assert_array_equal(out["rewards"][t], r)
        np.testing.assert_array_equal(out["dones"][t].astype(bool), term | trunc)
        np.testing.assert_array_equal(out["timeouts"][t].astype(bool), trunc)
        np.testing.assert_array_equal(out["ep_length"][t], epl)
        np.testing.assert_array_equal(out["ep_return"][t], epr)
    np.testing.assert_array_equal(out["last_obs"], obs)
    np.testing.assert_array_equal(cur.cpu().numpy(), obs)
    np.testing.assert_array_equal(out["last_values"], E.policy_values(p_dev, obs))
    sf, ef = fused_env.get_state()
    ss, es = step_env.get_state()
    np.testing.assert_array_equal(sf, ss)
    np.testing.assert_array_equal(ef, es)
    assert out["dones"].sum() > 0


def test_two_collects_continue_one_trajectory():
    import engine_api as E

    n, T = 130, 10
    p_dev = E.dev_params(P.random_params(4, (64, 64), 2, seed=4))
    e1 = E.DevEnv("CartPole-v1", n, seed=5)
    e2 = E.DevEnv("CartPole-v1", n, seed=5)
    c1, c2 = E.cu(e1.reset()), E.cu(e2.reset())
    a = _collect(e1, p_dev, T, c1, 9, 0, with_next=False)
    b = _collect(e1, p_dev, T, c1, 9, T, with_next=False)
    full = _collect(e2, p_dev, 2 * T, c2, 9, 0, with_next=False)
    for k in ("obs", "actions", "logprobs", "values", "rewards", "dones", "timeouts", "ep_length"):
        np.testing.assert_array_equal(np.concatenate([a[k], b[k]]), full[k], err_msg=k)
    np.testing.assert_array_equal(b["last_values"], full["last_values"])


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Acrobot-v1", "MountainCar-v0"])
@pytest.mark.parametrize("deterministic", [False, True])
def test_fused_collect_matches_cpu_oracle_rollout(env_id, deterministic):
    import engine_api as E

    n, T, seed, rng_seed, step0 = 96, 40, 21, 555, 12345
    kind = OE.ENV_KINDS[env_id]
    D, A = OE.OBS_DIM[kind], OE.N_ACTIONS[kind]
    p = P.random_params(D, (64, 64), A, seed=8, scale=1.5)
    dev_env = E.DevEnv(env_id, n, seed=seed, max_episode_steps=15)
    cur = E.cu(dev_env.reset())
    out = _collect(dev_env, E.dev_params(p), T, cur, rng_seed, step0, deterministic=deterministic)

    oenv = OE.OracleVecEnv(env_id, n, seed=seed, max_episode_steps=15)
    obs, _ = oenv.reset()
    U = _action_uniforms(rng_seed, 0, n, step0, T)
    alive = np.ones(n, bool)  # envs whose action sequence still agrees (a draw within 1e-5 of a CDF edge may flip)
    for t in range(T):
        a, lp, v, logits = P.act(p, torch.from_numpy(obs), deterministic=deterministic, uniforms=torch.from_numpy(U[t]))
        a = a.numpy().astype(np.int32)
        alive &= out["actions"][t] == a
        np.testing.assert_allclose(out["obs"][t][alive], obs[alive], atol=1e-6, rtol=0)
        np.testing.assert_allclose(out["logprobs"][t][alive], lp.numpy()[alive], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(out["values"][t][alive], v.numpy()[alive], rtol=1e-5, atol=2e-6)
        obs, r, term, trunc, info = oenv.step(np.where(alive, a, out["actions"][t]))
        np.testing.assert_allclose(out["rewards"][t][alive], r.astype(np.float32)[alive], rtol=1e-6)
        np.testing.assert_array_equal(out["dones"][t].astype(bool)[alive], (term | trunc)[alive])
        np.testing.assert_array_equal(out["timeouts"][t].astype(bool)[alive], trunc[alive])
        done = (term | trunc) & alive
        if done.any():
            np.testing.assert_array_equal(out["ep_length"][t][done], info["episode"]["l"][done])
    assert alive.mean() > 0.97
    assert out["dones"].sum() > n  # several episodes per env


def test_collect_errors_fail_loudly():
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    env = E.DevEnv("CartPole-v1", 8)
    cur = E.cu(env.reset())
    p_dev = E.dev_params(P.random_params(4, (64, 64), 3))  # 3 actions on CartPole
    with pytest.raises(N.EngineError, match="actions"):
        _collect(env, p_dev, 4, cur, 0, 0)
