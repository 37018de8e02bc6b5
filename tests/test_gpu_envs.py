"""GPU parity: device vector envs vs the fp64 CPU restatement (oracle/envs.c) on identical Philox reset streams and
identical action sequences.  Bars (BASELINE.json north_star): observations within 1e-6, termination / truncation
flags and episode lengths bit-exact."""
import numpy as np
import pytest

from oracle import envs as OE

pytestmark = pytest.mark.gpu

ENVS = ["CartPole-v1", "Acrobot-v1", "MountainCar-v0"]


def _rollout_pair(env_id, n, steps, seed, max_episode_steps=None, wrappers=(), rng_seed=0, env_id_offset=0, resync_every=None):
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    dev_wr = []
    for spec in wrappers:
        k, p = OE.wrapper_params(spec)
        assert N.WRAPPER_KINDS[spec["id"]] == k
        dev_wr.append((k, list(p)))
    o = OE.OracleVecEnv(env_id, n, seed=seed, max_episode_steps=max_episode_steps, env_wrappers=list(wrappers), env_id_offset=env_id_offset)
    g = E.DevEnv(env_id, n, seed=seed, max_episode_steps=max_episode_steps or 0, wrappers=dev_wr, env_id_offset=env_id_offset)
    obs_o, _ = o.reset()
    obs_g = g.reset()
    np.testing.assert_array_equal(obs_g, obs_o)  # Philox + IEEE affine map: bit-exact
    rng = np.random.default_rng(rng_seed)
    nA = OE.N_ACTIONS[o.kind]
    n_done = 0
    for t in range(steps):
        if resync_every and t > 0 and t % resync_every == 0:
            # chaotic dynamics (Acrobot) amplify 1-ulp libm differences exponentially: re-inject the oracle's fp64 state so
            # every comparison is "same state, same action -> same step", which is what the parity bar states
            so, eo = o.get_state()
            g.set_state(so, eo)
        a = rng.integers(0, nA, n).astype(np.int32)
        oo, ro, to, tro, info = o.step(a)
        og, rg, tg, trg, epr, epl = g.step(a)
        np.testing.assert_allclose(og, oo, rtol=1e-6, atol=1e-6, err_msg=f"obs step {t}")
        np.testing.assert_array_equal(tg, to, err_msg=f"terminated step {t}")
        np.testing.assert_array_equal(trg, tro, err_msg=f"truncated step {t}")
        np.testing.assert_allclose(rg, ro.astype(np.float32), rtol=1e-6, atol=1e-7, err_msg=f"reward step {t}")
        done = to | tro
        if done.any():
            n_done += int(done.sum())
            np.testing.assert_array_equal(epl[done], info["episode"]["l"][done], err_msg=f"episode length step {t}")
            np.testing.assert_allclose(epr[done], info["episode"]["r"][done], rtol=1e-9, err_msg=f"episode return step {t}")
        assert (epl[~done] == 0).all()
    so, eo = o.get_state()
    sg, eg = g.get_state()
    np.testing.assert_allclose(sg, so, rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(eg, eo)
    return n_done


@pytest.mark.parametrize("env_id", ENVS)
def test_env_trajectories_match_oracle(env_id):
    steps = {"CartPole-v1": 200, "Acrobot-v1": 520, "MountainCar-v0": 420}[env_id]
    n_done = _rollout_pair(env_id, 300, steps, seed=42, resync_every=40 if env_id == "Acrobot-v1" else None)
    assert n_done > 0  # autoreset / TimeLimit path exercised (CartPole terminates, Acrobot truncates at 500, MountainCar at 200)


def test_acrobot_free_running_divergence_stays_inside_the_stated_growth_bound():
    """Acrobot WITHOUT re-injecting the oracle's state: the device's sincos / RK4 and glibc differ in the last ulp, and the chaotic
    double pendulum amplifies that.  Measured on 2,048 envs x 500 random-action steps (no resets: the time limit is lifted): the MEDIAN
    env stays at ~1e-14 for all 500 steps, the WORST env grows ~10x per 25 steps once it leaves the ulp floor: 2e-14 at step 128 (one
    rollout of the bench configurations), 4e-10 at step 256, 1e-6 at step 300.  Stated bound checked here: max |state diff| <= 1e-12 up to
    step 128, <= 1e-8 up to step 256, median <= 1e-12 up to step 500; observations (float32) within 1e-6 and flags exact up to step 256."""
    import engine_api as E

    n = 2048
    o = OE.OracleVecEnv("Acrobot-v1", n, seed=42, max_episode_steps=100000)
    g = E.DevEnv("Acrobot-v1", n, seed=42, max_episode_steps=100000)
    o.reset(); g.reset()
    rng = np.random.default_rng(0)
    for t in range(1, 501):
        a = rng.integers(0, 3, n).astype(np.int32)
        oo, ro, to, tro, info = o.step(a)
        og, rg, tg, trg, epr, epl = g.step(a)
        if t <= 256:
            np.testing.assert_allclose(og, oo, rtol=0, atol=1e-6, err_msg=f"obs step {t}")
            np.testing.assert_array_equal(tg, to, err_msg=f"terminated step {t}")
        if t in (64, 128, 192, 256, 384, 500):
            so, _ = o.get_state(); sg, _ = g.get_state()
            d = np.abs(so - sg).max(axis=0)
            if t <= 128:
                assert d.max() <= 1e-12, (t, d.max())
            elif t <= 256:
                assert d.max() <= 1e-8, (t, d.max())
            assert np.median(d) <= 1e-12, (t, np.median(d))


@pytest.mark.parametrize("env_id", ENVS)
def test_env_timelimit_truncation_and_autoreset(env_id):
    n_done = _rollout_pair(env_id, 65, 40, seed=7, max_episode_steps=6)
    assert n_done >= 65 * 5


def test_env_shards_reproduce_global_streams():
    """Rank r owning envs [r*N/W, ...) sees the same streams as the unsharded run."""
    import engine_api as E

    full = E.DevEnv("CartPole-v1", 64, seed=9)
    shard = E.DevEnv("CartPole-v1", 16, seed=9, env_id_offset=48)
    np.testing.assert_array_equal(full.reset()[48:], shard.reset())
    _rollout_pair("CartPole-v1", 16, 60, seed=9, env_id_offset=48)


def test_env_set_state_single_step_known_answers():
    import engine_api as E

    g = E.DevEnv("CartPole-v1", 2)
    g.reset()
    g.set_state(np.zeros((4, 2)), np.zeros(2, np.int32))
    obs, r, term, trunc, _, _ = g.step([1, 0])
    temp = 10.0 / 1.1
    thetaacc = (-temp) / (0.5 * (4.0 / 3.0 - 0.1 / 1.1))
    xacc = temp - 0.05 * thetaacc / 1.1
    s, el = g.get_state()
    np.testing.assert_allclose(s[:, 0], [0.0, 0.02 * xacc, 0.0, 0.02 * thetaacc], rtol=1e-15)
    np.testing.assert_allclose(s[:, 1], [0.0, -0.02 * xacc, 0.0, -0.02 * thetaacc], rtol=1e-15)
    assert (r == 1.0).all() and not term.any() and not trunc.any() and (el == 1).all()


@pytest.mark.parametrize("spec", [
    dict(id="MountainCarV0_StateCountBonus", position_bins=50, velocity_bins=50, bonus_scale=0.1, bonus_type="count"),
    dict(id="MountainCarV0_StateCountBonus", position_bins=10, velocity_bins=7, bonus_scale=1.0, bonus_type="inverse", min_count=2),
    dict(id="MountainCarV0_StateCountBonus", position_bins=20, velocity_bins=20, bonus_scale=0.5, bonus_type="log"),
    dict(id="MountainCarV0_RewardShaper", position_reward_scale=100.0, velocity_reward_scale=10.0, height_reward_scale=50.0),
])
def test_mountaincar_wrappers_match_oracle(spec):
    _rollout_pair("MountainCar-v0", 130, 260, seed=3, wrappers=[spec])


def test_cartpole_reward_shaper_matches_oracle():
    spec = dict(id="CartPoleV1_RewardShaper", angle_reward_scale=1.0, position_reward_scale=0.25, clip_potential=True)
    _rollout_pair("CartPole-v1", 130, 150, seed=11, wrappers=[spec])


def test_wrapper_errors_fail_loudly():
    import ctypes as C
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    g = E.DevEnv("CartPole-v1", 4)
    arr = (C.c_double * 5)(50, 50, 0.1, 0, 1)
    assert N.lib().gs_wrapper_attach(g.h, 1, arr, 5) != 0  # MountainCar wrapper on CartPole
    assert b"does not apply" in N.lib().gs_last_error()


@pytest.mark.parametrize("tag", ["mcar_count", "mcar_count_log", "mcar_count_inverse", "mcar_shaper", "cartpole_shaper", "cartpole_shaper_noclip"])
def test_device_reward_wrappers_match_fixtures_generated_by_the_reference_wrappers(golden_dir, tag):
    """The fused device wrappers (gs_wrapper_attach) against rewards produced by EXECUTING the reference's wrapper classes
    (tests/golden/wrappers.npz, see tests/golden/make_golden.py::golden_wrappers): same Philox reset stream, same actions, several
    episodes with autoresets; done flags exact, rewards within 1e-6."""
    import json
    import os

    import engine_api as E

    d = np.load(os.path.join(golden_dir, "wrappers.npz"))
    spec = json.loads(str(d[f"{tag}_kwargs"]))
    env_id = "CartPole-v1" if tag.startswith("cartpole") else "MountainCar-v0"
    k, p = OE.wrapper_params(spec)
    g = E.DevEnv(env_id, 1, seed=int(d[f"{tag}_seed"]), max_episode_steps=int(d[f"{tag}_max_steps"]), wrappers=[(k, list(p))])
    g.reset()
    s0, _ = g.get_state()
    np.testing.assert_array_equal(s0, d[f"{tag}_state0"])
    got, dones = [], []
    for t, a in enumerate(d[f"{tag}_actions"]):
        _, r, te, tr, _, _ = g.step(np.array([a], dtype=np.int32))
        got.append(float(r[0])); dones.append(bool(te[0] or tr[0]))
    np.testing.assert_array_equal(np.array(dones), d[f"{tag}_done"])
    np.testing.assert_allclose(np.array(got), d[f"{tag}_reward"].astype(np.float32), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["mcar_count", "mcar_shaper", "cartpole_shaper"])
def test_env_surface_with_registry_wrappers_matches_reference_fixture(golden_dir, tag):
    """The reference-facing surface end to end: build_env(env_id, env_wrappers=[YAML spec]) -> EnvWrapperRegistry.apply ->
    DeviceVecEnv.reset()/step() (Gymnasium vector protocol: 5-tuple, infos["episode"] / ["_episode"]) against the rewards the
    reference's wrapper classes produced on the same trajectory (tests/golden/wrappers.npz)."""
    import json
    import os

    import torch

    from gymnasium_solver_b200.utils.environment import build_env

    d = np.load(os.path.join(golden_dir, "wrappers.npz"))
    spec = json.loads(str(d[f"{tag}_kwargs"]))
    env_id = "CartPole-v1" if tag.startswith("cartpole") else "MountainCar-v0"
    env = build_env(env_id, n_envs=1, seed=int(d[f"{tag}_seed"]), max_episode_steps=int(d[f"{tag}_max_steps"]), env_wrappers=[spec])
    assert env.num_envs == 1 and env.wrappers and env.wrappers[0]["id"] == spec["id"]
    obs, info = env.reset()
    assert obs.shape == (1, env.single_observation_space.shape[0]) and isinstance(info, dict)
    got, n_done, ep_len = [], 0, 0
    for a in d[f"{tag}_actions"]:
        obs, r, term, trunc, infos = env.step(torch.tensor([int(a)], dtype=torch.int32, device=env.device))
        got.append(float(r[0]))
        ep_len += 1
        if bool(term[0] | trunc[0]):
            n_done += 1
            assert bool(infos["_episode"][0]) and int(infos["episode"]["l"][0]) == ep_len
            ep_len = -1                      # the autoreset step that follows is not part of any episode
    assert n_done == int(d[f"{tag}_done"].sum())
    np.testing.assert_allclose(np.array(got), d[f"{tag}_reward"].astype(np.float32), rtol=1e-6, atol=1e-6)
    with pytest.raises(KeyError):
        build_env(env_id, n_envs=1, env_wrappers=[{"id": "NoSuchWrapper"}])


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Acrobot-v1", "MountainCar-v0"])
def test_observation_space_bounds_and_recorder_protocol(golden_dir, env_id):
    """SURVEY 8(b) env protocol: ``single_observation_space`` is a float32 Box whose finite bounds agree with the reference's YAML `spec:`
    block (config/environments/<env>.yaml, in host_logic.json) and hold along a random trajectory; ``recorder()`` is a context
    manager (no-op: nothing renders on the device), ``render_mode`` is None, ``unwrapped`` is the env."""
    import json
    import os

    import torch

    from gymnasium_solver_b200.envs.device_vec_env import DeviceVecEnv

    spec = json.load(open(os.path.join(str(golden_dir), "host_logic.json")))["env_specs"][env_id]
    env = DeviceVecEnv(env_id, 256, seed=3)
    box = env.single_observation_space
    comps = spec["observation_state"]["components"]
    assert box.shape == tuple(spec["observation_state"]["shape"]) and box.dtype == np.float32 and box.low.dtype == np.float32
    assert env.observation_space.shape == (256,) + box.shape
    for d, c in enumerate(comps):
        lo, hi = float(c["range"][0]), float(c["range"][1])
        if np.isfinite(lo):                    # the YAML rounds (0.418 for 12 degrees * 2 in radians)
            assert abs(box.low[d] - lo) < 1e-3 and abs(box.high[d] - hi) < 1e-3, (d, box.low[d], box.high[d])
        assert box.low[d] < box.high[d]
    obs, _ = env.reset()
    g = torch.Generator().manual_seed(0)
    lo_t, hi_t = torch.from_numpy(box.low).cuda(), torch.from_numpy(box.high).cuda()
    for _ in range(300):
        # a terminal CartPole observation may leave the Box by one step (the env reports the state that crossed the threshold),
        # like upstream: bounds are twice the termination thresholds
        assert bool(((obs >= lo_t) & (obs <= hi_t)).all())
        obs, *_ = env.step(torch.randint(0, env.n_actions, (256,), generator=g))
    with env.recorder("unused.mp4", record_video=True) as e:
        assert e is env
    assert env.render_mode is None and env.unwrapped is env and not os.path.exists("unused.mp4")
    env.close()


def _vec_normalize_static(obs, low, high):
    """numpy restatement of VecNormalizeStatic._normalize_obs (gym_wrappers/vec_normalize_static.py:44-60), float32 like the wrapper."""
    obs = obs.astype(np.float32, copy=False)
    low, high = low.astype(np.float32), high.astype(np.float32)
    finite = np.isfinite(low) & np.isfinite(high)
    pos, zero = finite & (high > low), finite & (high == low)
    scale = np.where(pos, (high - low).astype(np.float32), 1.0).astype(np.float32)
    out = np.empty_like(obs, dtype=np.float32)
    out[..., pos] = (obs[..., pos] - low[pos]) / (scale[pos] + 1e-8)
    out[..., zero] = 0.0
    out[..., ~(pos | zero)] = obs[..., ~(pos | zero)]
    return out


@pytest.mark.parametrize("env_id", ENVS)
def test_static_observation_normalisation_is_fused_bit_exactly(env_id):
    """normalize_obs="static" (utils/environment.py:215-216 -> VecNormalizeStatic): the device env with the fused normalisation shows,
    bit for bit, what the wrapper computes from the raw env's observations -- at reset, through env.step and inside the fused collect
    kernel -- and rewrites its observation space like the wrapper (bounded dims [0, 1], unbounded dims untouched)."""
    import torch
    from gymnasium_solver_b200.utils.environment import build_env
    from gymnasium_solver_b200.utils.models import MLPActorCritic
    from gymnasium_solver_b200.utils.rollout_collector import RolloutCollector

    n, T = 300, 60
    raw = build_env(env_id, n_envs=n, seed=5, max_episode_steps=25)
    nrm = build_env(env_id, n_envs=n, seed=5, max_episode_steps=25, normalize_obs="static")
    low, high = raw.single_observation_space.low, raw.single_observation_space.high
    finite = np.isfinite(low) & np.isfinite(high)
    np.testing.assert_array_equal(nrm.single_observation_space.low[finite], 0.0)
    np.testing.assert_array_equal(nrm.single_observation_space.high[finite], 1.0)
    np.testing.assert_array_equal(nrm.single_observation_space.high[~finite], high[~finite])
    o_raw, _ = raw.reset()
    o_nrm, _ = nrm.reset()
    np.testing.assert_array_equal(o_nrm.cpu().numpy(), _vec_normalize_static(o_raw.cpu().numpy(), low, high))
    rng = np.random.default_rng(1)
    for t in range(T):
        a = rng.integers(0, raw.n_actions, n)
        o_raw, r0, te0, tr0, _ = raw.step(a)
        o_nrm, r1, te1, tr1, _ = nrm.step(a)
        np.testing.assert_array_equal(o_nrm.cpu().numpy(), _vec_normalize_static(o_raw.cpu().numpy(), low, high), err_msg=f"step {t}")
        assert torch.equal(r0, r1) and torch.equal(te0, te1) and torch.equal(tr0, tr1)      # physics, rewards and flags are untouched
    if finite.any():
        got = o_nrm.cpu().numpy()[:, finite]
        assert got.min() >= -1e-6 and got.max() <= 1 + 1e-6
    # fused collect: the buffer holds normalised observations and the policy acted on them
    env2 = build_env(env_id, n_envs=64, seed=9, max_episode_steps=25, normalize_obs="static")
    ref2 = build_env(env_id, n_envs=64, seed=9, max_episode_steps=25)
    model = MLPActorCritic(input_shape=(env2.obs_dim,), hidden_dims=(64, 64), output_shape=(env2.n_actions,), activation="relu").to(env2.device)
    col = RolloutCollector(env2, model, 16, use_gae=True)
    traj = col.collect()
    obs_buf = traj.tm["obs"].cpu().numpy()
    acts = traj.tm["actions"].cpu().numpy()
    o, _ = ref2.reset()
    for t in range(16):
        np.testing.assert_array_equal(obs_buf[t], _vec_normalize_static(o.cpu().numpy(), low, high), err_msg=f"collect step {t}")
        o, *_ = ref2.step(acts[t])


def test_rolling_observation_normalisation_is_refused():
    from gymnasium_solver_b200.utils.environment import build_env

    with pytest.raises(ValueError, match="rolling"):
        build_env("CartPole-v1", n_envs=4, seed=0, normalize_obs="rolling")
    with pytest.raises(ValueError, match="rolling"):
        build_env("CartPole-v1", n_envs=4, seed=0, normalize_obs=True)
