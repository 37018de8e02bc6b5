"""Checks of oracle/envs.c.  Physics parity with gymnasium 1.1.1 is UNPINNED (package absent, no reference golden
trajectories); these tests pin the generator (Random123 known answers), hand-derived single steps, and the
TimeLimit / NEXT_STEP autoreset / RecordEpisodeStatistics bookkeeping the reference's collector depends on."""
import math
import os

import numpy as np
import pytest

from oracle import envs as E


def test_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(x) for x in E.philox4x32_10([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in E.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in E.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_reset_distributions_and_streams():
    n = 4096
    for env_id, lo, hi in (("CartPole-v1", -0.05, 0.05), ("Acrobot-v1", -0.1, 0.1)):
        e = E.OracleVecEnv(env_id, n, seed=3)
        e.reset()
        s, el = e.get_state()
        assert s.min() >= lo and s.max() < hi and abs(s.mean()) < 0.01 * (hi - lo) * 10
        assert (el == 0).all()
    e = E.OracleVecEnv("MountainCar-v0", n, seed=3)
    e.reset()
    s, _ = e.get_state()
    assert s[0].min() >= -0.6 and s[0].max() < -0.4 and (s[1] == 0).all()
    # acrobot reset state is float32-valued (uniform(...).astype(np.float32))
    e = E.OracleVecEnv("Acrobot-v1", 8, seed=1)
    e.reset()
    s, _ = e.get_state()
    np.testing.assert_array_equal(s, s.astype(np.float32).astype(np.float64))
    # global env ids make shards reproduce the unsharded stream
    full = E.OracleVecEnv("CartPole-v1", 8, seed=9)
    shard = E.OracleVecEnv("CartPole-v1", 4, seed=9, env_id_offset=4)
    o_full, _ = full.reset()
    o_shard, _ = shard.reset()
    np.testing.assert_array_equal(o_full[4:], o_shard)


def test_cartpole_single_step_hand_derived():
    e = E.OracleVecEnv("CartPole-v1", 1)
    e.reset()
    e.set_state(np.zeros((4, 1)), np.zeros(1, np.int32))
    obs, r, term, trunc, _ = e.step(np.array([1]))
    # theta = 0: temp = 10/1.1, thetaacc = -temp / (0.5*(4/3 - 0.1/1.1)), xacc = temp - 0.05*thetaacc/1.1
    temp = 10.0 / 1.1
    thetaacc = (-temp) / (0.5 * (4.0 / 3.0 - 0.1 / 1.1))
    xacc = temp - 0.05 * thetaacc / 1.1
    s, el = e.get_state()
    np.testing.assert_allclose(s[:, 0], [0.0, 0.02 * xacc, 0.0, 0.02 * thetaacc], rtol=1e-15)
    assert r[0] == 1.0 and not term[0] and not trunc[0] and el[0] == 1
    np.testing.assert_array_equal(obs[0], s[:, 0].astype(np.float32))


def test_cartpole_termination_thresholds_and_reward_on_terminal_step():
    e = E.OracleVecEnv("CartPole-v1", 2)
    e.reset()
    th = 12 * 2 * math.pi / 360
    st = np.zeros((4, 2))
    st[0, 0] = 2.4  # x at threshold with positive velocity crosses it
    st[1, 0] = 1.0
    st[2, 1] = th - 1e-9  # theta just below threshold, falling outwards
    st[3, 1] = 1.0
    e.set_state(st, np.zeros(2, np.int32))
    _, r, term, trunc, info = e.step(np.array([1, 1]))
    assert term.all() and not trunc.any() and (r == 1.0).all()
    np.testing.assert_array_equal(info["episode"]["l"], [1, 1])
    np.testing.assert_array_equal(info["episode"]["r"], [1.0, 1.0])


def test_next_step_autoreset_and_episode_statistics():
    e = E.OracleVecEnv("CartPole-v1", 1, seed=5, max_episode_steps=3)
    obs0, _ = e.reset()
    lens, flags = [], []
    for t in range(9):
        obs, r, term, trunc, info = e.step(np.array([t % 2]))
        flags.append((bool(term[0]), bool(trunc[0]), float(r[0])))
        if "episode" in info:
            lens.append((t, int(info["episode"]["l"][0]), float(info["episode"]["r"][0])))
    # truncation at elapsed == 3, then ONE dummy reset step (reward 0, flags False), then a fresh 3-step episode
    assert flags[2] == (False, True, 1.0) and flags[3] == (False, False, 0.0)
    assert flags[6] == (False, True, 1.0) and flags[7] == (False, False, 0.0)
    assert lens == [(2, 3, 3.0), (6, 3, 3.0)]
    # the reset-step observation is the new episode's initial state (small uniform noise)
    e2 = E.OracleVecEnv("CartPole-v1", 1, seed=5, max_episode_steps=3)
    e2.reset()
    for t in range(4):
        obs, *_ = e2.step(np.array([t % 2]))
    assert np.abs(obs).max() < 0.05
    s, el = e2.get_state()
    assert el[0] == 0


def test_mountaincar_step_and_bounds():
    e = E.OracleVecEnv("MountainCar-v0", 3)
    e.reset()
    st = np.array([[-0.5, -1.2, 0.49], [0.0, -0.01, 0.07]])
    e.set_state(st, np.zeros(3, np.int32))
    obs, r, term, trunc, _ = e.step(np.array([2, 0, 2]))
    s, _ = e.get_state()
    v0 = 0.0 + 1 * 0.001 + math.cos(3 * -0.5) * (-0.0025)
    np.testing.assert_allclose(s[:, 0], [-0.5 + v0, v0], rtol=1e-15)
    assert s[0, 1] == -1.2 and s[1, 1] == 0.0  # inelastic left wall
    assert term[2] and not term[0] and (r == -1.0).all()
    # default TimeLimit is 200
    e = E.OracleVecEnv("MountainCar-v0", 1, seed=1)
    e.reset()
    for t in range(200):
        _, _, term, trunc, info = e.step(np.array([1]))
    assert trunc[0] and not term[0] and info["episode"]["l"][0] == 200 and info["episode"]["r"][0] == -200.0


def test_acrobot_rest_state_and_termination_reward():
    e = E.OracleVecEnv("Acrobot-v1", 2)
    e.reset()
    st = np.zeros((4, 2))
    st[0, 1] = math.pi  # upright first link: -cos(pi) - cos(pi) = 2 > 1 -> terminal after the step
    e.set_state(st, np.zeros(2, np.int32))
    obs, r, term, trunc, _ = e.step(np.array([1, 1]))  # zero torque
    s, _ = e.get_state()
    # hanging at rest with zero torque is an equilibrium up to cos(-pi/2) rounding (6e-17 * g)
    assert np.abs(s[:, 0]).max() < 1e-14
    assert not term[0] and r[0] == -1.0
    assert term[1] and r[1] == 0.0
    np.testing.assert_allclose(obs[0], [1, 0, 1, 0, 0, 0], atol=1e-6)
    # velocities are clamped to +-4pi / +-9pi
    st = np.zeros((4, 1)); st[2, 0] = 100.0; st[3, 0] = -100.0
    e = E.OracleVecEnv("Acrobot-v1", 1); e.reset(); e.set_state(st, np.zeros(1, np.int32))
    e.step(np.array([0]))
    s, _ = e.get_state()
    assert abs(s[2, 0]) <= 4 * math.pi + 1e-12 and abs(s[3, 0]) <= 9 * math.pi + 1e-12
    assert -math.pi <= s[0, 0] <= math.pi and -math.pi <= s[1, 0] <= math.pi


def test_state_count_bonus_wrapper_matches_reference_formula():
    spec = dict(id="MountainCarV0_StateCountBonus", position_bins=50, velocity_bins=50, bonus_scale=0.1, bonus_type="count")
    e = E.OracleVecEnv("MountainCar-v0", 2, seed=2, env_wrappers=[spec])
    e.reset()
    counts = {0: {}, 1: {}}
    for t in range(60):
        obs, r, term, trunc, _ = e.step(np.array([0, 2]))
        for i in range(2):
            if t > 0 and prev_done[i]:
                assert r[i] == 0.0
                continue
            p, v = float(obs[i, 0]), float(obs[i, 1])  # float32 values promoted to float64 (numpy 1.26 scalar rules)
            pn = min(max((p - -1.2) / (0.6 - -1.2), 0.0), 0.999999)
            vn = min(max((v - -0.07) / (0.07 - -0.07), 0.0), 0.999999)
            key = (int(pn * 50), int(vn * 50))
            c = counts[i].get(key, 0)
            expect = -1.0 + 0.1 * (1.0 / math.sqrt(max(c, 1)))
            assert r[i] == expect
            counts[i][key] = c + 1
        prev_done = term | trunc


def test_cartpole_reward_shaper_is_potential_difference():
    spec = dict(id="CartPoleV1_RewardShaper", angle_reward_scale=1.0, position_reward_scale=0.25, clip_potential=True)
    e = E.OracleVecEnv("CartPole-v1", 1, seed=4, env_wrappers=[spec])
    obs, _ = e.reset()
    th = 12 * 2 * math.pi / 360

    def phi(o):
        x, t = float(o[0]), float(o[2])
        return 1.0 * min(max(1.0 - abs(t) / th, 0.0), 1.0) + 0.25 * min(max(1.0 - abs(x) / 2.4, 0.0), 1.0)

    for t in range(12):
        nobs, r, term, trunc, _ = e.step(np.array([1]))
        assert r[0] == 1.0 + (phi(nobs[0]) - phi(obs[0]))
        obs = nobs
        if term[0] or trunc[0]:
            break


WRAPPER_TAGS = ["mcar_count", "mcar_count_log", "mcar_count_inverse", "mcar_shaper", "cartpole_shaper", "cartpole_shaper_noclip"]


@pytest.mark.parametrize("tag", WRAPPER_TAGS)
def test_reward_wrappers_match_fixtures_generated_by_the_reference_wrappers(golden_dir, tag):
    """tests/golden/wrappers.npz holds rewards produced by EXECUTING the reference's wrapper classes
    (gym_wrappers/MountainCarV0/state_count_bonus.py, MountainCarV0/reward_shaper.py, CartPoleV1/reward_shaper.py) over a scripted
    sub-env replaying a physics trajectory through SyncVectorEnv's NEXT_STEP autoreset protocol.  The oracle with its restated
    wrapper attached must reproduce them on the same trajectory (several episodes, persistent count tables, potentials re-based
    at every reset)."""
    import json

    d = np.load(os.path.join(golden_dir, "wrappers.npz"))
    spec = json.loads(str(d[f"{tag}_kwargs"]))
    env_id = "CartPole-v1" if tag.startswith("cartpole") else "MountainCar-v0"
    env = E.OracleVecEnv(env_id, 1, seed=int(d[f"{tag}_seed"]), max_episode_steps=int(d[f"{tag}_max_steps"]), env_wrappers=[spec])
    env.reset()
    s0, e0 = env.get_state()
    np.testing.assert_array_equal(s0, d[f"{tag}_state0"])          # same physics start as the fixture's trajectory
    got, dones = [], []
    for a in d[f"{tag}_actions"]:
        _, r, te, tr, _ = env.step(np.array([a], dtype=np.int32))
        got.append(float(r[0])); dones.append(bool(te[0] or tr[0]))
    np.testing.assert_array_equal(np.array(dones), d[f"{tag}_done"])
    assert not np.allclose(d[f"{tag}_reward"], d[f"{tag}_base_reward"])          # the wrapper did something
    np.testing.assert_allclose(np.array(got), d[f"{tag}_reward"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Acrobot-v1", "MountainCar-v0"])
def test_envs_satisfy_the_reference_yaml_spec_blocks(golden_dir, env_id):
    """The physics itself is unpinned (gymnasium is not vendored), but the reference states each env's spaces, per-step reward,
    return range and action meanings in the `spec:` block of its YAML (config/environments/<env>.yaml, extracted into
    tests/golden/host_logic.json["env_specs"]).  The restatement must satisfy every one of those statements."""
    import json

    spec = json.load(open(os.path.join(golden_dir, "host_logic.json")))["env_specs"][env_id]
    n, steps = 64, 1200
    env = E.OracleVecEnv(env_id, n, seed=5)
    obs, _ = env.reset()
    assert env.single_action_space.n == spec["action_space"]["discrete"]
    assert list(obs.shape[1:]) == spec["observation_state"]["shape"] and str(obs.dtype) == spec["observation_state"]["dtype"]
    lo = np.array([float(c["range"][0]) for c in spec["observation_state"]["components"]])
    hi = np.array([float(c["range"][1]) for c in spec["observation_state"]["components"]])
    rng = np.random.default_rng(0)
    per_step, prev_done = float(spec["rewards"]["per_step"]), np.zeros(n, bool)
    ep_returns, ep_lengths = [], []
    for t in range(steps):
        obs, r, te, tr, info = env.step(rng.integers(0, env.single_action_space.n, n).astype(np.int32))
        assert (obs >= lo - 1e-6).all() and (obs <= hi + 1e-6).all()
        real = ~prev_done                                     # the step after a done is the autoreset step: reward 0, no flags
        if env_id == "Acrobot-v1":
            assert np.isin(r[real], (per_step, 0.0)).all() and (r[real & ~te] == per_step).all()      # 0 only on the terminating step
        else:
            assert (r[real] == per_step).all()
        assert (r[~real] == 0).all() and not (te | tr)[~real].any()
        done = te | tr
        if done.any():
            ep_returns += info["episode"]["r"][done].tolist()
            ep_lengths += info["episode"]["l"][done].tolist()
        prev_done = done
    assert len(ep_returns) > n
    rr = spec["returns"].get("range") or spec["rewards"]["range"]                # Acrobot states the return range under rewards
    assert min(ep_returns) >= rr[0] and max(ep_returns) <= rr[1]
    assert max(ep_lengths) <= max(abs(rr[0]), abs(rr[1]))                         # TimeLimit: 500 / 500 / 200 steps
    if env_id != "CartPole-v1":
        assert max(ep_lengths) == max(abs(rr[0]), abs(rr[1]))                     # random play runs into the time limit
    # action labels: the direction each action pushes
    labels = spec["action_space"]["labels"]
    one = E.OracleVecEnv(env_id, 1, seed=1)
    one.reset()
    s0, e0 = one.get_state()
    vel_index = {"CartPole-v1": 1, "MountainCar-v0": 1, "Acrobot-v1": 5}[env_id]
    after = {}
    for a in range(env.single_action_space.n):
        one.set_state(s0, e0)
        o, *_ = one.step(np.array([a], dtype=np.int32))
        after[a] = float(o[0, vel_index])
    neg = [int(k) for k, v in labels.items() if v in ("push_left", "torque_negative")][0]
    pos = [int(k) for k, v in labels.items() if v in ("push_right", "torque_positive")][0]
    assert after[neg] < after[pos]
    mid = [int(k) for k, v in labels.items() if v in ("no_push", "torque_zero")]
    if mid:
        assert after[neg] < after[mid[0]] < after[pos]
