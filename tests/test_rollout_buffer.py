"""RolloutBuffer / RolloutTrajectory protocol surface (reference: utils/rollout_buffer.py:16-173, pinned there by
tests/test_rollout_buffer.py) on a CPU torch device — the class is device-agnostic storage; the kernels write the same arrays on
the GPU.  The env-major order is checked against tests/golden/buffer_env_major.npz, produced by running the reference's own
RolloutBuffer.add + flatten_slice_env_major on the same inputs."""
import os

import numpy as np
import pytest
import torch

from gymnasium_solver_b200.utils.rollout_buffer import RolloutBuffer, RolloutTrajectory

CPU = torch.device("cpu")


def _step(n, obs_shape, t, rng):
    return dict(obs=rng.standard_normal((n, *obs_shape)).astype(np.float32), next_obs=rng.standard_normal((n, *obs_shape)).astype(np.float32),
                actions=rng.integers(0, 3, n), logps=rng.standard_normal(n).astype(np.float32), values=rng.standard_normal(n).astype(np.float32),
                rewards=np.full(n, float(t), np.float32), dones=rng.random(n) < 0.4, timeouts=rng.random(n) < 0.2)


def test_begin_rollout_wraps_and_tracks_the_high_water_mark():
    buf = RolloutBuffer(2, (3,), np.float32, CPU, maxsize=5)
    assert (buf.begin_rollout(3), buf.pos, buf.size) == (0, 3, 3)
    assert (buf.begin_rollout(2), buf.pos, buf.size) == (3, 5, 5)
    assert (buf.begin_rollout(2), buf.pos, buf.size) == (0, 2, 5)      # the tail is too short: wrap to 0, size keeps the maximum
    with pytest.raises(ValueError, match="exceeds buffer maxsize"):
        RolloutBuffer(1, (1,), np.float32, CPU, maxsize=4).begin_rollout(5)


def test_add_checks_the_observation_shape_and_keeps_dtypes():
    buf = RolloutBuffer(2, (2, 2), np.float32, CPU, maxsize=3)
    idx = buf.begin_rollout(1)
    obs = np.arange(8, dtype=np.float32).reshape(2, 2, 2)
    z = np.zeros(2, np.float32)
    buf.add(idx, obs, obs + 1, np.array([1, 2]), z + 0.1, z + 0.3, z, np.array([False, True]), np.array([False, False]))
    np.testing.assert_array_equal(buf.obs_buf[idx].numpy(), obs)
    np.testing.assert_array_equal(buf.next_obs_buf[idx].numpy(), obs + 1)
    assert buf.logprobs_buf.dtype == buf.values_buf.dtype == buf.rewards_buf.dtype == torch.float32
    assert buf.dones_buf[idx].tolist() == [0, 1]
    with pytest.raises(AssertionError, match="Expected shape"):
        buf.add(idx, np.zeros((2, 3), np.float32), obs, np.array([0, 0]), z, z, z, z > 1, z > 1)


def test_flatten_is_env_major_like_the_reference(golden_dir):
    d = np.load(os.path.join(golden_dir, "buffer_env_major.npz"))
    T, n, D = d["in_obs"].shape
    buf = RolloutBuffer(n, (D,), np.float32, CPU, maxsize=T)
    start = buf.begin_rollout(T)
    for t in range(T):
        buf.add(start + t, d["in_obs"][t], d["in_next_obs"][t], d["in_actions"][t], d["in_logps"][t], d["in_values"][t], d["in_rewards"][t],
                d["in_dones"][t], d["in_timeouts"][t])
    traj = buf.flatten_slice_env_major(start, start + T, d["in_adv"], d["in_ret"])
    assert isinstance(traj, RolloutTrajectory) and traj._fields == ("observations", "actions", "rewards", "dones", "logprobs", "values",
                                                                    "advantages", "returns", "next_observations")
    for name in traj._fields:
        got, want = getattr(traj, name).numpy(), d[f"out_{name}"]
        assert got.dtype == want.dtype, (name, got.dtype, want.dtype)
        np.testing.assert_array_equal(got, want, err_msg=name)
    # sample i = env * T + t
    np.testing.assert_array_equal(traj.observations[1 * T + 2].numpy(), d["in_obs"][2, 1])


def test_scalar_observations_flatten_to_a_column_and_flags_propagate():
    T, n = 4, 3
    rng = np.random.default_rng(3)
    buf = RolloutBuffer(n, (), np.int64, CPU, maxsize=T)
    start = buf.begin_rollout(T)
    steps = []
    for t in range(T):
        s = _step(n, (), t, rng)
        s["obs"], s["next_obs"] = rng.integers(0, 16, n), rng.integers(0, 16, n)
        steps.append(s)
        buf.add(start + t, s["obs"], s["next_obs"], s["actions"], s["logps"], s["values"], s["rewards"], s["dones"], s["timeouts"])
    adv = np.zeros((T, n), np.float32)
    traj = buf.flatten_slice_env_major(start, start + T, adv, adv + 1)
    assert traj.observations.shape == (n * T, 1) and traj.next_observations.shape == (n * T, 1) and traj.observations.dtype == torch.int64
    for e in range(n):
        for t in range(T):
            i = e * T + t
            assert int(traj.observations[i, 0]) == int(steps[t]["obs"][e])
            assert bool(traj.dones[i]) == bool(steps[t]["dones"][e]) and float(traj.rewards[i]) == float(t)
    assert traj.dones.dtype == torch.bool and traj.actions.dtype == torch.int64
    assert float(traj.returns.min()) == 1.0 and float(traj.advantages.max()) == 0.0


def test_vector_actions_are_kept_as_float_rows():
    buf = RolloutBuffer(2, (3,), np.float32, CPU, maxsize=2, action_shape=(4,), action_dtype=np.int8)     # MultiBinary(4)-style actions
    idx = buf.begin_rollout(1)
    z = np.zeros(2, np.float32)
    acts = np.array([[1, 0, 1, 1], [0, 0, 1, 0]])
    buf.add(idx, np.zeros((2, 3), np.float32), np.zeros((2, 3), np.float32), acts, z, z, z, z > 1, z > 1)
    traj = buf.flatten_slice_env_major(idx, idx + 1, z[None], z[None])
    assert traj.actions.shape == (2, 4) and traj.actions.dtype == torch.float32
    np.testing.assert_array_equal(traj.actions.numpy(), acts.astype(np.float32))
