"""The reference's collector known-answer tests that need a SCRIPTED env, on the DEVICE collector (fused collect kernel + device
targets): the fake envs of tests/test_rollouts_extra.py:161-240, tests/test_mc_baseline_mask.py:75-78 and tests/test_rollouts.py:95-123
are reproduced by the engine's `ScriptedReplay` device wrapper (table replay over MountainCar-v0's spaces); the scripted policies by an
MLP with hand-set weights whose argmax follows the observation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _collector(n_steps, script, *, n_envs=1, **kw):
    from gymnasium_solver_b200.gym_wrappers import EnvWrapperRegistry
    from gymnasium_solver_b200.utils.environment import build_env
    from gymnasium_solver_b200.utils.models import MLPActorCritic
    from gymnasium_solver_b200.utils.rollout_collector import RolloutCollector

    env = build_env("MountainCar-v0", n_envs=n_envs, seed=0)
    env = EnvWrapperRegistry.apply(env, dict(id="ScriptedReplay", **script))
    model = MLPActorCritic(input_shape=(2,), hidden_dims=(64, 64), output_shape=(3,), activation="relu").to(env.device)
    # logits follow the observation: units relu(x0), relu(x1), relu(-x0-x1) pass through both hidden layers to the three logits
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
        sd = model.state_dict()
        w1 = torch.zeros(64, 2); w1[0, 0] = 1; w1[1, 1] = 1; w1[2, 0] = -1; w1[2, 1] = -1
        w2 = torch.zeros(64, 64); w2[0, 0] = w2[1, 1] = w2[2, 2] = 1
        wp = torch.zeros(3, 64); wp[0, 0] = wp[1, 1] = wp[2, 2] = 10
        sd["backbone.0.weight"].copy_(w1); sd["backbone.2.weight"].copy_(w2); sd["policy_head.weight"].copy_(wp)
    return RolloutCollector(env, model, n_steps, **kw), env


ONE_HOT = {0: [1.0, 0.0], 1: [0.0, 1.0], 2: [-1.0, -1.0]}


def test_action_histogram_counts_and_metrics():
    """reference tests/test_rollouts_extra.py:161-178: actions 0,1,2,0,1 -> histogram [2,2,1], roll/actions/mean and std from it, reset zeroes."""
    obs = [ONE_HOT[t % 3] for t in range(6)]
    col, _ = _collector(5, dict(rewards=[0.0] * 5, terminated=[False] * 5, observations=obs), use_gae=True)
    traj = col.collect(deterministic=True)
    np.testing.assert_array_equal(traj.actions.cpu().numpy(), [0, 1, 2, 0, 1])
    counts = col.get_action_histogram_counts(reset=False)
    np.testing.assert_array_equal(counts, np.array([2, 2, 1], dtype=np.int64))
    m = col.get_metrics()
    idxs = np.arange(3, dtype=np.float32)
    mean = float((idxs * counts).sum() / counts.sum())
    std = float(np.sqrt(((idxs - mean) ** 2 * counts).sum() / counts.sum()))
    assert abs(m["roll/actions/mean"] - mean) < 1e-6 and abs(m["roll/actions/std"] - std) < 1e-6
    np.testing.assert_array_equal(col.get_action_histogram_counts(reset=True), [2, 2, 1])
    np.testing.assert_array_equal(col.get_action_histogram_counts(reset=False), [0, 0, 0])


TRAILING = dict(rewards=[1.0, 2.0, 3.0, 4.0], terminated=[False, True, False, False])


def test_slice_trajectories_remaps_trailing_partial_for_mc():
    """reference tests/test_rollouts_extra.py:229-240: terminal at t=1 of 4, gamma=1 -> returns [3,2,7,4]; indices 2, 3 (trailing partial
    episode) are remapped to the nearest previous valid index 1."""
    col, _ = _collector(4, TRAILING, use_gae=False, normalize_advantages=False, gamma=1.0)
    traj = col.collect(deterministic=True)
    flat = traj.returns.reshape(-1).cpu().numpy()
    np.testing.assert_array_equal(flat, np.array([3.0, 2.0, 7.0, 4.0], dtype=np.float32))
    out = col.slice_trajectories(traj, np.array([2, 3]))
    np.testing.assert_array_equal(out.returns.cpu().numpy(), flat[[1, 1]])


def test_mc_baseline_uses_masked_values_only():
    """reference tests/test_mc_baseline_mask.py:60-78: the running baseline sees only the valid positions (t <= last terminal):
    mean of returns {3, 2} = 2.5."""
    col, _ = _collector(4, TRAILING, use_gae=False, normalize_advantages=False, gamma=1.0, returns_type="mc:rtg")
    col.collect(deterministic=True)
    assert abs(col.get_metrics()["roll/baseline/mean"] - 2.5) < 1e-6


def test_mc_episode_returns_constant_within_episode():
    """reference tests/test_rollouts.py:95-123: one 5-step episode of reward 1, episode-mode MC returns: every step carries 5.0."""
    col, _ = _collector(5, dict(rewards=[1.0] * 5, terminated=[False] * 4 + [True]), use_gae=False, normalize_advantages=False, gamma=1.0,
                        returns_type="mc:episode")
    traj = col.collect(deterministic=True)
    rets = traj.returns.cpu().numpy().reshape(-1)
    assert rets.shape[0] == 5
    np.testing.assert_allclose(rets, np.full_like(rets, rets[0]))
    assert rets[0] == 5.0
    m = col.get_metrics()
    assert m["roll/ep_rew/mean"] == 5.0 and m["roll/ep_len/mean"] == 5       # RecordEpisodeStatistics of the scripted episode


def test_scripted_replay_several_envs_and_episode_statistics_restart():
    """The replay is per env (same tables), clamps at the last row, and episode statistics restart on the step after a done."""
    col, env = _collector(6, dict(rewards=[1.0, 2.0, 3.0], terminated=[False, True, False]), n_envs=5, use_gae=True)
    traj = col.collect(deterministic=True)
    r = traj.tm["rewards"].cpu().numpy()
    np.testing.assert_array_equal(r, np.tile(np.array([[1.0], [2.0], [3.0], [3.0], [3.0], [3.0]], np.float32), (1, 5)))
    d = traj.tm["dones"].cpu().numpy().astype(bool)
    np.testing.assert_array_equal(d[:, 0], [False, True, False, False, False, False])
    eplen = col._buffer.ep_length_buf[:6].cpu().numpy()
    assert (eplen[1] == 2).all() and (eplen[[0, 2, 3, 4, 5]] == 0).all()
