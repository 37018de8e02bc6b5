"""Device versions of the reference's per-env reward wrappers, registered under the SAME ids so YAML ``env_wrappers`` entries
resolve unchanged.  "Wrapping" a DeviceVecEnv attaches the wrapper's arithmetic to the env handle: it is fused into the
step / collect kernels (csrc/env_dynamics.cuh), and the call returns the same env object.

  MountainCarV0_StateCountBonus   gym_wrappers/MountainCarV0/state_count_bonus.py:9-126  (per-env 2-D visit table in HBM)
  CartPoleV1_RewardShaper         gym_wrappers/CartPoleV1/reward_shaper.py:6-77
  MountainCarV0_RewardShaper      gym_wrappers/MountainCarV0/reward_shaper.py:6-102
"""
from __future__ import annotations

from ..envs.device_vec_env import DeviceVecEnv


def _require_device_env(env, wrapper_id, env_id):
    if not isinstance(env, DeviceVecEnv):
        raise TypeError(f"{wrapper_id} (device version) wraps a DeviceVecEnv, got {type(env).__name__}")
    if env.env_id != env_id:
        raise ValueError(f"{wrapper_id} applies to {env_id}, not {env.env_id}")
    return env


def MountainCarV0_StateCountBonus(env, position_bins=50, velocity_bins=50, bonus_scale=1.0, bonus_type="count", min_count=1):
    env = _require_device_env(env, "MountainCarV0_StateCountBonus", "MountainCar-v0")
    codes = {"count": 0, "inverse": 1, "log": 2}
    if bonus_type not in codes:
        raise ValueError(f"Unknown bonus_type: {bonus_type}")
    env.attach_wrapper("MountainCarV0_StateCountBonus", [position_bins, velocity_bins, bonus_scale, codes[bonus_type], min_count],
                       dict(id="MountainCarV0_StateCountBonus", position_bins=position_bins, velocity_bins=velocity_bins,
                            bonus_scale=bonus_scale, bonus_type=bonus_type, min_count=min_count))
    return env


def CartPoleV1_RewardShaper(env, angle_reward_scale: float = 1.0, position_reward_scale: float = 0.25, clip_potential: bool = True):
    env = _require_device_env(env, "CartPoleV1_RewardShaper", "CartPole-v1")
    env.attach_wrapper("CartPoleV1_RewardShaper", [angle_reward_scale, position_reward_scale, 1.0 if clip_potential else 0.0],
                       dict(id="CartPoleV1_RewardShaper", angle_reward_scale=angle_reward_scale,
                            position_reward_scale=position_reward_scale, clip_potential=clip_potential))
    return env


def MountainCarV0_RewardShaper(env, position_reward_scale=100.0, velocity_reward_scale=10.0, height_reward_scale=50.0):
    env = _require_device_env(env, "MountainCarV0_RewardShaper", "MountainCar-v0")
    env.attach_wrapper("MountainCarV0_RewardShaper", [position_reward_scale, velocity_reward_scale, height_reward_scale],
                       dict(id="MountainCarV0_RewardShaper", position_reward_scale=position_reward_scale,
                            velocity_reward_scale=velocity_reward_scale, height_reward_scale=height_reward_scale))
    return env


def ScriptedReplay(env, rewards=(), terminated=(), truncated=None, observations=None):
    """Engine-only test double (no reference counterpart in gym_wrappers/): a MountainCar-v0-shaped env (2 observations, 3 actions)
    that ignores its actions and replays per-step tables -- the scripted fake envs of the reference's collector tests
    (tests/test_rollouts_extra.py:161-240, tests/test_mc_baseline_mask.py, tests/test_rollouts.py:95-123) as a DEVICE env, so their
    known answers are checked on the fused collect kernel.  Step k pays rewards[k] with terminated[k] / truncated[k] (k clamps at
    the last entry) and shows observations[k+1]; reset shows observations[0].  No autoreset, no time limit."""
    env = _require_device_env(env, "ScriptedReplay", "MountainCar-v0")
    L = len(rewards)
    if L < 1 or len(terminated) != L:
        raise ValueError("ScriptedReplay: rewards and terminated must have the same non-zero length")
    truncated = [False] * L if truncated is None else list(truncated)
    observations = [[0.0, 0.0]] * (L + 1) if observations is None else [list(o) for o in observations]
    if len(truncated) != L or len(observations) != L + 1 or any(len(o) != 2 for o in observations):
        raise ValueError("ScriptedReplay: truncated needs L entries, observations L + 1 rows of 2")
    params = [float(L)] + [float(r) for r in rewards] + [1.0 if t else 0.0 for t in terminated] + [1.0 if t else 0.0 for t in truncated]
    for o in observations:
        params += [float(o[0]), float(o[1])]
    env.attach_wrapper("ScriptedReplay", params, dict(id="ScriptedReplay"))
    return env
