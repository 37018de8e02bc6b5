"""Environment construction (reference: utils/environment.py:115-425) for the engine's device environments.

``build_env`` keeps the reference's keyword surface; the Gymnasium pipeline it used to assemble per sub-env
(gym.make -> YAML wrappers -> TimeLimit -> SyncVectorEnv -> RecordEpisodeStatistics) is one DeviceVecEnv handle here."""
from __future__ import annotations

from ..envs.device_vec_env import DeviceVecEnv
from ..gym_wrappers import EnvWrapperRegistry

DEVICE_ENV_IDS = ("CartPole-v1", "Acrobot-v1", "MountainCar-v0")


def get_env_type(env_id: str) -> str:
    return "device" if env_id in DEVICE_ENV_IDS else "unsupported"


def build_env(env_id: str, *, n_envs: int = 1, seed: int = 0, max_episode_steps=None, env_wrappers=(), env_spec=None, device=None,
              env_id_offset: int = 0, **unused) -> DeviceVecEnv:
    if env_id not in DEVICE_ENV_IDS:
        raise KeyError(f"{env_id!r} is outside the b200 engine's scope; device envs: {DEVICE_ENV_IDS}")
    for k in ("frame_stack", "frame_skip", "grayscale_obs", "resize_obs", "record_video"):
        if unused.get(k):
            raise ValueError(f"{k} is not supported for device environments")
    normalize_obs = unused.get("normalize_obs")
    if normalize_obs in (True, "rolling"):            # gymnasium's NormalizeObservation (running statistics): not built
        raise ValueError("normalize_obs='rolling' is not supported for device environments; use 'static' (VecNormalizeStatic)")
    if normalize_obs not in (None, False, "static"):
        raise ValueError(f"unknown normalize_obs: {normalize_obs!r}")
    env = DeviceVecEnv(env_id, int(n_envs), int(seed), max_episode_steps=max_episode_steps, device=device,
                       env_id_offset=env_id_offset, spec=env_spec)
    for spec in env_wrappers or ():
        env = EnvWrapperRegistry.apply(env, spec)
    if normalize_obs == "static":                      # utils/environment.py:215-216: after the per-env wrappers and the episode statistics
        env.normalize_observations_static()
    return env


def build_env_from_config(config, **kwargs) -> DeviceVecEnv:
    args = config.get_env_args()
    args.update(kwargs)
    return build_env(args.pop("env_id"), **args)
