"""Wrapper registration (reference: gym_wrappers/__init__.py:29-46) — only the wrappers of the three device envs."""
from .device_wrappers import CartPoleV1_RewardShaper, MountainCarV0_RewardShaper, MountainCarV0_StateCountBonus, ScriptedReplay
from .env_wrapper_registry import EnvWrapperRegistry

EnvWrapperRegistry.register([MountainCarV0_StateCountBonus, CartPoleV1_RewardShaper, MountainCarV0_RewardShaper, ScriptedReplay])

__all__ = ["EnvWrapperRegistry", "MountainCarV0_StateCountBonus", "CartPoleV1_RewardShaper", "MountainCarV0_RewardShaper"]
