"""Policy construction from env spaces + config (reference: utils/policy_factory.py:55-130)."""
from __future__ import annotations

from .models import MLPActorCritic, MLPPolicy

_POLICIES = {"mlp": MLPPolicy, "mlp_actorcritic": MLPActorCritic}


def build_policy(policy_type, *, input_shape, hidden_dims, output_shape, activation, **policy_kwargs):
    key = getattr(policy_type, "value", policy_type)
    if key not in _POLICIES:
        raise KeyError(f"policy {key!r} is outside the engine's scope (mlp, mlp_actorcritic)")
    return _POLICIES[key](input_shape=input_shape, hidden_dims=hidden_dims, output_shape=output_shape, activation=activation,
                          **policy_kwargs)


def build_policy_from_env_and_config(env, config):
    obs_space = getattr(env, "single_observation_space", None) or env.observation_space
    act_space = getattr(env, "single_action_space", None) or env.action_space
    output_shape = act_space.shape or (act_space.n,)
    valid_actions = None
    spec = getattr(config, "spec", None) or {}
    if "valid" in (spec.get("action_space") or {}):
        valid_actions = spec["action_space"]["valid"]
    kwargs = dict(config.policy_kwargs)
    if valid_actions is not None:
        kwargs["valid_actions"] = valid_actions
    kwargs["action_space_type"] = "discrete"
    return build_policy(config.policy, input_shape=obs_space.shape, output_shape=output_shape, hidden_dims=config.hidden_dims,
                        activation=config.activation, **kwargs)
