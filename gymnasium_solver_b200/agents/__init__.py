"""Agent factory (reference: agents/__init__.py:1-8): ``build_agent(config, **engine_kwargs)`` -> the agent class of ``config.algo_id``.

``engine_kwargs`` are the engine's own constructor options (``device``, ``rank``, ``world_size``); under ``torchrun`` rank and world size
come from the environment and none are needed.  An ``algo_id`` without an engine agent raises ``ValueError`` (the reference falls
through and returns ``None``)."""
import importlib

# algo_id -> (module below this package, class): imported on demand so loading one agent does not pull in the other
_AGENTS = {"ppo": ("ppo.ppo_agent", "PPOAgent"), "reinforce": ("reinforce.reinforce_agent", "REINFORCEAgent")}


def build_agent(config, **engine_kwargs):
    try:
        module, cls_name = _AGENTS[config.algo_id]
    except KeyError:
        raise ValueError(f"Unknown algo_id: {config.algo_id}") from None
    agent_cls = getattr(importlib.import_module(f"{__name__}.{module}"), cls_name)
    return agent_cls(config, **engine_kwargs)
