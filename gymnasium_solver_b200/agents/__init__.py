"""build_agent (reference: agents/__init__.py:1-8)."""


def build_agent(config, **kw):
    algo_id = config.algo_id
    if algo_id == "ppo":
        from .ppo.ppo_agent import PPOAgent
        return PPOAgent(config, **kw)
    if algo_id == "reinforce":
        from .reinforce.reinforce_agent import REINFORCEAgent
        return REINFORCEAgent(config, **kw)
    raise ValueError(f"Unknown algo_id: {algo_id}")
