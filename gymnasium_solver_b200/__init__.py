"""B200-native rollout-and-update engine behind gymnasium-solver's BaseAgent / RolloutCollector / RolloutBuffer /
EnvWrapperRegistry / ``train.py <env>:<algo>`` surface.  See DESIGN.md and include/gs_engine.h."""

__all__ = ["build_agent", "load_config"]


def build_agent(config, **kw):
    from .agents import build_agent as _b
    return _b(config, **kw)


def load_config(config_id, variant_id=None, config_dir="config/environments"):
    from .utils.config import load_config as _l
    return _l(config_id, variant_id, config_dir)
