"""EnvWrapperRegistry (reference: gym_wrappers/env_wrapper_registry.py:1-16): wrapper classes looked up by ``__name__`` and applied
from the YAML ``env_wrappers`` list as ``wrapper_cls(env, **kwargs)``.  On the engine the registered classes are the device wrappers
of gym_wrappers/device_wrappers.py (they attach a fused reward term to the env handle instead of wrapping ``step``)."""
from __future__ import annotations

from typing import Any, Dict, List, Mapping


class EnvWrapperRegistry:
    _registry: Dict[str, type] = {}

    @classmethod
    def register(cls, wrapper_classes) -> None:
        """One class or a ``list`` of classes (exactly a list, like the reference: a tuple is taken for a single entry and fails on
        ``__name__``)."""
        batch = wrapper_classes if type(wrapper_classes) is list else [wrapper_classes]
        cls._registry.update((w.__name__, w) for w in batch)

    @classmethod
    def apply(cls, env, wrapper_spec: Mapping[str, Any]):
        """``{"id": name, **kwargs}`` -> ``registered[name](env, **kwargs)``; KeyError for an unknown id or a spec without one."""
        spec = dict(wrapper_spec)
        factory = cls._registry[spec.pop("id")]
        return factory(env, **spec)

    @classmethod
    def registered(cls) -> List[str]:
        return sorted(cls._registry)
