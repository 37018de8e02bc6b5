"""EnvWrapperRegistry (reference: gym_wrappers/env_wrapper_registry.py:1-16): name -> wrapper class, applied from the
YAML ``env_wrappers`` list as ``wrapper_cls(env, **kwargs)``.  Unknown ids raise KeyError like the reference."""


class EnvWrapperRegistry:
    _registry = {}

    @classmethod
    def register(cls, wrapper_classes):
        if type(wrapper_classes) is not list:
            wrapper_classes = [wrapper_classes]
        for wrapper_cls in wrapper_classes:
            cls._registry[wrapper_cls.__name__] = wrapper_cls

    @classmethod
    def apply(cls, env, wrapper_spec):
        wrapper_id = wrapper_spec["id"]
        kwargs = {k: v for k, v in wrapper_spec.items() if k != "id"}
        return cls._registry[wrapper_id](env, **kwargs)
