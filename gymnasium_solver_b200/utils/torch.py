"""Torch-side helpers kept from the reference surface (utils/torch.py): activation map, weight init, inference context,
batch normalisation and KL diagnostics.  None of these sit on the engine's hot path — they serve model construction,
checkpoint-compatible nn.Modules and host-side consumers."""
from __future__ import annotations

import math
from contextlib import contextmanager

import torch
import torch.nn as nn

ACTIVATION_MAPPING = {
    "tanh": nn.Tanh, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "elu": nn.ELU, "selu": nn.SELU, "gelu": nn.GELU,
    "silu": nn.SiLU, "swish": nn.SiLU, "identity": nn.Identity,
}


@contextmanager
def inference_ctx(*modules):
    """eval() + no_grad for the given modules, restoring train flags afterwards (utils/torch.py:28-49).  no_grad rather
    than inference_mode: the collector allocates its persistent device buffers inside this context and updates them in
    place later, which inference tensors forbid."""
    mods = []
    for m in modules:
        mods.extend(m if isinstance(m, (list, tuple)) else [m])
    mods = [m for m in mods if isinstance(m, nn.Module)]
    flags = [m.training for m in mods]
    try:
        for m in mods:
            m.eval()
        with torch.no_grad():
            yield
    finally:
        for m, f in zip(mods, flags):
            if f:
                m.train()


def _device_of(module) -> torch.device:
    if not hasattr(module, "parameters"):
        return torch.device("cpu")
    return next(module.parameters()).device


def assert_detached(*tensors) -> bool:
    for t in tensors:
        assert not t.requires_grad, "Tensor still requires grad"
        assert t.grad_fn is None, "Tensor is still connected to a computation graph"
    return True


def batch_normalize(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    return (x - x.mean()) / (x.std() + eps)


def compute_kl_diagnostics(old_logprobs, new_logprobs):
    diff = torch.clamp(new_logprobs - old_logprobs, min=-20.0, max=20.0)
    ratio = torch.exp(diff)
    return (old_logprobs - new_logprobs).mean(), ((ratio - 1) - torch.log(ratio)).mean()


def _gain_for(act: nn.Module):
    if isinstance(act, nn.ReLU):
        return nn.init.calculate_gain("relu")
    if isinstance(act, nn.LeakyReLU):
        return nn.init.calculate_gain("leaky_relu", act.negative_slope)
    if isinstance(act, nn.Tanh):
        return nn.init.calculate_gain("tanh")
    if isinstance(act, (nn.ELU, nn.GELU, nn.SiLU)):
        return nn.init.calculate_gain("relu")
    if isinstance(act, (nn.SELU, nn.Identity)):
        return 1.0
    return None


def init_model_weights(model: nn.Module, *, default_activation="relu", policy_heads=None, value_heads=None) -> None:
    """Orthogonal init (utils/torch.py:204-258): hidden layers use the gain of the activation that follows them,
    policy heads 0.01, value heads 1.0, all biases zero."""
    policy_heads, value_heads = tuple(policy_heads or ()), tuple(value_heads or ())
    heads = set(map(id, policy_heads + value_heads))
    act = default_activation
    if isinstance(act, str):
        act = ACTIVATION_MAPPING.get(act.lower(), nn.ReLU)()
    elif isinstance(act, type):
        act = act()
    default_gain = _gain_for(act) or 1.0
    follow = {}
    for seq in model.modules():
        if isinstance(seq, nn.Sequential):
            kids = list(seq.children())
            for a, b in zip(kids, kids[1:]):
                g = _gain_for(b)
                if isinstance(a, nn.Linear) and g is not None:
                    follow[id(a)] = g

    def init(layer, gain):
        nn.init.orthogonal_(layer.weight, gain=gain)
        if layer.bias is not None:
            nn.init.constant_(layer.bias, 0.0)

    for m in model.modules():
        if isinstance(m, nn.Linear) and id(m) not in heads:
            init(m, follow.get(id(m), default_gain))
    for h in policy_heads:
        init(h, 0.01)
    for h in value_heads:
        init(h, 1.0)


def compute_param_group_grad_norm(params) -> float:
    sq, seen = 0.0, False
    for p in params:
        g = getattr(p, "grad", None)
        if g is None:
            continue
        seen = True
        sq += float(g.detach().norm(2).item() ** 2)
    return math.sqrt(sq) if seen else 0.0
