"""MLP policy / actor-critic modules (reference: utils/models.py:20-346).

These nn.Modules exist so that parameters, ``state_dict`` keys (``backbone.{0,2}.*``, ``policy_head.*``, ``value_head.*``),
checkpoints and the torch optimizer stay exactly what the reference produces.  Their parameters are views into ONE flat
fp32 buffer (and their ``.grad`` views into one flat gradient buffer) so the update kernel writes gradients in place and a
single all-reduce covers the model.  ``forward`` is the plain torch definition for host-side consumers; the engine's
collect and update kernels read the parameter storage directly and never call it."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .policy_ops import create_action_distribution
from .torch import ACTIVATION_MAPPING, compute_param_group_grad_norm, init_model_weights


def build_mlp(input_shape, hidden_dims, activation: str) -> nn.Sequential:
    is_int = type(input_shape) in (int, np.int32, np.int64)
    assert is_int or len(input_shape) == 1, "Input shape must be 1D"
    act_cls = ACTIVATION_MAPPING[activation.lower()]
    layers, dims = [], list(hidden_dims)
    if is_int:
        layers.append(nn.Embedding(int(input_shape), dims[0]))
        last, dims = dims[0], dims[1:]
    else:
        last = int(input_shape[0])
    for h in dims:
        layers += [nn.Linear(last, h), act_cls()]
        last = h
    return nn.Sequential(*layers)


class BaseModel(nn.Module):
    """Grad-norm and activation-stat reporting surface (utils/models.py:113-230).  On the engine path both are computed
    on device by gs_clip_grad_norm / the update kernel; ``set_engine_metrics`` feeds them back in here."""

    def __init__(self):
        super().__init__()
        self._activation_stats: Dict[str, Dict[str, float]] = {}
        self._track_activations = False
        self._engine_grad_norms: Dict[str, float] | None = None
        self._flat_params = None
        self._flat_grads = None

    # -- flat storage --------------------------------------------------------------------------------------------
    def flatten_parameters_(self):
        """Re-home every parameter (and its grad) as a view into one contiguous fp32 buffer, parameters() order."""
        params = list(self.parameters())
        dev = params[0].device
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grads = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = flat[off:off + n].view_as(p)
            p.grad = grads[off:off + n].view_as(p)
            off += n
        self._flat_params, self._flat_grads = flat, grads
        return flat, grads

    @property
    def flat_params(self):
        return self._flat_params

    @property
    def flat_grads(self):
        return self._flat_grads

    # -- metrics --------------------------------------------------------------------------------------------------
    def set_engine_metrics(self, grad_norms=None, activation_stats=None) -> None:
        if grad_norms is not None:
            self._engine_grad_norms = dict(grad_norms)
        if activation_stats is not None:
            self._activation_stats = dict(activation_stats)

    def compute_activation_stats(self) -> Dict[str, float]:
        out = {}
        for name, st in self._activation_stats.items():
            for k in ("mean", "std", "dead_pct", "dead_max"):
                out[f"opt/activations/{name}/{k}"] = st[k]
        self._activation_stats = {}
        return out

    def compute_grad_norms(self) -> Dict[str, float]:
        if self._engine_grad_norms is not None:
            out, self._engine_grad_norms = self._engine_grad_norms, None
            return out
        groups = {n: list(getattr(self, n).parameters()) for n in ("backbone", "policy_head", "value_head")
                  if isinstance(getattr(self, n, None), nn.Module)}
        out = {"opt/grads/norm/all": compute_param_group_grad_norm(list(self.parameters()))}
        for n, ps in groups.items():
            out[f"opt/grads/norm/{n}"] = compute_param_group_grad_norm(ps)
        return out


class MLPPolicy(BaseModel):
    def __init__(self, input_dim=None, hidden_dims=(64,), output_dim=None, activation: str = "relu", *, input_shape=None,
                 output_shape=None, valid_actions=None, action_space_type: str = "discrete"):
        super().__init__()
        if input_dim is None:
            assert input_shape is not None and len(input_shape) == 1, "Input shape must be 1D"
            input_dim = int(input_shape[0])
        if output_dim is None:
            assert output_shape is not None and len(output_shape) == 1, "Output shape must be 1D"
            output_dim = int(output_shape[0])
        if isinstance(hidden_dims, int):
            hidden_dims = (hidden_dims,)
        self.backbone = build_mlp((input_dim,), hidden_dims, activation)
        self.policy_head = nn.Linear(hidden_dims[-1], output_dim)
        self.valid_actions, self.action_space_type = valid_actions, action_space_type
        init_model_weights(self, default_activation=activation, policy_heads=[self.policy_head])

    def forward(self, obs):
        return create_action_distribution(self.policy_head(self.backbone(obs)), self.valid_actions, self.action_space_type), None


class MLPActorCritic(BaseModel):
    def __init__(self, *, input_shape, hidden_dims, output_shape, activation: str, valid_actions=None,
                 action_space_type: str = "discrete"):
        super().__init__()
        assert type(input_shape) in (int, np.int32, np.int64) or len(input_shape) == 1, "Input shape must be 1D"
        assert len(output_shape) == 1, "Output shape must be 1D"
        self.backbone = build_mlp(input_shape, hidden_dims, activation)
        self.policy_head = nn.Linear(hidden_dims[-1], output_shape[0])
        self.value_head = nn.Linear(hidden_dims[-1], 1)
        self.valid_actions, self.action_space_type = valid_actions, action_space_type
        init_model_weights(self, default_activation=activation, policy_heads=[self.policy_head], value_heads=[self.value_head])

    def forward(self, obs):
        x = self.backbone(obs)
        if x.ndim > 2 and x.shape[1] == 1:
            x = x.squeeze(1)
        dist = create_action_distribution(self.policy_head(x), self.valid_actions, self.action_space_type)
        return dist, self.value_head(x).squeeze(-1)
