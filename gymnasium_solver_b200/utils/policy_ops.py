"""policy_act / policy_predict_values (reference: utils/policy_ops.py:14-75).

For the engine-supported MLPs these run the fused CUDA forward + categorical sample / log-prob kernel (gs_policy_act);
there is no torch fallback on that path.  ``create_action_distribution`` is kept for ``model.forward`` consumers."""
from __future__ import annotations

import ctypes as C

import torch
from torch.distributions import Bernoulli, Categorical, Independent

from .. import _native as N
from .distributions import MaskedCategorical

_counter = [0]


def policy_act(model, obs, *, deterministic: bool = False, return_dist: bool = False, rng_seed: int | None = None,
               rng_offset: int | None = None, uniforms=None):
    """(actions int64, logprobs, values[, dist]) for a batch of observations, computed by the engine."""
    obs_t = torch.as_tensor(obs, dtype=torch.float32)
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise N.EngineError("policy_act runs on the CUDA engine: move the model to a CUDA device")
    obs_t = obs_t.to(dev).contiguous()
    if getattr(model, "valid_actions", None) is not None:
        raise N.EngineError("action masking is not supported by the engine kernels")
    m = N.mlp_struct(model)
    n = obs_t.shape[0]
    a = torch.empty(n, dtype=torch.int32, device=dev)
    lp = torch.empty(n, dtype=torch.float32, device=dev)
    v = torch.empty(n, dtype=torch.float32, device=dev)
    logits = torch.empty(n, m.n_actions, dtype=torch.float32, device=dev) if return_dist else None
    if rng_seed is None:
        rng_seed = torch.initial_seed() & (2**63 - 1)
    if rng_offset is None:
        rng_offset = _counter[0]
        _counter[0] += 1
    u = None if uniforms is None else torch.as_tensor(uniforms, dtype=torch.float32, device=dev).contiguous()
    with torch.cuda.device(dev):
        N.check(N.lib().gs_policy_act(C.byref(m), N.ptr(obs_t), n, rng_seed, rng_offset, 0, int(deterministic), N.ptr(u), N.ptr(a),
                                      N.ptr(lp), N.ptr(v), N.ptr(logits), N.stream()))
    if return_dist:
        return a.long(), lp, v, Categorical(logits=logits)
    return a.long(), lp, v


def policy_predict_values(model, obs):
    obs_t = torch.as_tensor(obs, dtype=torch.float32)
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise N.EngineError("policy_predict_values runs on the CUDA engine: move the model to a CUDA device")
    obs_t = obs_t.to(dev).contiguous()
    m = N.mlp_struct(model)
    v = torch.empty(obs_t.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().gs_policy_values(C.byref(m), N.ptr(obs_t), obs_t.shape[0], N.ptr(v), N.stream()))
    return v


def create_action_distribution(logits, valid_actions, action_space_type: str = "discrete"):
    assert action_space_type in ("discrete", "multibinary"), \
        f"action_space_type must be 'discrete' or 'multibinary', got {action_space_type}"
    if valid_actions is not None:
        mask = torch.ones_like(logits, dtype=torch.bool)
        mask[:, valid_actions] = False
        logits = logits.masked_fill(mask, float("-inf"))
    if action_space_type == "multibinary":
        return Independent(Bernoulli(probs=torch.sigmoid(logits)), 1)
    return MaskedCategorical(logits=logits) if valid_actions is not None else Categorical(logits=logits)
