"""torch restatement of the policy / loss side of the reference — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates (plain torch, autograd for the gradients):
  utils/models.py:285-346        MLPActorCritic.forward (shared backbone, policy_head, value_head)
  utils/models.py:233-282        MLPPolicy.forward
  utils/policy_ops.py:14-34      policy_act (sample / mode, log_prob, value)
  utils/distributions.py:8-82    MaskedCategorical entropy
  agents/ppo/ppo_agent.py:21-152 PPO clipped surrogate + clipped value loss + entropy + metrics
  agents/reinforce/reinforce_agent.py:11-88  REINFORCE loss
  utils/torch.py:97-145          batch_normalize, KL diagnostics
  utils/models.py:196-230 + agents/base_agent.py:591-621  grad norms and global-norm clip
Pinned by tests/golden/ppo_*.npz and reinforce_*.npz, produced by running the reference's own
PPOAgent / REINFORCEAgent.losses_for_batch in the build container (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

PARAM_ORDER = ("w1", "b1", "w2", "b2", "wp", "bp", "wv", "bv")


def init_params(obs_dim, hidden, n_actions, has_value=True, seed=0, dtype=torch.float32):
    """Orthogonal init like utils/torch.py:204-258 (gain sqrt(2) hidden, 0.01 policy head, 1.0 value head)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    dims = [obs_dim] + list(hidden)
    names = [("w1", "b1"), ("w2", "b2")]
    for i, (wn, bn) in enumerate(names[: len(hidden)]):
        w = torch.empty(dims[i + 1], dims[i])
        torch.nn.init.orthogonal_(w, gain=math.sqrt(2.0), generator=g)
        p[wn], p[bn] = w.to(dtype), torch.zeros(dims[i + 1], dtype=dtype)
    wp = torch.empty(n_actions, dims[-1])
    torch.nn.init.orthogonal_(wp, gain=0.01, generator=g)
    p["wp"], p["bp"] = wp.to(dtype), torch.zeros(n_actions, dtype=dtype)
    if has_value:
        wv = torch.empty(1, dims[-1])
        torch.nn.init.orthogonal_(wv, gain=1.0, generator=g)
        p["wv"], p["bv"] = wv.to(dtype), torch.zeros(1, dtype=dtype)
    return p


def random_params(obs_dim, hidden, n_actions, has_value=True, seed=0, scale=0.5, dtype=torch.float32):
    """Dense random weights with non-zero biases: exercises every term of the gradient."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    dims = [obs_dim] + list(hidden)
    for i, (wn, bn) in enumerate([("w1", "b1"), ("w2", "b2")][: len(hidden)]):
        p[wn] = (torch.randn(dims[i + 1], dims[i], generator=g) * scale / math.sqrt(dims[i])).to(dtype)
        p[bn] = (torch.randn(dims[i + 1], generator=g) * 0.1).to(dtype)
    p["wp"] = (torch.randn(n_actions, dims[-1], generator=g) * scale / math.sqrt(dims[-1])).to(dtype)
    p["bp"] = (torch.randn(n_actions, generator=g) * 0.1).to(dtype)
    if has_value:
        p["wv"] = (torch.randn(1, dims[-1], generator=g) * scale / math.sqrt(dims[-1])).to(dtype)
        p["bv"] = (torch.randn(1, generator=g) * 0.1).to(dtype)
    return p


def _act(x, activation):
    return torch.relu(x) if activation == "relu" else torch.tanh(x)


def forward(p, obs, activation="relu", return_pre=False):
    """logits (B,A), value (B,) or None — utils/models.py:328-346."""
    z1 = F.linear(obs, p["w1"], p["b1"])
    h = _act(z1, activation)
    pre = [z1]
    if "w2" in p:
        z2 = F.linear(h, p["w2"], p["b2"])
        h = _act(z2, activation)
        pre.append(z2)
    logits = F.linear(h, p["wp"], p["bp"])
    value = F.linear(h, p["wv"], p["bv"]).squeeze(-1) if "wv" in p else None
    if return_pre:
        return logits, value, pre
    return logits, value


def categorical_entropy(logits):
    """torch.distributions.Categorical.entropy (logits clamped at finfo.min before p*log p)."""
    logp = logits - logits.logsumexp(dim=-1, keepdim=True)
    logp_c = torch.clamp(logp, min=torch.finfo(logp.dtype).min)
    return -(logp_c * logp.exp()).sum(-1)


def masked_categorical_entropy(logits):
    """utils/distributions.py:34-58: -sum p*log(p+1e-8) over finite-logit entries."""
    mask = torch.isfinite(logits)
    logp = logits - logits.logsumexp(dim=-1, keepdim=True)
    p = logp.exp()
    log_p = torch.where(mask, torch.log(p + 1e-8), torch.zeros_like(p))
    return -(p * log_p).sum(-1)


def act(p, obs, *, deterministic=False, uniforms=None, activation="relu"):
    """policy_act — utils/policy_ops.py:14-34.  Sampling convention of the engine: inverse CDF over
    softmax(logits) with u ~ U[0,1): action = min{k : cdf_k > u} (clamped to A-1)."""
    logits, value = forward(p, obs, activation)
    logp_all = logits - logits.logsumexp(dim=-1, keepdim=True)
    if deterministic:
        actions = logits.argmax(dim=-1)  # dist.mode
    else:
        cdf = logp_all.exp().cumsum(-1)
        u = uniforms.to(cdf.dtype).unsqueeze(-1)
        actions = (cdf <= u).sum(-1).clamp(max=logits.shape[-1] - 1)
    logp = logp_all.gather(-1, actions.unsqueeze(-1)).squeeze(-1)
    if value is None:
        value = torch.zeros(obs.shape[0], dtype=torch.float32)
    return actions, logp, value, logits


def batch_normalize(x, eps=1e-8):
    """utils/torch.py:97-99 (torch.std is unbiased)."""
    return (x - x.mean()) / (x.std() + eps)


def kl_diagnostics(old_logp, new_logp):
    """utils/torch.py:102-119."""
    d = torch.clamp(new_logp - old_logp, min=-20.0, max=20.0)
    ratio = torch.exp(d)
    return (old_logp - new_logp).mean(), ((ratio - 1) - torch.log(ratio)).mean()


def activation_stats(pre):
    """utils/models.py:121-146 on the hooked Linear outputs (pre-activation)."""
    out = {}
    for name, z in zip(("backbone.0", "backbone.2"), pre):
        flat = z.detach().flatten(start_dim=1)
        dead = (flat.abs() < 1e-6).float().mean(dim=0)
        out[name] = dict(mean=flat.mean().item(), std=flat.std().item(), dead_pct=dead.mean().item(), dead_max=dead.max().item())
    return out


def ppo_loss(p, obs, actions, old_logp, values_old, adv, ret, *, clip_range, clip_range_vf, vf_coef, ent_coef,
             normalize_adv=True, activation="relu"):
    """PPOAgent.losses_for_batch — agents/ppo/ppo_agent.py:21-152.  Returns (loss, metrics dict of tensors)."""
    m = {}
    if normalize_adv:
        adv = batch_normalize(adv)
        m["roll/adv/norm/mean"], m["roll/adv/norm/std"] = adv.mean(), adv.std()
    logits, v, pre = forward(p, obs, activation, return_pre=True)
    logp_all = logits - logits.logsumexp(dim=-1, keepdim=True)
    new_logp = logp_all.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1)
    ratio = torch.exp(new_logp - old_logp)
    s1 = adv * ratio
    s2 = adv * torch.clamp(ratio, 1.0 - clip_range, 1.0 + clip_range)
    policy_loss = -torch.min(s1, s2).mean()
    vd = v - values_old
    lu = (v - ret) ** 2
    vc = values_old + torch.clamp(vd, -clip_range_vf, clip_range_vf)
    lc = (vc - ret) ** 2
    value_loss = torch.max(lu, lc).mean()
    entropy = categorical_entropy(logits).mean()
    entropy_loss = -entropy
    loss = policy_loss + vf_coef * value_loss + ent_coef * entropy_loss
    with torch.no_grad():
        m["opt/ppo/clip_fraction"] = ((ratio < 1.0 - clip_range) | (ratio > 1.0 + clip_range)).float().mean()
        m["opt/ppo/clip_fraction_vf"] = ((vd < -clip_range_vf) | (vd > clip_range_vf)).float().mean()
        m["opt/value/explained_var"] = 1 - torch.var(ret - v) / torch.var(ret)
        m["opt/ppo/kl"], m["opt/ppo/approx_kl"] = kl_diagnostics(old_logp, new_logp)
    m.update({
        "opt/loss/total": loss.detach(), "opt/loss/policy": policy_loss.detach(),
        "opt/loss/entropy": entropy_loss.detach(), "opt/policy/entropy": entropy.detach(),
        "opt/loss/entropy_scaled": (ent_coef * entropy_loss).detach(), "opt/loss/value": value_loss.detach(),
        "opt/loss/value_scaled": (vf_coef * value_loss).detach(),
    })
    m["_activations"] = activation_stats(pre)
    return loss, m


def reinforce_loss(p, obs, actions, old_logp, adv, ret, *, ent_coef, policy_targets="returns",
                   normalize_returns=False, normalize_adv=False, activation="relu"):
    """REINFORCEAgent.losses_for_batch — agents/reinforce/reinforce_agent.py:11-88 (with the missing
    config.normalize_advantages field supplied)."""
    m = {}
    if normalize_returns:
        ret = batch_normalize(ret)
        m["roll/return/norm/mean"], m["roll/return/norm/std"] = ret.mean(), ret.std()
    if normalize_adv:
        adv = batch_normalize(adv)
        m["roll/adv/norm/mean"], m["roll/adv/norm/std"] = adv.mean(), adv.std()
    targets = ret if policy_targets == "returns" else adv
    logits, _, pre = forward(p, obs, activation, return_pre=True)
    logp_all = logits - logits.logsumexp(dim=-1, keepdim=True)
    logp = logp_all.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1)
    policy_loss = -(logp * targets).mean()
    entropy = categorical_entropy(logits).mean()
    entropy_loss = -entropy
    loss = policy_loss + ent_coef * entropy_loss
    with torch.no_grad():
        m["opt/ppo/kl"], m["opt/ppo/approx_kl"] = kl_diagnostics(old_logp, logp)
    m.update({
        "opt/loss/total": loss.detach(), "opt/loss/policy": policy_loss.detach(),
        "opt/loss/entropy": entropy_loss.detach(), "opt/policy/entropy": entropy.detach(),
        "policy_targets_mean": targets.mean().detach(), "policy_targets_std": targets.std().detach(),
    })
    m["_activations"] = activation_stats(pre)
    return loss, m


def loss_and_grads(loss_fn, p, *args, **kw):
    """Autograd gradients in PARAM_ORDER as one flat vector (the layout of gs_ppo_step's grads_flat)."""
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss, metrics = loss_fn(q, *args, **kw)
    loss.backward()
    flat = torch.cat([q[k].grad.reshape(-1) if q[k].grad is not None else torch.zeros(q[k].numel()) for k in PARAM_ORDER if k in q])
    return loss.detach(), flat, metrics


def grad_norms_and_clip(p, flat, max_norm):
    """BaseModel.compute_grad_norms (utils/models.py:196-230) then torch clip_grad_norm_ semantics
    (scale = max_norm / (total + 1e-6), clamped to 1)."""
    sizes = [(k, p[k].numel()) for k in PARAM_ORDER if k in p]
    off, sq = 0, {}
    for k, n in sizes:
        sq[k] = float((flat[off:off + n].double() ** 2).sum())
        off += n
    norms = {
        "all": math.sqrt(sum(sq.values())),
        "backbone": math.sqrt(sum(v for k, v in sq.items() if k in ("w1", "b1", "w2", "b2"))),
        "policy_head": math.sqrt(sq["wp"] + sq["bp"]),
        "value_head": math.sqrt(sq.get("wv", 0.0) + sq.get("bv", 0.0)),
    }
    coef = min(1.0, max_norm / (norms["all"] + 1e-6)) if max_norm and max_norm > 0 else 1.0
    return norms, coef, flat * coef
