"""GAE / Monte-Carlo targets on the CUDA engine (reference: utils/returns_advantages.py).

Same function names and argument meaning as the reference; inputs are time-major ``(T, N)`` arrays (CUDA tensors stay on
device; numpy / CPU inputs are shipped to the current CUDA device and the results come back as numpy).  Arithmetic is
the reference's fp32 sequence, so results are bit-identical to its numpy implementation."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native as N


def _dev(x, dtype, device):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype).contiguous()


def _target_device(*xs) -> Tuple[torch.device, bool]:
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device, True
    if not torch.cuda.is_available():
        raise N.EngineError("returns/advantages run on the CUDA engine; no CUDA device is available")
    return torch.device("cuda", torch.cuda.current_device()), False


def _out(t: torch.Tensor, keep_on_device: bool):
    return t if keep_on_device else t.cpu().numpy()


def compute_batched_gae_advantages_and_returns(values, rewards, dones, timeouts, last_values, bootstrapped_next_values,
                                               gamma: float, gae_lambda: float, *, out=None):
    """returns_advantages.py:115-155 — (advantages, returns), both (T, N) fp32."""
    dev, on_dev = _target_device(values, rewards)
    v, r = _dev(values, torch.float32, dev), _dev(rewards, torch.float32, dev)
    d, to = _dev(dones, torch.uint8, dev), _dev(timeouts, torch.uint8, dev)
    lv = _dev(last_values, torch.float32, dev)
    b = None if bootstrapped_next_values is None else _dev(bootstrapped_next_values, torch.float32, dev)
    T, n = r.shape
    adv, ret = out if out is not None else (torch.empty_like(r), torch.empty_like(r))
    with torch.cuda.device(dev):
        N.check(N.lib().gs_gae(N.ptr(v), N.ptr(r), N.ptr(d), N.ptr(to), N.ptr(lv), N.ptr(b), T, n, float(gamma), float(gae_lambda),
                               N.ptr(adv), N.ptr(ret), N.stream()))
    return _out(adv, on_dev), _out(ret, on_dev)


def compute_batched_mc_returns(rewards, dones, timeouts, gamma: float, *, episode_mode: bool = False, return_last_terminal: bool = False):
    """returns_advantages.py:67-91 (+ :93-113 when ``episode_mode``).  ``timeouts=None`` == all False."""
    dev, on_dev = _target_device(rewards)
    r, d = _dev(rewards, torch.float32, dev), _dev(dones, torch.uint8, dev)
    to = None if timeouts is None else _dev(timeouts, torch.uint8, dev)
    T, n = r.shape
    ret = torch.empty_like(r)
    lt = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().gs_mc_returns(N.ptr(r), N.ptr(d), N.ptr(to), T, n, float(gamma), int(episode_mode), N.ptr(ret), N.ptr(lt), N.stream()))
    if return_last_terminal:
        return _out(ret, on_dev), _out(lt, on_dev)
    return _out(ret, on_dev)


def convert_returns_to_full_episode(returns, dones, timeouts):
    """returns_advantages.py:93-113 — every step takes the reward-to-go of its segment's first step (in place on device
    tensors, like the reference mutates its argument)."""
    dev, on_dev = _target_device(returns)
    ret = _dev(returns, torch.float32, dev)
    if not on_dev:
        ret = ret.clone()
    d = _dev(dones, torch.uint8, dev)
    to = None if timeouts is None else _dev(timeouts, torch.uint8, dev)
    T, n = ret.shape
    with torch.cuda.device(dev):
        N.check(N.lib().gs_returns_to_full_episode(N.ptr(ret), N.ptr(d), N.ptr(to), T, n, N.stream()))
    return _out(ret, on_dev)


def _build_valid_mask_and_index_map(dones, timeouts):
    """returns_advantages.py:33-52 — env-major (N*T,) valid mask and nearest-previous-valid index map, or (None, None)."""
    dev, on_dev = _target_device(dones)
    d = _dev(dones, torch.uint8, dev)
    if d.numel() == 0:
        return None, None
    to = None if timeouts is None else _dev(timeouts, torch.uint8, dev)
    T, n = d.shape
    zeros = torch.zeros(T, n, dtype=torch.float32, device=dev)
    _, lt = compute_batched_mc_returns(zeros, d, to, 1.0, return_last_terminal=True)
    mask, imap, n_valid = valid_mask_and_index_map_from_last_terminal(lt, T)
    if int(n_valid.item()) == 0:
        return None, None
    return _out(mask.bool(), on_dev), _out(imap, on_dev)


def valid_mask_and_index_map_from_last_terminal(last_terminal: torch.Tensor, T: int):
    """Device triple (mask uint8 (N*T,), idx_map int64 (N*T,), n_valid int64[1]) from the per-env last real terminal."""
    dev = last_terminal.device
    n = last_terminal.numel()
    mask = torch.empty(n * T, dtype=torch.uint8, device=dev)
    imap = torch.empty(n * T, dtype=torch.int64, device=dev)
    n_valid = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        wsb = N.lib().gs_valid_index_map_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        N.check(N.lib().gs_valid_index_map(N.ptr(last_terminal), T, n, N.ptr(mask), N.ptr(imap), N.ptr(n_valid), N.ptr(ws), wsb, N.stream()))
    return mask, imap, n_valid


def moments_into(x: torch.Tensor, out: torch.Tensor, last_terminal: Optional[torch.Tensor] = None, n_valid: Optional[torch.Tensor] = None) -> None:
    """out[0:3] += (sum, sumsq, count) of a (T, N) device array (optionally only t <= last_terminal[n]).  With ``n_valid`` (the
    device count gs_valid_index_map wrote) a rollout without any valid entry is counted in full, like the reference's None mask
    (rollout_collector.py:435-455)."""
    T = x.shape[0]
    n = x.numel() // T
    with torch.cuda.device(x.device):
        if last_terminal is not None and n_valid is not None:
            N.check(N.lib().gs_moments_valid(N.ptr(x), N.ptr(last_terminal), N.ptr(n_valid), T, n, N.ptr(out), N.stream()))
        else:
            N.check(N.lib().gs_moments(N.ptr(x), N.ptr(last_terminal), T, n, N.ptr(out), N.stream()))


def _normalize(x, eps: float = 1e-8):
    dev, on_dev = _target_device(x)
    xt = _dev(x, torch.float32, dev)
    mom = torch.zeros(3, dtype=torch.float64, device=dev)
    flat = xt.reshape(1, -1)
    moments_into(flat, mom)
    y = torch.empty_like(xt)
    with torch.cuda.device(dev):
        N.check(N.lib().gs_normalize(N.ptr(xt), xt.numel(), N.ptr(mom), float(eps), N.ptr(y), N.stream()))
    return _out(y, on_dev)


def _normalize_returns(returns, eps: float = 1e-8):
    """returns_advantages.py:55-58 (population std)."""
    return _normalize(returns, eps)


def _normalize_advantages(advantages, eps: float = 1e-8):
    """returns_advantages.py:61-64 (population std)."""
    return _normalize(advantages, eps)
