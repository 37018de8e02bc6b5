"""Device rollout storage (reference: utils/rollout_buffer.py:16-173).

Same constructor, ``begin_rollout`` / ``add`` / ``flatten_slice_env_major`` and public ``*_buf`` arrays as the reference,
but the arrays are torch tensors on ``device`` (HBM when it is a CUDA device) laid out time-major ``(maxsize, n_envs, ...)``
— exactly the layout the fused collect kernel writes and the GAE / update kernels read, so no copy or transpose sits
between collection and training.  Actions are int32 on device; ``RolloutTrajectory.actions`` is int64 like the reference.
"""
from __future__ import annotations

from typing import NamedTuple, Tuple

import numpy as np
import torch


class RolloutTrajectory(NamedTuple):
    observations: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    dones: torch.Tensor
    logprobs: torch.Tensor
    values: torch.Tensor
    advantages: torch.Tensor
    returns: torch.Tensor
    next_observations: torch.Tensor


_TORCH_OF_NP = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int64): torch.int64,
                np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8, np.dtype(bool): torch.bool}


def _torch_dtype(dt) -> torch.dtype:
    return dt if isinstance(dt, torch.dtype) else _TORCH_OF_NP[np.dtype(dt)]


def _env_major(x: torch.Tensor) -> torch.Tensor:
    """(T, N, ...) -> (N*T, ...) with env-major order (sample id = n*T + t)."""
    return x.transpose(0, 1).reshape(x.shape[0] * x.shape[1], *x.shape[2:])


class RolloutBuffer:
    def __init__(self, n_envs: int, obs_shape: Tuple[int, ...], obs_dtype, device, maxsize: int,
                 action_shape: Tuple[int, ...] = (), action_dtype=np.int64, store_next_obs: bool = True) -> None:
        self.n_envs, self.obs_shape, self.obs_dtype = int(n_envs), tuple(obs_shape), obs_dtype
        self.device = torch.device(device)
        self.maxsize = int(maxsize)
        self.action_shape, self.action_dtype = tuple(action_shape), action_dtype
        z = lambda *shape, dt: torch.zeros(self.maxsize, self.n_envs, *shape, dtype=dt, device=self.device)
        self.obs_buf = z(*self.obs_shape, dt=_torch_dtype(obs_dtype))
        self.next_obs_buf = z(*self.obs_shape, dt=_torch_dtype(obs_dtype)) if store_next_obs else None
        discrete = len(self.action_shape) == 0
        self.actions_buf = z(*self.action_shape, dt=torch.int32 if discrete else torch.float32)
        self.rewards_buf = z(dt=torch.float32)
        self.values_buf = z(dt=torch.float32)
        self.dones_buf = z(dt=torch.uint8)
        self.logprobs_buf = z(dt=torch.float32)
        self.timeouts_buf = z(dt=torch.uint8)
        # The reference's `bootstrapped_values_buf` (values of final observations at truncations).  The device vector envs autoreset on
        # the NEXT step and emit no final observation, so nothing ever writes it: it is materialised (zeros) only when somebody
        # asks for it, and until then the GAE scan runs its zero-bootstrap variant without reading 4 bytes per element.
        self._bootstrapped_values_buf = None
        # episode accounting written by the collect kernel (RecordEpisodeStatistics values where done)
        self.ep_return_buf = z(dt=torch.float64)
        self.ep_length_buf = z(dt=torch.int32)
        self.pos = 0
        self.size = 0

    @property
    def bootstrapped_values_buf(self) -> torch.Tensor:
        if self._bootstrapped_values_buf is None:
            self._bootstrapped_values_buf = torch.zeros(self.maxsize, self.n_envs, dtype=torch.float32, device=self.device)
        return self._bootstrapped_values_buf

    @bootstrapped_values_buf.setter
    def bootstrapped_values_buf(self, value) -> None:
        self._bootstrapped_values_buf = value

    def begin_rollout(self, T: int) -> int:
        """Reserve T contiguous steps (wrapping to 0 when the tail is too short); returns the start index."""
        if T > self.maxsize:
            raise ValueError(f"Rollout length T={T} exceeds buffer maxsize={self.maxsize}")
        if self.pos + T > self.maxsize:
            self.pos = 0
        start = self.pos
        self.pos += T
        self.size = max(self.size, self.pos)
        return start

    def add(self, idx: int, obs_np, next_obs_np, actions_np, logps_np, values_np, rewards_np, dones_np, timeouts_np) -> None:
        """Per-step write, the reference's parameter names (utils/rollout_buffer.py:82-93; its tests pass them by keyword).  Protocol
        compatibility: the fused collect kernel writes the same arrays directly.  numpy arrays or tensors on any device."""
        obs_t = torch.as_tensor(obs_np)
        assert tuple(obs_t.shape) == (self.n_envs, *self.obs_shape), \
            f"Expected shape {(self.n_envs, *self.obs_shape)}, got {tuple(obs_t.shape)}"
        put = lambda buf, v: buf[idx].copy_(torch.as_tensor(v).to(buf.dtype))
        put(self.obs_buf, obs_t)
        if self.next_obs_buf is not None:
            put(self.next_obs_buf, next_obs_np)
        put(self.actions_buf, actions_np)
        put(self.logprobs_buf, logps_np)
        put(self.values_buf, values_np)
        put(self.rewards_buf, rewards_np)
        put(self.dones_buf, dones_np)
        put(self.timeouts_buf, timeouts_np)

    def flatten_slice_env_major(self, start: int, end: int, advantages_buf, returns_buf) -> RolloutTrajectory:
        """Materialise the reference's env-major (N*T, ...) training tensors for [start, end).

        Compatibility surface only: the engine's update kernels read the time-major arrays in place and translate
        env-major sample ids themselves, so the training fast path never calls this."""
        T, n = end - start, self.n_envs
        sl = slice(start, end)
        obs = _env_major(self.obs_buf[sl])
        if len(self.obs_shape) == 0:
            obs = obs.reshape(n * T, 1)
        nxt_src = self.next_obs_buf if self.next_obs_buf is not None else self.obs_buf
        nxt = _env_major(nxt_src[sl])
        if len(self.obs_shape) == 0:
            nxt = nxt.reshape(n * T, 1)
        act = _env_major(self.actions_buf[sl])
        act = act.to(torch.int64) if len(self.action_shape) == 0 else act.to(torch.float32)
        as_t = lambda x: torch.as_tensor(x, device=self.device)
        return RolloutTrajectory(
            observations=obs, actions=act, rewards=_env_major(self.rewards_buf[sl]), dones=_env_major(self.dones_buf[sl]).bool(),
            logprobs=_env_major(self.logprobs_buf[sl]), values=_env_major(self.values_buf[sl]),
            advantages=_env_major(as_t(advantages_buf).to(torch.float32)), returns=_env_major(as_t(returns_buf).to(torch.float32)),
            next_observations=nxt)
