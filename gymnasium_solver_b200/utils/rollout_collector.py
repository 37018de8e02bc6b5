"""RolloutCollector on the CUDA engine (reference: utils/rollout_collector.py:22-777).

Same constructor, ``collect`` / ``evaluate_episodes`` / ``slice_trajectories`` / ``get_metrics`` / ``pop_recent_episodes`` /
``get_action_histogram_counts`` and counters as the reference, but one ``collect()`` is: ONE fused collect launch
(policy forward + sample + env step + buffer write for all ``n_steps``), one target kernel (GAE or MC), and a handful
of reductions — everything stays in HBM.  Running statistics are (sum, sumsq, count) accumulators on device and episode
bookkeeping is resolved lazily, so ``collect()`` itself never synchronises with the host.
"""
from __future__ import annotations

import ctypes as C
import threading
import time
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _native as N
from ..envs.device_vec_env import DeviceVecEnv
from .returns_advantages import moments_into, valid_mask_and_index_map_from_last_terminal
from .rollout_buffer import RolloutBuffer, RolloutTrajectory, _env_major
from .rollout_stats import RollingWindow, RunningStats
from .torch import _device_of, inference_ctx

_ALIASES = {"episode": "mc:episode", "reward_to_go": "mc:rtg", "rtg": "mc:rtg", "mc_episode": "mc:episode", "mc_rtg": "mc:rtg",
            "gae_rtg": "gae:rtg", "gae": "gae", "baseline": "baseline"}
_STAT_NAMES = ("obs", "rew", "base", "adv", "adv_norm", "ret")


class DeviceTrajectory:
    """What ``collect()`` returns: quacks like the reference's RolloutTrajectory NamedTuple (same 9 fields, env-major
    ``(N*T, ...)`` tensors, materialised lazily and cached) and carries the time-major views the update kernels read in
    place (``tm``)."""

    _fields = RolloutTrajectory._fields

    def __init__(self, tm: dict, T: int, n_envs: int):
        self.tm, self.T, self.n_envs = tm, int(T), int(n_envs)
        self._cache = {}

    def _get(self, name):
        if name not in self._cache:
            src = {"observations": "obs", "actions": "actions", "rewards": "rewards", "dones": "dones", "logprobs": "logprobs",
                   "values": "values", "advantages": "adv", "returns": "ret", "next_observations": "next_obs"}[name]
            x = self.tm[src]
            if x is None:
                # not stored: with NEXT_STEP autoreset the observation returned by step t IS the policy input of step
                # t+1, and the one after the last step is last_obs — rebuild it on demand, bit-identical
                x = torch.cat([self.tm["obs"][1:], self.tm["last_obs"][None]], dim=0)
            y = _env_major(x)
            if name == "actions":
                y = y.to(torch.int64)
            elif name == "dones":
                y = y.bool()
            self._cache[name] = y
        return self._cache[name]

    def __getattr__(self, name):
        if name in RolloutTrajectory._fields:
            return self._get(name)
        raise AttributeError(name)

    def __len__(self):
        return self.T * self.n_envs

    def __iter__(self):
        return (self._get(f) for f in self._fields)

    def as_namedtuple(self) -> RolloutTrajectory:
        return RolloutTrajectory(*self)


class RolloutCollector:
    def __init__(self, env, policy_model, n_steps, *, stats_window_size=100, gamma: float = 0.99, gae_lambda: float = 0.95,
                 returns_type: Optional[str] = None, normalize_returns: bool = False, advantages_type: Optional[str] = None,
                 normalize_advantages: bool = False, buffer_maxsize: Optional[int] = None,
                 mc_treat_timeouts_as_terminals: bool = True, use_gae: Optional[bool] = None, **kwargs) -> None:
        if not isinstance(env, DeviceVecEnv):
            raise N.EngineError("the b200 RolloutCollector drives DeviceVecEnv handles only (CartPole-v1, Acrobot-v1, "
                                "MountainCar-v0); host Gymnasium envs belong to the reference collector")
        self.env, self.policy_model, self.n_steps = env, policy_model, int(n_steps)
        self.stats_window_size, self.gamma, self.gae_lambda = stats_window_size, gamma, gae_lambda

        def _s(x):
            return None if x is None else str(getattr(x, "value", x))

        rtype, atype = _s(returns_type), _s(advantages_type)
        if rtype is None or atype is None:
            if bool(use_gae):
                rtype, atype = rtype or "gae:rtg", atype or "gae"
            else:
                rtype, atype = rtype or "mc:rtg", atype or "baseline"
        self.returns_type, self.advantages_type = _ALIASES.get(rtype, rtype), _ALIASES.get(atype, atype)
        self.normalize_returns, self.normalize_advantages = normalize_returns, normalize_advantages
        self.mc_treat_timeouts_as_terminals = mc_treat_timeouts_as_terminals
        self.kwargs = kwargs
        self.buffer_maxsize = buffer_maxsize
        self.store_next_obs = bool(kwargs.get("store_next_obs", True))
        self.rng_seed = int(kwargs.get("rng_seed", env.seed)) & (2**63 - 1)

        self.device = _device_of(policy_model)
        self.n_envs = env.num_envs

        self.rollout_fpss = RollingWindow(stats_window_size)
        self.episode_reward_deque = RollingWindow(stats_window_size)
        self.episode_length_deque = RollingWindow(stats_window_size)
        # per-env windows exist on the reference but nothing reads them; kept for small vectors only
        self._per_env_windows = self.n_envs <= 1024
        self.env_episode_reward_deques = [RollingWindow(stats_window_size) for _ in range(self.n_envs)] if self._per_env_windows else []
        self.env_episode_length_deques = [RollingWindow(stats_window_size) for _ in range(self.n_envs)] if self._per_env_windows else []
        self._last_episode_reward, self._last_episode_length = 0.0, 0
        self._best_episode_reward = -float("inf")

        # running statistics: host mirrors + device accumulators [6][3] = (sum, sumsq, count)
        self._stats = {name: RunningStats() for name in _STAT_NAMES}
        self._obs_stats, self._rew_stats, self._base_stats = self._stats["obs"], self._stats["rew"], self._stats["base"]
        self._adv_stats, self._adv_norm_stats, self._ret_stats = self._stats["adv"], self._stats["adv_norm"], self._stats["ret"]
        self._stats_dev = None
        self._stats_synced = None
        self._action_counts = None
        self._action_counts_dev = None
        self._supports_action_probs = None

        self.total_rollouts = self.total_steps = self.total_vec_steps = self.total_episodes = 0
        self.rollout_steps = self.rollout_vec_steps = self.rollout_episodes = 0

        self.obs = None
        self._buffer: Optional[RolloutBuffer] = None
        self._last_rollout_index_map = None
        self._last_rollout_valid_count = None
        self.terminal_obs_info = []
        self._recent_episodes: List[Tuple[int, float, int, bool]] = []
        self._pending = None       # (start, end) of a rollout whose episode bookkeeping has not been resolved yet
        self._events = None

    # ------------------------------------------------------------------------------------------------ helpers
    def _sync_device_and_prepare_buffers(self) -> None:
        if self.obs is not None:
            return
        self.device = _device_of(self.policy_model)
        if self.device.type != "cuda":
            raise N.EngineError("the policy model must live on a CUDA device")
        if self.device != self.env.device:
            raise N.EngineError(f"model on {self.device} but environments on {self.env.device}")
        self.obs, _ = self.env.reset()
        if self._buffer is None:
            maxsize = self.buffer_maxsize if self.buffer_maxsize is not None else self.n_steps
            self._buffer = RolloutBuffer(self.n_envs, (self.env.obs_dim,), np.float32, self.device, maxsize,
                                         store_next_obs=self.store_next_obs)
            n = self.n_envs
            self._last_values = torch.zeros(n, dtype=torch.float32, device=self.device)
            self._last_obs = torch.zeros(n, self.env.obs_dim, dtype=torch.float32, device=self.device)
            self._adv = torch.zeros(maxsize, n, dtype=torch.float32, device=self.device)
            self._ret = torch.zeros(maxsize, n, dtype=torch.float32, device=self.device)
            self._last_terminal = torch.zeros(n, dtype=torch.int32, device=self.device)
            self._stats_dev = torch.zeros(len(_STAT_NAMES), 3, dtype=torch.float64, device=self.device)
            self._scratch_mom = torch.zeros(2, 3, dtype=torch.float64, device=self.device)
            self._stats_synced = np.zeros((len(_STAT_NAMES), 3))
            self._action_counts_dev = torch.zeros(self.env.n_actions, dtype=torch.int64, device=self.device)

    def _rollout_struct(self, start: int, end: int) -> N.GsRollout:
        b, r = self._buffer, N.GsRollout()
        r.T, r.obs_dim, r.N = end - start, self.env.obs_dim, self.n_envs
        sl = slice(start, end)
        r.obs, r.actions, r.logprobs = N.ptr(b.obs_buf[sl]), N.ptr(b.actions_buf[sl]), N.ptr(b.logprobs_buf[sl])
        r.next_obs = N.ptr(b.next_obs_buf[sl]) if b.next_obs_buf is not None else None
        r.values, r.rewards = N.ptr(b.values_buf[sl]), N.ptr(b.rewards_buf[sl])
        r.dones, r.timeouts = N.ptr(b.dones_buf[sl]), N.ptr(b.timeouts_buf[sl])
        r.last_obs, r.last_values = N.ptr(self._last_obs), N.ptr(self._last_values)
        r.ep_return, r.ep_length = N.ptr(b.ep_return_buf[sl]), N.ptr(b.ep_length_buf[sl])
        return r

    def _stat_row(self, name: str) -> torch.Tensor:
        return self._stats_dev[_STAT_NAMES.index(name)]

    def _flush_stats(self) -> None:
        """Bring the device accumulators into the host RunningStats mirrors (one small D2H copy)."""
        if self._stats_dev is None:
            return
        cur = self._stats_dev.cpu().numpy()
        delta = cur - self._stats_synced
        for i, name in enumerate(_STAT_NAMES):
            self._stats[name].update_moments(delta[i, 0], delta[i, 1], int(round(delta[i, 2])))
        self._stats_synced = cur
        counts = self._action_counts_dev.cpu().numpy()
        self._action_counts = counts.copy() if counts.sum() > 0 else self._action_counts

    def resolve_episodes_async(self) -> None:
        """Episode bookkeeping of the last rollout on a SIDE stream that only waits for the collect / target kernels.  Called by the
        agent after it has queued the rollout's minibatch passes: the host synchronisation inside (``nonzero`` needs a count) then
        returns while the training stream still has the passes to run, instead of draining it at the next ``collect()``."""
        if self._pending is None or not self._events_pending:
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._side_stream):
            self._side_stream.wait_event(self._events[1])        # recorded after the rollout's last kernel (targets)
            self._resolve_pending_episodes()
            self._side_stream.synchronize()

    _side_stream = None

    def _resolve_pending_episodes(self) -> None:
        """Episode bookkeeping of the last rollout (reference: _process_done_infos, rollout_collector.py:210-294),
        resolved lazily from the (T, N) ``dones`` / ``ep_return`` / ``ep_length`` arrays the collect kernel wrote."""
        if self._pending is None:
            return
        start, end = self._pending
        self._pending = None
        b = self._buffer
        dones = b.dones_buf[start:end]
        flat = dones.reshape(-1)
        pos = torch.nonzero(flat, as_tuple=False).squeeze(-1)     # row-major == (step, env) order of the reference loop
        n_done = int(pos.numel())
        self.rollout_episodes = n_done
        self.total_episodes += n_done
        if n_done == 0:
            return
        ep_r = b.ep_return_buf[start:end].reshape(-1)[pos]
        ep_l = b.ep_length_buf[start:end].reshape(-1)[pos]
        self._best_episode_reward = max(self._best_episode_reward, float(ep_r.max().item()))
        k = self.stats_window_size
        tail_r, tail_l = ep_r[-k:].cpu().numpy(), ep_l[-k:].cpu().numpy()
        self.episode_reward_deque.extend(tail_r.tolist())
        self.episode_length_deque.extend(tail_l.tolist())
        self._last_episode_reward, self._last_episode_length = float(tail_r[-1]), int(tail_l[-1])
        self._recent_dev = (pos, ep_r, ep_l, b.timeouts_buf[start:end].reshape(-1)[pos].bool())
        if self._per_env_windows:
            env_idx = (pos % self.n_envs).cpu().numpy()
            all_r, all_l = ep_r.cpu().numpy(), ep_l.cpu().numpy()
            for e, r_, l_ in zip(env_idx.tolist(), all_r.tolist(), all_l.tolist()):
                self.env_episode_reward_deques[e].append(r_)
                self.env_episode_length_deques[e].append(l_)

    # ------------------------------------------------------------------------------------------------ collect
    @torch.no_grad()
    def collect(self, *args, **kwargs):
        """One rollout slice with the current policy (reference: rollout_collector.py:296-304, 459-567)."""
        with inference_ctx(self.policy_model):
            return self._collect(*args, **kwargs)

    def _collect(self, deterministic: bool = False) -> DeviceTrajectory:
        self._sync_device_and_prepare_buffers()
        self._resolve_pending_episodes()          # before the buffer slice is overwritten
        start = self._buffer.begin_rollout(self.n_steps)
        end = start + self.n_steps
        self.rollout_steps = self.rollout_vec_steps = self.rollout_episodes = 0

        if self._events is None:
            self._events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        elif self._events_pending:
            # fps of the PREVIOUS rollout, measured on device without stalling this one
            self._events[1].synchronize()
            ms = self._events[0].elapsed_time(self._events[1])
            if ms > 0:
                self.rollout_fpss.append(self.n_envs * self.n_steps / (ms * 1e-3))
        wall0 = time.time()
        mlp = N.mlp_struct(self.policy_model)
        roll = self._rollout_struct(start, end)
        with torch.cuda.device(self.device):
            self._events[0].record()
            N.check(N.lib().gs_rollout_collect(self.env.handle, C.byref(mlp), C.byref(roll), N.ptr(self.obs), self.rng_seed,
                                               self.total_vec_steps, int(deterministic), N.stream()))
        self.rollout_steps = self.n_envs * self.n_steps
        self.rollout_vec_steps = self.n_steps
        self.total_steps += self.rollout_steps
        self.total_vec_steps += self.rollout_vec_steps
        self._pending = (start, end)

        self._update_running_stats_after_rollout(start, end)
        adv, ret = self._compute_targets(start, end)
        with torch.cuda.device(self.device):
            self._events[1].record()
        self._events_pending = True
        self._wall_last = time.time() - wall0

        b = self._buffer
        sl = slice(start, end)
        tm = dict(obs=b.obs_buf[sl], next_obs=None if b.next_obs_buf is None else b.next_obs_buf[sl], actions=b.actions_buf[sl],
                  rewards=b.rewards_buf[sl], dones=b.dones_buf[sl], timeouts=b.timeouts_buf[sl], logprobs=b.logprobs_buf[sl],
                  values=b.values_buf[sl], adv=adv, ret=ret, idx_map=self._last_rollout_index_map, last_obs=self._last_obs)
        self.total_rollouts += 1
        return DeviceTrajectory(tm, self.n_steps, self.n_envs)

    _events_pending = False

    def _update_running_stats_after_rollout(self, start: int, end: int) -> None:
        b = self._buffer
        moments_into(b.obs_buf[start:end].reshape(end - start, -1), self._stat_row("obs"))
        moments_into(b.rewards_buf[start:end], self._stat_row("rew"))
        a = b.actions_buf[start:end].reshape(-1)
        for k in range(self.env.n_actions):
            self._action_counts_dev[k] += (a == k).sum()

    def _compute_targets(self, start: int, end: int):
        """Advantages / returns for [start, end) (reference: rollout_collector.py:362-457)."""
        b, L = self._buffer, N.lib()
        T, n = end - start, self.n_envs
        sl = slice(start, end)
        adv, ret = self._adv[sl], self._ret[sl]
        valid_lt = n_valid = None
        with torch.cuda.device(self.device):
            st = N.stream()
            if self.returns_type == "gae:rtg" and self.advantages_type == "gae":
                if b._bootstrapped_values_buf is None:       # never materialised == identically zero: the scan does not read it
                    N.check(L.gs_gae_zero_boot(N.ptr(b.values_buf[sl]), N.ptr(b.rewards_buf[sl]), N.ptr(b.dones_buf[sl]), N.ptr(b.timeouts_buf[sl]),
                                               N.ptr(self._last_values), T, n, float(self.gamma), float(self.gae_lambda), N.ptr(adv), N.ptr(ret), st))
                else:
                    N.check(L.gs_gae(N.ptr(b.values_buf[sl]), N.ptr(b.rewards_buf[sl]), N.ptr(b.dones_buf[sl]), N.ptr(b.timeouts_buf[sl]),
                                     N.ptr(self._last_values), N.ptr(b.bootstrapped_values_buf[sl]), T, n, float(self.gamma),
                                     float(self.gae_lambda), N.ptr(adv), N.ptr(ret), st))
                self._last_rollout_index_map = None
            elif self.returns_type in ("mc:episode", "mc:rtg"):
                to = None if self.mc_treat_timeouts_as_terminals else b.timeouts_buf[sl]
                N.check(L.gs_mc_returns(N.ptr(b.rewards_buf[sl]), N.ptr(b.dones_buf[sl]), N.ptr(to), T, n, float(self.gamma),
                                        int(self.returns_type == "mc:episode"), N.ptr(ret), N.ptr(self._last_terminal), st))
                _, imap, n_valid = valid_mask_and_index_map_from_last_terminal(self._last_terminal, T)
                self._last_rollout_index_map, self._last_rollout_valid_count = imap, n_valid
                valid_lt = self._last_terminal
                # global baseline over valid (non-trailing) returns of every rollout so far, then adv = ret - baseline
                moments_into(ret, self._stat_row("base"), valid_lt)
                if self.advantages_type == "baseline":
                    N.check(L.gs_shift_by_mean(N.ptr(ret), ret.numel(), N.ptr(self._stat_row("base")), N.ptr(adv), st))
                else:
                    adv.copy_(ret)
            else:
                raise ValueError(f"Invalid returns_type: {self.returns_type} and advantages_type: {self.advantages_type}")
            if self.normalize_returns:
                mom = self._scratch_mom[0].zero_()
                moments_into(ret, mom)
                N.check(L.gs_normalize(N.ptr(ret), ret.numel(), N.ptr(mom), 1e-8, N.ptr(ret), st))
            # valid positions only; a rollout WITHOUT any real terminal has no mask in the reference and counts in full (n_valid == 0)
            moments_into(adv, self._stat_row("adv"), valid_lt, n_valid)
            if self.normalize_advantages:
                mom = self._scratch_mom[1].zero_()
                moments_into(adv, mom)
                N.check(L.gs_normalize(N.ptr(adv), adv.numel(), N.ptr(mom), 1e-8, N.ptr(adv), st))
                moments_into(adv, self._stat_row("adv_norm"), valid_lt, n_valid)
            moments_into(ret, self._stat_row("ret"), valid_lt, n_valid)
        return adv, ret

    # ------------------------------------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def evaluate_episodes(self, *, n_episodes: int, deterministic: bool = True, timeout_seconds: Optional[float] = None) -> dict:
        """Exactly-N-episodes evaluation with balanced per-env quotas (reference: rollout_collector.py:569-655)."""
        self._busy_thread = threading.get_ident()
        try:
            return self._evaluate_episodes(n_episodes, deterministic, timeout_seconds)
        finally:
            self._busy_thread = None

    def _evaluate_episodes(self, n_episodes: int, deterministic: bool, timeout_seconds: Optional[float]) -> dict:
        n = self.n_envs
        base, rem = int(n_episodes) // n, int(n_episodes) % n
        targets = torch.tensor([base + (1 if i < rem else 0) for i in range(n)], device=self.env.device) if n_episodes > 0 else \
            torch.zeros(n, dtype=torch.int64, device=self.env.device)
        counts = torch.zeros(n, dtype=torch.int64, device=self.env.device)
        reward_sum, length_sum, total_timesteps, total_vec_steps = 0.0, 0, 0, 0
        self._resolve_pending_episodes()
        self.obs = None
        self._recent_episodes, self._recent_dev = [], None
        self._sync_device_and_prepare_buffers()
        t0 = time.time()
        while bool((counts < targets).any()):
            self.collect(deterministic=deterministic)
            total_timesteps += self.n_envs * self.n_steps
            total_vec_steps += self.rollout_vec_steps
            start, end = self._pending
            self._resolve_pending_episodes()
            b = self._buffer
            dones = b.dones_buf[start:end].to(torch.int64)
            rank = counts[None, :] + torch.cumsum(dones, dim=0) - 1       # k-th finished episode of each env
            take = (dones > 0) & (rank < targets[None, :])
            reward_sum += float((b.ep_return_buf[start:end] * take).sum().item())
            length_sum += int((b.ep_length_buf[start:end].to(torch.int64) * take).sum().item())
            counts += take.sum(dim=0)
            if timeout_seconds is not None and (time.time() - t0) >= float(timeout_seconds):
                break
        collected = int(counts.sum().item())
        metrics = self.get_metrics()
        metrics.pop("action_dist")
        metrics.update({"cnt/total_episodes": collected, "cnt/total_env_steps": int(total_timesteps), "cnt/total_vec_steps": int(total_vec_steps)})
        if collected > 0:
            metrics["roll/ep_rew/mean"] = float(reward_sum / collected)
            metrics["roll/ep_len/mean"] = float(length_sum / collected)
        return metrics

    # ------------------------------------------------------------------------------------------------ slicing / metrics
    def slice_trajectories(self, trajectories, idxs):
        """Gather a minibatch by env-major sample ids (reference: rollout_collector.py:657-682); MC index remap included.
        Compatibility surface: the engine's update kernels gather in place from the time-major buffer instead."""
        idx = torch.as_tensor(idxs, dtype=torch.int64, device=self.device)
        if self._last_rollout_index_map is not None and self.advantages_type != "gae":
            idx = self._last_rollout_index_map[idx]
        return RolloutTrajectory(*(getattr(trajectories, f)[idx] for f in RolloutTrajectory._fields))

    _busy_thread = None           # ident of the thread inside evaluate_episodes()
    _metrics_snapshot: dict = {}

    def get_metrics(self) -> dict:
        busy = self._busy_thread
        if busy is not None and busy != threading.get_ident():
            # an evaluation is running on another thread and stream (BaseAgent eval_async): the collector's buffers are in use there,
            # so a reader on this thread gets the view of the last completed call instead of touching them
            return dict(self._metrics_snapshot)
        self._resolve_pending_episodes()
        self._flush_stats()
        if self._events_pending and self._events is not None:
            self._events[1].synchronize()
            ms = self._events[0].elapsed_time(self._events[1])
            if ms > 0:
                self.rollout_fpss.append(self.n_envs * self.n_steps / (ms * 1e-3))
            self._events_pending = False
        if self._action_counts is not None and self._action_counts.sum() > 0:
            ids = np.arange(self._action_counts.shape[0], dtype=np.float32)
            tot = float(self._action_counts.sum())
            a_mean = float((ids * self._action_counts).sum() / tot)
            a_std = float(np.sqrt(max(0.0, float(((ids - a_mean) ** 2 * self._action_counts).sum() / tot))))
            a_dist = self._action_counts.copy()
        else:
            a_mean, a_std, a_dist = 0.0, 0.0, None
        m = {
            "cnt/total_env_steps": self.total_steps, "cnt/total_vec_steps": self.total_vec_steps,
            "cnt/total_episodes": self.total_episodes, "cnt/total_rollouts": self.total_rollouts,
            "roll/env_steps": self.rollout_steps, "roll/vec_steps": self.rollout_vec_steps, "roll/episodes": self.rollout_episodes,
            "roll/fps": float(self.rollout_fpss.mean()) if self.rollout_fpss else 0.0,
            "roll/obs/mean": self._obs_stats.mean(), "roll/obs/std": self._obs_stats.std(),
            "roll/reward/mean": self._rew_stats.mean(), "roll/reward/std": self._rew_stats.std(),
            "roll/return/mean": self._ret_stats.mean(), "roll/return/std": self._ret_stats.std(),
            "roll/adv/mean": self._adv_stats.mean(), "roll/adv/std": self._adv_stats.std(),
            "roll/actions/mean": a_mean, "roll/actions/std": a_std, "action_dist": a_dist,
            "roll/baseline/mean": self._base_stats.mean(), "roll/baseline/std": self._base_stats.std(),
        }
        if self.normalize_advantages and self._adv_norm_stats.count > 0:
            m["roll/adv_norm/mean"], m["roll/adv_norm/std"] = self._adv_norm_stats.mean(), self._adv_norm_stats.std()
        if self.episode_reward_deque:
            m["roll/ep_rew/mean"] = float(self.episode_reward_deque.mean())
            m["roll/ep_len/mean"] = int(self.episode_length_deque.mean())
            m["roll/ep_rew/best"] = float(self._best_episode_reward)
            m["roll/ep_rew/last"] = float(self._last_episode_reward)
            m["roll/ep_len/last"] = int(self._last_episode_length)
        self._metrics_snapshot = dict(m)
        return m

    _recent_dev = None

    def pop_recent_episodes(self) -> List[Tuple[int, float, int, bool]]:
        """(env_idx, reward, length, was_timeout) of the episodes finished in the last rollout."""
        self._resolve_pending_episodes()
        if self._recent_dev is None:
            return []
        pos, ep_r, ep_l, to = self._recent_dev
        self._recent_dev = None
        env_idx = (pos % self.n_envs).cpu().tolist()
        return list(zip(env_idx, ep_r.cpu().tolist(), ep_l.cpu().tolist(), to.cpu().tolist()))

    def get_action_histogram_counts(self, reset: bool = False):
        self._flush_stats()
        if self._action_counts is None:
            return None
        out = self._action_counts.copy()
        if reset:
            self._action_counts = np.zeros_like(out)
            self._action_counts_dev.zero_()
        return out
