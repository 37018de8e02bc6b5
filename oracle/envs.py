"""ctypes front-end of oracle/envs.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Exposes the Gymnasium VectorEnv protocol the reference's collector consumes
(utils/rollout_collector.py:317, :504, :223-240): ``num_envs``, ``reset() -> (obs, info)``,
``step(actions) -> (obs, rewards, terminated, truncated, infos)`` with
``infos["episode"]["r"|"l"]`` / ``infos["_episode"]`` from RecordEpisodeStatistics.
Physics parity is UNPINNED (see the header of envs.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_envs.so")

ENV_KINDS = {"CartPole-v1": 0, "Acrobot-v1": 1, "MountainCar-v0": 2}
OBS_DIM = {0: 4, 1: 6, 2: 2}
STATE_DIM = {0: 4, 1: 4, 2: 2}
N_ACTIONS = {0: 2, 1: 3, 2: 3}
WRAPPERS = {"MountainCarV0_StateCountBonus": 1, "CartPoleV1_RewardShaper": 2, "MountainCarV0_RewardShaper": 3}

_lib = None


def build(force: bool = False) -> str:
    """Compile envs.c with the committed Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "envs.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.orc_env_create.restype = vp
        L.orc_env_create.argtypes = [i32, i64, i64, u64, i32]
        L.orc_env_destroy.argtypes = [vp]
        L.orc_env_attach_wrapper.restype = i32
        L.orc_env_attach_wrapper.argtypes = [vp, i32, vp, i32]
        L.orc_env_set_state.argtypes = [vp, vp, vp]
        L.orc_env_get_state.argtypes = [vp, vp, vp]
        L.orc_env_counts.restype = vp
        L.orc_env_counts.argtypes = [vp]
        L.orc_env_reset.argtypes = [vp, vp]
        L.orc_env_step.argtypes = [vp] + [vp] * 7
        L.orc_env_step_range.argtypes = [vp, i64, i64] + [vp] * 7
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_reset_uniforms.argtypes = [u64, u64, C.c_uint32, i32, vp]
        _lib = L
    return _lib


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def reset_uniforms(seed: int, gid: int, reset_idx: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float64)
    lib().orc_reset_uniforms(seed, gid, reset_idx, n, out.ctypes.data)
    return out


def wrapper_params(spec: dict):
    """(kind, params) for an EnvWrapperRegistry spec dict {"id": ..., **kwargs} using the wrapper defaults."""
    wid = spec["id"]
    kw = {k: v for k, v in spec.items() if k != "id"}
    if wid == "MountainCarV0_StateCountBonus":  # state_count_bonus.py:13-21
        btype = {"count": 0, "inverse": 1, "log": 2}[kw.get("bonus_type", "count")]
        p = [kw.get("position_bins", 50), kw.get("velocity_bins", 50), kw.get("bonus_scale", 1.0), btype, kw.get("min_count", 1)]
    elif wid == "CartPoleV1_RewardShaper":  # reward_shaper.py:24-30
        p = [kw.get("angle_reward_scale", 1.0), kw.get("position_reward_scale", 0.25), 1.0 if kw.get("clip_potential", True) else 0.0]
    elif wid == "MountainCarV0_RewardShaper":  # reward_shaper.py:9
        p = [kw.get("position_reward_scale", 100.0), kw.get("velocity_reward_scale", 10.0), kw.get("height_reward_scale", 50.0)]
    else:
        raise KeyError(wid)
    return WRAPPERS[wid], np.asarray(p, dtype=np.float64)


class OracleVecEnv:
    """SyncVectorEnv(+TimeLimit, +RecordEpisodeStatistics) stand-in backed by the C restatement."""

    def __init__(self, env_id: str, n_envs: int, seed: int = 0, *, env_id_offset: int = 0,
                 max_episode_steps: int | None = None, env_wrappers=(), threads: int = 1):
        self.kind = ENV_KINDS[env_id]
        self.env_id = env_id
        self.num_envs = int(n_envs)
        self._h = lib().orc_env_create(self.kind, self.num_envs, env_id_offset, seed, int(max_episode_steps or 0))
        if not self._h:
            raise ValueError("orc_env_create failed")
        for spec in env_wrappers:
            k, p = wrapper_params(spec)
            if lib().orc_env_attach_wrapper(self._h, k, p.ctypes.data, len(p)) != 0:
                raise ValueError(f"cannot attach {spec['id']} to {env_id}")
        D = OBS_DIM[self.kind]
        self.single_observation_space = SimpleNamespace(shape=(D,), dtype=np.float32)
        self.single_action_space = SimpleNamespace(n=N_ACTIONS[self.kind], shape=())
        self.render_mode = None
        self._threads = max(1, int(threads))
        self._pool = ThreadPoolExecutor(self._threads) if self._threads > 1 else None

    def close(self):
        if self._h:
            lib().orc_env_destroy(self._h)
            self._h = None

    __del__ = close

    def set_state(self, state: np.ndarray, elapsed: np.ndarray | None = None):
        s = np.ascontiguousarray(state, dtype=np.float64)
        assert s.shape == (STATE_DIM[self.kind], self.num_envs)
        e = None if elapsed is None else np.ascontiguousarray(elapsed, dtype=np.int32)
        lib().orc_env_set_state(self._h, s.ctypes.data, None if e is None else e.ctypes.data)

    def get_state(self):
        s = np.zeros((STATE_DIM[self.kind], self.num_envs), dtype=np.float64)
        e = np.zeros(self.num_envs, dtype=np.int32)
        lib().orc_env_get_state(self._h, s.ctypes.data, e.ctypes.data)
        return s, e

    def reset(self, **_):
        obs = np.zeros((self.num_envs, OBS_DIM[self.kind]), dtype=np.float32)
        lib().orc_env_reset(self._h, obs.ctypes.data)
        return obs, {}

    def step(self, actions):
        n = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((n, OBS_DIM[self.kind]), dtype=np.float32)
        rew = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        ep_r = np.zeros(n, dtype=np.float64)
        ep_l = np.zeros(n, dtype=np.int32)
        ptrs = [x.ctypes.data for x in (a, obs, rew, term, trunc, ep_r, ep_l)]
        if self._pool is None:
            lib().orc_env_step(self._h, *ptrs)
        else:
            cuts = np.linspace(0, n, self._threads + 1).astype(np.int64)
            futs = [self._pool.submit(lib().orc_env_step_range, self._h, int(lo), int(hi), *ptrs)
                    for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]
            for f in futs:
                f.result()
        term_b, trunc_b = term.astype(bool), trunc.astype(bool)
        infos = {}
        done = term_b | trunc_b
        if done.any():
            infos["episode"] = {"r": ep_r, "l": ep_l}
            infos["_episode"] = done
        return obs, rew, term_b, trunc_b, infos
