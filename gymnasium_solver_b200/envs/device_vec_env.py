"""GPU-resident vector environments behind the Gymnasium VectorEnv protocol.

Replaces what the reference builds at utils/environment.py:305-418 for the three classic-control ids
(``gym.make`` + per-env wrappers + ``TimeLimit`` + ``SyncVectorEnv`` + ``RecordEpisodeStatistics``): state lives in HBM
inside an opaque ``gs_env_t`` handle, every call is one kernel launch, observations/rewards/flags come back as CUDA
tensors.  ``RolloutCollector`` drives the handle directly through the fused collect kernel; ``reset``/``step`` exist for
protocol compatibility and for the parity tests.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from .. import _native as N

RETURN_THRESHOLDS = {"CartPole-v1": 475.0, "Acrobot-v1": -100.0, "MountainCar-v0": -110.0}
DEFAULT_MAX_EPISODE_STEPS = {"CartPole-v1": 500, "Acrobot-v1": 500, "MountainCar-v0": 200}


class _Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.int64

    def sample(self):
        return int(np.random.randint(self.n))


# observation_space.high of the upstream envs (gymnasium 1.1.1 classic_control; low = -high except MountainCar's position).  Stated from
# the upstream sources like the physics (un-vendored, so unpinned); the reference's YAML `spec:` blocks agree where finite.
_OBS_HIGH = {"CartPole-v1": [4.8, np.inf, 12 * 2 * np.pi / 360 * 2, np.inf],
             "Acrobot-v1": [1.0, 1.0, 1.0, 1.0, 4 * np.pi, 9 * np.pi],
             "MountainCar-v0": [0.6, 0.07]}
_OBS_LOW = {"MountainCar-v0": [-1.2, -0.07]}


class _Box:
    def __init__(self, shape, low=None, high=None):
        self.shape, self.dtype = tuple(shape), np.float32
        self.high = np.broadcast_to(np.asarray(np.inf if high is None else high, dtype=np.float32), self.shape).copy()
        self.low = -self.high if low is None else np.broadcast_to(np.asarray(low, dtype=np.float32), self.shape).copy()


class DeviceVecEnv:
    """``num_envs`` classic-control environments stepped by CUDA kernels (one thread per env)."""

    def __init__(self, env_id: str, n_envs: int, seed: int = 0, *, max_episode_steps: int | None = None,
                 device: int | torch.device | str | None = None, env_id_offset: int = 0, spec: dict | None = None):
        if env_id not in N.ENV_KINDS:
            raise KeyError(f"{env_id!r} is not a device environment; supported: {sorted(N.ENV_KINDS)}")
        if not torch.cuda.is_available():
            raise N.EngineError("DeviceVecEnv needs a CUDA device (the engine has no CPU path)")
        self.env_id = env_id
        self.kind = N.ENV_KINDS[env_id]
        self.num_envs = int(n_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed = int(seed)
        self.env_id_offset = int(env_id_offset)
        self.max_episode_steps = int(max_episode_steps) if max_episode_steps else DEFAULT_MAX_EPISODE_STEPS[env_id]
        self.obs_dim = N.lib().gs_env_obs_dim(self.kind)
        self.state_dim = N.lib().gs_env_state_dim(self.kind)
        self.n_actions = N.lib().gs_env_n_actions(self.kind)
        self.single_observation_space = _Box((self.obs_dim,), _OBS_LOW.get(env_id), _OBS_HIGH[env_id])
        self.single_action_space = _Discrete(self.n_actions)
        self.observation_space = _Box((self.num_envs, self.obs_dim), _OBS_LOW.get(env_id), _OBS_HIGH[env_id])
        self.action_space = SimpleNamespace(shape=(self.num_envs,), n=self.n_actions)
        self.render_mode = None
        self.spec_dict = dict(spec or {})
        self.wrappers: list[dict] = []
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_create(self.kind, self.num_envs, self.env_id_offset, self.seed & (2**64 - 1),
                                          int(max_episode_steps or 0), self.device.index, C.byref(h)))
        self._h = h

    # ---- handle management --------------------------------------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise N.EngineError("environment is closed")
        return self._h

    @property
    def unwrapped(self):
        return self

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib().gs_env_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def normalize_observations_static(self) -> None:
        """VecNormalizeStatic of the reference (gym_wrappers/vec_normalize_static.py:20-60; ``normalize_obs: static``), fused into the
        kernels that emit observations: bounded dims -> (x - low) / ((high - low) + 1e-8) in fp32, degenerate dims -> 0, unbounded
        dims pass through.  The observation spaces are rewritten the way the wrapper rewrites them."""
        sp = self.single_observation_space
        low, high = sp.low.astype(np.float32), sp.high.astype(np.float32)
        lo_c = (C.c_float * self.obs_dim)(*[float(x) for x in low])
        hi_c = (C.c_float * self.obs_dim)(*[float(x) for x in high])
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_set_obs_normalization(self.handle, lo_c, hi_c, self.obs_dim))
        finite = np.isfinite(low) & np.isfinite(high)
        pos, zero = finite & (high > low), finite & (high == low)
        low_n = np.where(pos | zero, 0.0, low).astype(np.float32)
        high_n = np.where(pos, 1.0, np.where(zero, 0.0, high)).astype(np.float32)
        self.single_observation_space = _Box((self.obs_dim,), low_n, high_n)
        self.observation_space = _Box((self.num_envs, self.obs_dim), low_n, high_n)
        self.normalize_obs = "static"

    def attach_wrapper(self, wrapper_id: str, params: list[float], spec: dict) -> None:
        arr = (C.c_double * len(params))(*[float(p) for p in params])
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_wrapper_attach(self.handle, N.WRAPPER_KINDS[wrapper_id], arr, len(params)))
        self.wrappers.append(dict(spec))

    # ---- VectorEnv protocol ---------------------------------------------------------------------------------------
    def reset(self, **_):
        obs = torch.empty(self.num_envs, self.obs_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_reset(self.handle, N.ptr(obs), N.stream()))
        return obs, {}

    def step(self, actions):
        """(obs, rewards, terminated, truncated, infos) as CUDA tensors; infos carries RecordEpisodeStatistics keys."""
        a = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
        if a.shape != (self.num_envs,):
            raise ValueError(f"expected actions of shape ({self.num_envs},), got {tuple(a.shape)}")
        n, dev = self.num_envs, self.device
        obs = torch.empty(n, self.obs_dim, dtype=torch.float32, device=dev)
        rew = torch.empty(n, dtype=torch.float32, device=dev)
        term = torch.empty(n, dtype=torch.uint8, device=dev)
        trunc = torch.empty(n, dtype=torch.uint8, device=dev)
        ep_r = torch.empty(n, dtype=torch.float64, device=dev)
        ep_l = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().gs_env_step(self.handle, N.ptr(a), N.ptr(obs), N.ptr(rew), N.ptr(term), N.ptr(trunc), N.ptr(ep_r),
                                        N.ptr(ep_l), N.stream()))
        term_b, trunc_b = term.bool(), trunc.bool()
        infos = {"episode": {"r": ep_r, "l": ep_l}, "_episode": term_b | trunc_b}
        return obs, rew, term_b, trunc_b, infos

    # ---- parity hooks ---------------------------------------------------------------------------------------------
    def set_state(self, state, elapsed=None) -> None:
        s = torch.as_tensor(state, dtype=torch.float64).to(self.device).contiguous()
        if s.shape != (self.state_dim, self.num_envs):
            raise ValueError(f"state must be ({self.state_dim}, {self.num_envs}) SoA, got {tuple(s.shape)}")
        e = None if elapsed is None else torch.as_tensor(elapsed, dtype=torch.int32).to(self.device).contiguous()
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_set_state(self.handle, N.ptr(s), N.ptr(e), N.stream()))
            torch.cuda.current_stream().synchronize()

    def get_state(self):
        s = torch.empty(self.state_dim, self.num_envs, dtype=torch.float64, device=self.device)
        e = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_get_state(self.handle, N.ptr(s), N.ptr(e), N.stream()))
        return s, e

    # ---- exact snapshot (checkpoint / resume) ------------------------------------------------------------------------
    def snapshot(self) -> torch.Tensor:
        """Everything the handle holds on the device as one uint8 tensor (gs_env_save)."""
        nbytes = N.lib().gs_env_snapshot_bytes(self.handle)
        if nbytes <= 0:
            raise N.EngineError(N.lib().gs_last_error().decode())
        blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_save(self.handle, N.ptr(blob), N.stream()))
        return blob

    def restore(self, blob: torch.Tensor) -> None:
        nbytes = N.lib().gs_env_snapshot_bytes(self.handle)
        if blob.dtype != torch.uint8 or blob.numel() != nbytes:
            raise ValueError(f"snapshot of {blob.numel()} bytes does not fit this env ({nbytes} bytes: kind, size or wrapper differ)")
        blob = blob.to(self.device).contiguous()
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_env_load(self.handle, N.ptr(blob), N.stream()))
            torch.cuda.current_stream().synchronize()

    # ---- EnvInfoWrapper surface used by callbacks (gym_wrappers/env_info.py) -----------------------------------------
    def get_return_threshold(self):
        try:
            return float(self.spec_dict["returns"]["threshold_solved"])
        except (KeyError, TypeError):
            return RETURN_THRESHOLDS[self.env_id]

    def get_spec(self):
        return self.spec_dict

    def get_max_episode_steps(self):
        return self.max_episode_steps

    @contextlib.contextmanager
    def recorder(self, video_path=None, record_video: bool = True, **_):
        """``with env.recorder(path, record_video=...)`` around an evaluation (reference gym_wrappers/env_video_recorder.py:180).  Device
        environments have no renderer (``render_mode`` is None), so like the reference's own non-rendering vector adapter
        (gym_wrappers/ale_vec_env_adapter.py:87-97) this is a no-op context: nothing is written to ``video_path``."""
        yield self
