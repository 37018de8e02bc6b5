"""``<env>:<variant>`` YAML configuration (reference: utils/config.py:17-889).

Same file format (YAML anchors / ``<<:`` merges; every top-level mapping that is not a Config field and does not start
with ``_`` is a variant), same field names, defaults, fractional ``batch_size``, numeric-string coercion, ``model_id``
resolution and schedule-dict expansion as the reference, restricted to what the rollout-and-update path consumes.
Engine additions are declared fields (unknown keys are dropped, like the reference): ``engine``, ``devices``,
``store_next_obs``, ``minibatch_shuffle``.
"""
from __future__ import annotations

import json
import os
from dataclasses import MISSING, asdict, dataclass, field, fields
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import yaml

from .model_registry import resolve_model_spec

_SCHEDULABLE_BASE = ("policy_lr", "ent_coef")
_SCHEDULABLE_PPO = ("vf_coef", "clip_range", "clip_range_vf")


def _sanitize(name: str) -> str:
    return name.replace("/", "-").replace("\\", "-")


@dataclass
class Config:
    project_id: str = ""
    env_id: str = ""
    description: str = ""
    spec: Dict[str, Any] = field(default_factory=dict)
    n_steps: Optional[int] = None
    batch_size: Optional[Union[int, float]] = None
    n_epochs: Optional[int] = None
    max_epochs: Optional[int] = None
    max_env_steps: Optional[int] = None
    max_episode_steps: Optional[int] = None
    seed: int = 42
    seed_train: int = 42
    seed_val: int = 1042
    seed_test: int = 2042
    n_envs: Union[int, str] = "auto"
    reward_threshold: Optional[float] = None
    env_wrappers: list = field(default_factory=list)
    env_kwargs: dict = field(default_factory=dict)
    vectorization_mode: Optional[str] = "auto"
    normalize_obs: bool = False
    obs_type: str = "vector"
    policy: str = "mlp"
    model_id: Optional[str] = None
    policy_lr: Optional[Union[float, Dict[str, Any]]] = None
    optimizer: str = "adam"
    max_grad_norm: Optional[float] = None
    gamma: Optional[float] = None
    ent_coef: Optional[Union[float, Dict[str, Any]]] = None
    returns_type: Optional[str] = None
    normalize_returns: Optional[str] = None
    policy_targets: Optional[str] = None
    eval_warmup_epochs: Union[int, float] = 0
    eval_episodes: int = 100
    eval_freq_epochs: Optional[int] = None
    eval_deterministic: bool = False
    eval_async: bool = False
    early_stop_on_train_threshold: Union[bool, float] = False
    early_stop_on_eval_threshold: Union[bool, float] = True
    accelerator: str = "auto"
    devices: Optional[Union[int, str]] = None
    quiet: bool = False
    enable_wandb: bool = True
    init_from_run: Optional[str] = None
    # ---- engine fields (new) ----
    engine: str = "b200"
    store_next_obs: bool = False          # the update path never reads next_observations; keep them only on request
    minibatch_shuffle: str = "device"     # "device": keyed bijection inside the kernel; "torch": argsort(rand) index tensors
    track_activations: bool = True
    fused_update: bool = True             # step tail (reduction, gradient exchange, clip, Adam) as ONE launch (gs_update_finish)
    grad_allreduce: str = "auto"          # several ranks: "peer" = NVLink P2P exchange inside gs_update_finish; "nccl" = torch.distributed;
                                          # "auto" = peer when every rank can map every other rank's buffer, else nccl
    _hidden_dims: Optional[Tuple[int, ...]] = field(default=None, init=False, repr=False)
    _activation: Optional[str] = field(default=None, init=False, repr=False)
    _policy_kwargs: Optional[Dict[str, Any]] = field(default=None, init=False, repr=False)

    # -------------------------------------------------------------------------------------------- construction
    @classmethod
    def build_from_dict(cls, config_dict: Dict[str, Any]) -> "Config":
        d = dict(config_dict)
        algo_id = d.pop("algo_id")
        config_cls = {"reinforce": REINFORCEConfig, "ppo": PPOConfig}[algo_id]
        valid = {f.name for f in fields(config_cls) if f.init}
        return config_cls(**{k: v for k, v in d.items() if k in valid})

    @classmethod
    def build_from_yaml(cls, config_id: str, variant_id: str = None, config_dir: str = "config/environments") -> "Config":
        root = Path(__file__).resolve().parent.parent.parent
        cfg_dir = Path(config_dir) if os.path.isabs(str(config_dir)) else root / config_dir
        known = {f.name for f in fields(cls)}
        table: Dict[str, Dict[str, Any]] = {}
        for path in sorted(cfg_dir.glob("*.yaml")):
            doc = yaml.safe_load(path.read_text()) or {}
            base = {k: v for k, v in doc.items() if k in known}
            for key, val in doc.items():
                if key in known or not isinstance(val, dict) or str(key).startswith("_"):
                    continue
                variant = dict(base)
                variant.update(val)
                if not variant.get("project_id"):
                    env_id = variant.get("env_id", "")
                    variant["project_id"] = f"{env_id}_{variant.get('obs_type', 'rgb')}" if env_id else path.stem
                names = {f"{variant['project_id']}_{key}", f"{_sanitize(variant['project_id'])}_{key}"}
                if variant.get("env_id"):
                    names |= {f"{variant['env_id']}_{key}", f"{_sanitize(variant['env_id'])}_{key}"}
                for nme in names:
                    table.setdefault(nme, variant)
        return cls.build_from_dict(table[f"{config_id}_{variant_id}"])

    def __post_init__(self):
        for f in fields(self):
            if getattr(self, f.name) is None:
                if f.default is not MISSING:
                    setattr(self, f.name, f.default)
                elif f.default_factory is not MISSING:
                    setattr(self, f.name, f.default_factory())
        if self.n_envs == "auto":
            self.n_envs = os.cpu_count() or 1
        for f in fields(self):                        # YAML reads 1e5 as a string
            v = getattr(self, f.name)
            if isinstance(v, str):
                try:
                    setattr(self, f.name, float(v))
                except ValueError:
                    pass
        self._resolve_batch_size()
        w = self.eval_warmup_epochs
        if 0 < w < 1:
            assert self.max_env_steps is not None, "Fractional eval_warmup_epochs requires max_env_steps to be set"
            self.eval_warmup_epochs = int(self.max_env_steps / (self.n_envs * self.n_steps) * w)
        self._resolve_schedules()
        assert self.model_id is not None, "model_id is required. Available models: mlp_tiny, mlp_64x64, mlp_small, mlp_medium, mlp_large"
        ms = resolve_model_spec(self.model_id)
        self.policy, self._hidden_dims, self._activation, self._policy_kwargs = ms.policy, ms.hidden_dims, ms.activation, dict(ms.policy_kwargs)
        self.validate()

    def _resolve_batch_size(self) -> None:
        if self.batch_size is None:
            self.batch_size = 64
        if self.batch_size > 1:
            self.batch_size = int(self.batch_size)
            return
        self.batch_size = max(1, int(int(self.n_envs) * int(self.n_steps) * self.batch_size))

    def _schedulable(self) -> Tuple[str, ...]:
        return _SCHEDULABLE_BASE

    def _resolve_schedules(self) -> None:
        for key in self._schedulable():
            v = getattr(self, key, None)
            if not isinstance(v, dict):
                continue
            assert v.get("start") is not None, f"{key} schedule dict must have 'start' key"
            start, end = float(v["start"]), float(v.get("end", 0.0))
            setattr(self, key, start)
            setattr(self, f"{key}_schedule", v.get("schedule", "linear"))
            setattr(self, f"{key}_schedule_start_value", start)
            setattr(self, f"{key}_schedule_end_value", end)
            setattr(self, f"{key}_schedule_start", float(v.get("from", 0.0)))
            setattr(self, f"{key}_schedule_end", float(v.get("to", 1.0)))
            if float(v.get("warmup", 0.0)) > 0.0:
                setattr(self, f"{key}_schedule_warmup", float(v["warmup"]))

    # -------------------------------------------------------------------------------------------- accessors
    @property
    def hidden_dims(self) -> Tuple[int, ...]:
        return self._hidden_dims

    @property
    def activation(self) -> str:
        return self._activation

    @property
    def policy_kwargs(self) -> Dict[str, Any]:
        return self._policy_kwargs

    @property
    def max_vec_steps(self) -> Optional[int]:
        return None if self.max_env_steps is None else int(self.max_env_steps) // int(self.n_envs)

    def get_env_args(self) -> Dict[str, Any]:
        return dict(env_id=self.env_id, project_id=self.project_id, env_spec=self.spec, n_envs=self.n_envs, seed=self.seed,
                    max_episode_steps=self.max_episode_steps, env_wrappers=self.env_wrappers, normalize_obs=self.normalize_obs,
                    obs_type=self.obs_type, render_mode=None, vectorization_mode=self.vectorization_mode, record_video=False,
                    record_video_kwargs={}, env_kwargs=self.env_kwargs)

    def rollout_collector_hyperparams(self) -> Dict[str, Any]:
        out = {"gamma": self.gamma, "normalize_returns": self.normalize_returns == "rollout", "returns_type": self.returns_type}
        if hasattr(self, "gae_lambda"):
            out["gae_lambda"] = self.gae_lambda
        if hasattr(self, "advantages_type"):
            out["advantages_type"] = self.advantages_type
        if hasattr(self, "normalize_advantages"):
            out["normalize_advantages"] = self.normalize_advantages == "rollout"
        return out

    def get_rollout_collector_kwargs(self) -> Dict[str, Any]:
        return {"n_steps": self.n_steps, **self.rollout_collector_hyperparams()}

    def save_to_json(self, path: str) -> None:
        data = {k: v for k, v in asdict(self).items() if not k.startswith("_")}
        data["algo_id"] = self.algo_id
        for k, v in vars(self).items():
            if "_schedule" in k and not k.startswith("_") and v is not None:
                data[k] = v
        Path(path).write_text(json.dumps(data, indent=2, default=str))

    # -------------------------------------------------------------------------------------------- validation
    def validate(self) -> None:
        def positive(name, allow_none=True):
            v = getattr(self, name, None)
            if v is None:
                if not allow_none:
                    raise ValueError(f"{name} must be set.")
                return
            if not v > 0:
                raise ValueError(f"{name} must be a positive number.")

        positive("seed", allow_none=False)
        positive("n_envs", allow_none=False)
        for name in ("policy_lr", "n_steps", "batch_size", "n_epochs", "max_env_steps", "max_epochs", "max_episode_steps", "max_grad_norm", "eval_episodes"):
            positive(name)
        if self.gamma is not None and not (0 < self.gamma <= 1):
            raise ValueError("gamma must be in (0, 1].")
        if self.ent_coef is not None and self.ent_coef < 0:
            raise ValueError("ent_coef must be a non-negative number.")
        rollout_size = int(self.n_envs) * int(self.n_steps)
        if self.batch_size > rollout_size:
            raise ValueError(f"batch_size ({self.batch_size}) must be <= n_envs*n_steps ({rollout_size}).")
        if rollout_size % int(self.batch_size) != 0:
            raise ValueError(f"batch_size must divide n_envs*n_steps exactly: rollout_size={rollout_size}, batch_size={self.batch_size}.")
        if self.policy_targets is not None and self.policy_targets not in ("returns", "advantages"):
            raise ValueError("policy_targets must be 'returns' or 'advantages'.")
        if self.engine != "b200":
            raise ValueError("engine must be 'b200' in this repository (the reference loop lives upstream).")


@dataclass
class REINFORCEConfig(Config):
    policy: str = "mlp"
    n_steps: int = 2048
    batch_size: Union[int, float] = 2048
    n_epochs: int = 1
    policy_lr: Union[float, Dict[str, Any]] = 1e-2
    gamma: float = 0.99
    ent_coef: Union[float, Dict[str, Any]] = 0.01
    max_grad_norm: float = 0.5
    returns_type: str = "mc:rtg"
    policy_targets: str = "returns"
    # the reference reads config.normalize_advantages in REINFORCEAgent but never declares it: declared here
    normalize_advantages: str = "off"
    advantages_type: str = "baseline"

    @property
    def algo_id(self) -> str:
        return "reinforce"


@dataclass
class PPOConfig(Config):
    policy: str = "mlp_actorcritic"
    n_steps: int = 2048
    batch_size: Union[int, float] = 64
    n_epochs: int = 10
    policy_lr: Union[float, Dict[str, Any]] = 3e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: Union[float, Dict[str, Any]] = 0.2
    clip_range_vf: Union[float, Dict[str, Any]] = 0.2
    target_kl: Optional[float] = None
    ent_coef: Union[float, Dict[str, Any]] = 0.0
    vf_coef: Union[float, Dict[str, Any]] = 0.5
    max_grad_norm: float = 0.5
    returns_type: str = "gae:rtg"
    advantages_type: str = "gae"
    policy_targets: str = "advantages"
    normalize_advantages: str = "batch"

    @property
    def algo_id(self) -> str:
        return "ppo"

    def _schedulable(self) -> Tuple[str, ...]:
        return _SCHEDULABLE_BASE + _SCHEDULABLE_PPO

    def validate(self) -> None:
        super().validate()
        if self.target_kl is not None and not self.target_kl > 0:
            raise ValueError("target_kl must be a positive number.")
        if not (0 <= self.gae_lambda <= 1):
            raise ValueError("gae_lambda must be in [0, 1].")
        for name in ("clip_range", "clip_range_vf"):
            v = getattr(self, name)
            if v is not None and not (0 < v < 1):
                raise ValueError(f"{name} must be in (0, 1).")
        if self.vf_coef is not None and self.vf_coef < 0:
            raise ValueError("vf_coef must be a non-negative number.")
        if self.normalize_advantages not in ("rollout", "batch", "off"):
            raise ValueError("normalize_advantages must be 'rollout', 'batch', or 'off'.")


def load_config(config_id: str, variant_id: str = None, config_dir: str = "config/environments") -> Config:
    return Config.build_from_yaml(config_id, variant_id, config_dir)
