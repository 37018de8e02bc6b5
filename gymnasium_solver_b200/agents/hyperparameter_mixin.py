"""Live hyper-parameter management of an agent (reference agents/hyperparameter_mixin.py:10-112).

What the reference does at the start of every training epoch (agents/base_agent.py:297-302): re-read the run's ``config.json`` -- a user
may have edited it while training runs -- apply every changed SCALAR that has no active schedule, then log the tunable values under
``hp/*``.  Same names, same rules here; the run object only has to offer ``load_config()`` (``RunConfigFile`` below reads a run
directory's ``config.json``; the reference's ``utils.run.Run`` satisfies the same protocol).
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict, is_dataclass
from typing import Any, Dict

_SKIPPED_TYPES = (list, tuple, dict)


def _as_mapping(config) -> Dict[str, Any]:
    if is_dataclass(config):
        return asdict(config)
    if isinstance(config, dict):
        return dict(config)
    return {k: v for k, v in vars(config).items() if not k.startswith("_")}


class RunConfigFile:
    """The part of the reference's ``Run`` (utils/run.py:119, 185) the agent reads while training: ``<run_dir>/config.json``."""

    def __init__(self, run_dir: str, filename: str = "config.json"):
        self.run_dir, self.path = run_dir, os.path.join(run_dir, filename)

    def save_config(self, config) -> None:
        os.makedirs(self.run_dir, exist_ok=True)
        with open(self.path, "w") as f:
            json.dump({k: v for k, v in _as_mapping(config).items() if isinstance(v, (int, float, str, bool, list, dict)) or v is None}, f, indent=1)

    def load_config(self) -> Dict[str, Any]:
        with open(self.path) as f:
            return json.load(f)


class HyperparameterMixin:
    run = None          # set to an object with load_config() to enable the live re-read (None: nothing to read)

    def _change_optimizers_lr(self, lr: float) -> None:
        self.policy_lr = lr                                  # kept in sync for logging / inspection
        optimizers = self.optimizers()
        for opt in optimizers if isinstance(optimizers, (list, tuple)) else [optimizers]:
            for group in opt.param_groups:
                group["lr"] = lr

    def _change_n_epochs(self, n_epochs: int) -> None:
        self.n_epochs = n_epochs
        loader = getattr(self, "_train_dataloader", None)    # loader-driven training: the sampler's pass count follows
        if loader is not None and hasattr(getattr(loader, "sampler", None), "num_passes"):
            loader.sampler.num_passes = n_epochs

    def _load_run_config(self):
        """The run's current configuration.  Several ranks: rank 0 reads the file and every rank takes ITS reading, so a file edited while
        the ranks pass this point cannot be applied in different epochs on different ranks (the ranks' weights must stay bit-identical)."""
        world = int(getattr(self, "world_size", 1) or 1)
        if world > 1:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                box = [self.run.load_config() if int(getattr(self, "rank", 0)) == 0 else None]
                dist.broadcast_object_list(box, src=0)
                return box[0]
        return self.run.load_config()

    def _read_hyperparameters_from_run(self) -> None:
        if getattr(self, "run", None) is None:
            return
        loaded, current = _as_mapping(self._load_run_config()), _as_mapping(self.config)
        scheduled = {k[: -len("_schedule")] for k, v in vars(self.config).items() if k.endswith("_schedule") and v}
        changes = {}
        for key, value in loaded.items():
            if isinstance(value, _SKIPPED_TYPES) or key in scheduled:     # structured fields and scheduled parameters are not reloaded
                continue
            if value != current.get(key):
                changes[key] = value
        if changes:
            self.on_hyperparams_change(changes)

    def on_hyperparams_change(self, changes_map: Dict[str, Any]) -> None:
        for key, value in changes_map.items():
            if not hasattr(self.config, key):
                continue
            setattr(self.config, key, value)
            if key == "policy_lr":
                self._change_optimizers_lr(value)
            elif key in ("clip_range", "clip_range_vf", "vf_coef", "ent_coef"):
                setattr(self, key, value)
            elif key == "n_epochs":
                self._change_n_epochs(value)

    def _log_hyperparameters(self) -> None:
        names = ["n_epochs", "ent_coef", "vf_coef", "clip_range", "policy_lr"] + (["clip_range_vf"] if hasattr(self, "clip_range_vf") else [])
        self.metrics_recorder.record("train", {f"hp/{k}": getattr(self, k) for k in names if hasattr(self, k)})

    def set_hyperparameter(self, param: str, value: float) -> None:
        """Called by the hyper-parameter scheduler (trainer_callbacks/hyperparameter_scheduler.py)."""
        setattr(self, param, value)
        if hasattr(self.config, param):
            setattr(self.config, param, value)
        if param == "policy_lr" and getattr(self, "_optimizer", None) is not None:
            self._change_optimizers_lr(value)
