"""Trainer shell: the hook protocol the reference drives its agents with (PyTorch-Lightning ``Trainer`` + ``Callback`` as used by
utils/trainer_factory.py:9-46, utils/callback_builder.py:24-175 and trainer_callbacks/*), without Lightning.

One "epoch" = one rollout + all its passes.  Hook order per epoch, as the reference's fit loop produces it:

    callbacks.on_train_epoch_start -> agent.on_train_epoch_start (budget check, collect) -> agent.train_on_rollout
    -> callbacks.on_train_epoch_end    (DispatchMetrics -> schedulers -> EarlyStopping, in callback_builder.py's order)
    -> every eval_freq_epochs after the warm-up: callbacks.on_validation_epoch_start -> agent.validation_epoch
       -> callbacks.on_validation_epoch_end (DispatchMetrics -> EarlyStopping -> ModelCheckpoint)

``BaseAgent.learn()`` is the same loop with the callbacks' arithmetic inlined; ``Trainer.fit`` exists so that callback classes written
against the reference's protocol (``on_*_epoch_*(trainer, pl_module)``, ``trainer.should_stop``, ``trainer.logged_metrics``,
``pl_module.log_dict / get_rollout_collector / set_hyperparameter / metrics_recorder``) run on the engine unchanged.
"""
from __future__ import annotations

import csv
import json
import math
import shutil
import time
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

from .utils.schedules import SCHEDULERS_MAP, position_to_env_steps, progress_fraction, scheduled_value


class Callback:
    """pytorch_lightning.Callback, as far as the reference's trainer_callbacks use it."""

    def on_fit_start(self, trainer, pl_module) -> None: ...
    def on_train_epoch_start(self, trainer, pl_module) -> None: ...
    def on_train_epoch_end(self, trainer, pl_module) -> None: ...
    def on_validation_epoch_start(self, trainer, pl_module) -> None: ...
    def on_validation_epoch_end(self, trainer, pl_module) -> None: ...
    def on_fit_end(self, trainer, pl_module) -> None: ...
    def on_exception(self, trainer, pl_module, exception) -> None: ...


class Trainer:
    def __init__(self, *, callbacks: Iterable[Callback] = (), max_epochs: Optional[int] = None, loggers: Iterable[Any] = ()):
        self.callbacks: List[Callback] = list(callbacks)
        self.max_epochs = max_epochs
        self.loggers = list(loggers)
        self.should_stop = False
        self.current_epoch = 0
        self.logged_metrics: Dict[str, Any] = {}

    def _call(self, hook: str, agent, *args) -> None:
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)          # callbacks written against Lightning need not define every hook
            if fn is not None:
                fn(self, agent, *args)

    def log_dict(self, metrics: Dict[str, Any]) -> None:
        self.logged_metrics.update(metrics)
        for lg in self.loggers:
            lg.log_metrics(metrics, step=self.current_epoch)

    def fit(self, agent) -> Dict[str, Any]:
        cfg = agent.config
        agent.trainer = self
        agent.on_fit_start()
        self._call("on_fit_start", agent)
        max_epochs = self.max_epochs if self.max_epochs is not None else cfg.max_epochs
        try:
            while not self.should_stop:
                if max_epochs is not None and agent.current_epoch >= max_epochs:
                    agent.set_early_stop_reason(f"max_epochs={max_epochs} reached.")
                    break
                self.current_epoch = agent.current_epoch
                self._call("on_train_epoch_start", agent)
                if not agent.on_train_epoch_start():            # env-step budget exhausted: no rollout is collected (base_agent.py:306-320)
                    self.should_stop = True
                    break
                agent.train_on_rollout(agent._trajectories)
                agent.get_rollout_collector("train").resolve_episodes_async()
                self._call("on_train_epoch_end", agent)
                if cfg.eval_freq_epochs and (agent.current_epoch + 1) % int(cfg.eval_freq_epochs) == 0 \
                        and agent.current_epoch + 1 >= int(cfg.eval_warmup_epochs):
                    self._call("on_validation_epoch_start", agent)
                    agent.validation_epoch()
                    self._call("on_validation_epoch_end", agent)
                agent.current_epoch += 1
                if getattr(agent, "world_size", 1) > 1:         # callbacks judged rank-local metrics: leave the lock-step loop together
                    from .utils import distributed as D
                    self.should_stop = D.agree_any(self.should_stop, agent.device, agent.world_size)
        except BaseException as exc:                        # Lightning's on_exception: callbacks, then the module (joins a background evaluation)
            self._call("on_exception", agent, exc)
            if hasattr(agent, "on_exception"):
                agent.on_exception(self, agent, exc)
            raise
        if hasattr(agent, "on_fit_end"):                     # module hook (Lightning calls it when defined): joins a background evaluation
            agent.on_fit_end()
        self._call("on_fit_end", agent)
        for lg in self.loggers:
            lg.close()
        col = agent.get_rollout_collector("train")
        return {"stop_reason": agent._early_stop_reason, "epochs": agent.current_epoch, "total_env_steps": col.total_steps * agent.world_size,
                "best_eval_reward": agent.best_eval_reward, "elapsed_s": time.time() - agent._fit_t0}


# ---------------------------------------------------------------------------------------------------------------- callbacks
class DispatchMetricsCallback(Callback):
    """trainer_callbacks/dispatch_metrics.py:8-101: collector metrics + epoch means of the per-minibatch metrics, prefixed with the
    stage, through ``log_dict`` (one D2H copy of the device metric sums per epoch)."""

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        pl_module.metrics_recorder.reset_epoch("train")
        self._t_epoch, self._steps_epoch = time.time(), pl_module.get_rollout_collector("train").total_steps

    def on_train_epoch_end(self, trainer, pl_module) -> None:
        self._dispatch(trainer, pl_module, "train")

    def on_validation_epoch_start(self, trainer, pl_module) -> None:
        pl_module.metrics_recorder.reset_epoch("val")

    def on_validation_epoch_end(self, trainer, pl_module) -> None:
        self._dispatch(trainer, pl_module, "val")

    def _dispatch(self, trainer, pl_module, stage: str) -> None:
        col = pl_module.get_rollout_collector(stage)
        rollout = {k: v for k, v in col.get_metrics().items() if not k.endswith("_dist")}
        epoch_means = pl_module.pop_epoch_metrics() if stage == "train" else pl_module.metrics_recorder.epoch_means(stage)
        elapsed = time.time() - pl_module._fit_t0
        out = {**rollout, **epoch_means, "sys/timing/time_elapsed": elapsed, "cnt/epoch": pl_module.current_epoch}
        if stage == "train":
            steps = col.total_steps * pl_module.world_size
            out["cnt/total_env_steps"] = steps
            out["sys/timing/fps"] = steps / elapsed if elapsed > 0 else 0.0
            dt = time.time() - getattr(self, "_t_epoch", pl_module._fit_t0)
            out["sys/timing/fps_instant"] = (col.total_steps - getattr(self, "_steps_epoch", 0)) * pl_module.world_size / dt if dt > 0 else 0.0
            out["sys/timing/eps"] = (pl_module.current_epoch + 1) / elapsed if elapsed > 0 else 0.0
            if pl_module.config.max_env_steps is not None:
                out["progress"] = float(pl_module.calc_training_progress())
                if out["sys/timing/fps"] > 0:
                    out["sys/timing/eta_s"] = float(pl_module.config.max_env_steps) / out["sys/timing/fps"]
        pl_module.log_dict({f"{stage}/{k}": v for k, v in out.items()})


class HyperparameterSchedulerCallback(Callback):
    """trainer_callbacks/hyperparameter_scheduler.py:39-116 (positions in env steps: same fraction as the reference's vec steps)."""

    def __init__(self, *, schedule: str, parameter: str, start_value: float, end_value: float, start_step: float, end_step: float,
                 warmup_fraction: float = 0.0, set_value_fn=None):
        if schedule not in SCHEDULERS_MAP:
            raise ValueError(f"invalid schedule: {schedule}")
        if end_step < start_step:
            raise ValueError("schedule end_step must be >= start_step")
        if not (0.0 <= warmup_fraction < 1.0):
            raise ValueError(f"warmup_fraction must be in [0, 1), got {warmup_fraction}")
        self.schedule, self.parameter, self.start_value, self.end_value = schedule, parameter, start_value, end_value
        self.start_step, self.end_step, self.warmup_fraction = start_step, end_step, warmup_fraction
        self.set_value_fn = set_value_fn or (lambda module, value: module.set_hyperparameter(self.parameter, value))

    def on_train_epoch_end(self, trainer, pl_module) -> None:
        col = pl_module.get_rollout_collector("train")
        steps = float(col.total_steps * getattr(pl_module, "world_size", 1))
        frac = progress_fraction(steps, self.start_step, self.end_step)
        self.set_value_fn(pl_module, scheduled_value(self.schedule, self.start_value, self.end_value, frac, self.warmup_fraction))


class EarlyStoppingCallback(Callback):
    """trainer_callbacks/early_stopping.py:15-77: stop when a logged metric crosses a threshold; a None threshold disables it."""

    def __init__(self, metric_key: str, threshold: Optional[float], mode: str = "max"):
        assert mode in {"max", "min"}, "mode must be 'max' or 'min'"
        self.metric_key, self.threshold, self.mode = metric_key, threshold, mode

    def on_train_epoch_end(self, trainer, pl_module) -> None:
        self._maybe_stop(trainer, pl_module)

    def on_validation_epoch_end(self, trainer, pl_module) -> None:
        self._maybe_stop(trainer, pl_module)

    def _maybe_stop(self, trainer, pl_module) -> None:
        if trainer.should_stop or self.threshold is None:
            return
        value = trainer.logged_metrics.get(self.metric_key)
        if value is None:
            return
        hit = float(value) >= self.threshold if self.mode == "max" else float(value) <= self.threshold
        if not hit:
            return
        trainer.should_stop = True
        stage = self.metric_key.split("/")[0] if "/" in self.metric_key else "train"
        pl_module.metrics_recorder.record(stage, {"solved": 1})
        op = ">=" if self.mode == "max" else "<="
        pl_module.set_early_stop_reason(f"'{self.metric_key}': {float(value):.2f} {op} {float(self.threshold):.2f}.")


class ModelCheckpointCallback(Callback):
    """trainer_callbacks/model_checkpoint.py:15-106: after a validation epoch save when it is the first evaluation, the best so far,
    or training is stopping; ``<dir>/epoch=NN`` directories with ``last`` / ``best`` links (utils/run.py:203-212)."""

    def __init__(self, checkpoint_dir, metric: str = "val/roll/ep_rew/mean", mode: str = "max"):
        self.dir, self.metric, self.mode = Path(checkpoint_dir), metric, mode
        self.best_value = float("-inf") if mode == "max" else float("inf")
        self.first_eval = True

    def on_validation_epoch_end(self, trainer, pl_module) -> None:
        if self.metric not in trainer.logged_metrics:
            return
        value = float(trainer.logged_metrics[self.metric])
        world = getattr(pl_module, "world_size", 1)
        stopping = trainer.should_stop
        if world > 1:                                           # save_checkpoint is collective (every rank writes its env shard): one decision
            from .utils import distributed as D
            value = D.broadcast_value(value, pl_module.device, world)
            stopping = D.agree_any(stopping, pl_module.device, world)
        is_best = value > self.best_value if self.mode == "max" else value < self.best_value
        if not (self.first_eval or is_best or stopping):
            return
        mark_best = is_best or self.first_eval
        if mark_best:
            self.best_value = value
        self.first_eval = False
        target = self.dir / f"epoch={pl_module.current_epoch:02d}"
        pl_module.save_checkpoint(target)
        if getattr(pl_module, "rank", 0) == 0:
            scalars = {k: float(v) for k, v in trainer.logged_metrics.items() if isinstance(v, (int, float)) and math.isfinite(float(v))}
            (target / "metrics.json").write_text(json.dumps(scalars, indent=2, sort_keys=True))
            for name in (["last", "best"] if mark_best else ["last"]):
                link = self.dir / name
                if link.is_symlink() or link.exists():
                    link.unlink() if link.is_symlink() or link.is_file() else shutil.rmtree(link)
                link.symlink_to(target.name, target_is_directory=True)


class CsvMetricsLogger:
    """loggers/metrics_csv_logger.py in spirit: one row per log_dict call, columns grow as keys appear (rank 0 only)."""

    def __init__(self, path, enabled: bool = True):
        self.path, self.enabled, self.rows = Path(path), enabled, []

    def log_metrics(self, metrics: Dict[str, Any], step: int) -> None:
        if self.enabled:
            self.rows.append({"epoch": step, **{k: v for k, v in metrics.items() if isinstance(v, (int, float, str))}})

    def close(self) -> None:
        if not self.enabled or not self.rows:
            return
        keys = sorted({k for r in self.rows for k in r})
        self.path.parent.mkdir(parents=True, exist_ok=True)
        with open(self.path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            w.writerows(self.rows)


def build_callbacks(agent, *, checkpoint_dir=None) -> List[Callback]:
    """utils/callback_builder.py:32-103 for the callbacks that exist here, in the reference's order: metrics dispatch, schedulers,
    early stopping (train threshold, then eval threshold), checkpointing (after early stopping so it sees trainer.should_stop)."""
    cfg = agent.config
    cbs: List[Callback] = [DispatchMetricsCallback()]
    for param in ("policy_lr", "ent_coef", "vf_coef", "clip_range", "clip_range_vf"):
        kind = getattr(cfg, f"{param}_schedule", None)
        if not kind:
            continue
        s0 = position_to_env_steps(getattr(cfg, f"{param}_schedule_start", None), param=param, default_to_max=False, max_env_steps=cfg.max_env_steps)
        s1 = position_to_env_steps(getattr(cfg, f"{param}_schedule_end", None), param=param, default_to_max=True, max_env_steps=cfg.max_env_steps)
        cbs.append(HyperparameterSchedulerCallback(schedule=kind, parameter=param, start_value=float(getattr(cfg, f"{param}_schedule_start_value")),
                                                   end_value=float(getattr(cfg, f"{param}_schedule_end_value")), start_step=s0, end_step=s1,
                                                   warmup_fraction=float(getattr(cfg, f"{param}_schedule_warmup", 0.0) or 0.0)))
    thr = cfg.early_stop_on_train_threshold
    if thr:
        cbs.append(EarlyStoppingCallback("train/roll/ep_rew/mean", agent.get_env("train").get_return_threshold() if thr is True else float(thr)))
    thr = cfg.early_stop_on_eval_threshold
    if thr:
        limit = cfg.reward_threshold if cfg.reward_threshold is not None else agent.get_env("val").get_return_threshold()
        cbs.append(EarlyStoppingCallback("val/roll/ep_rew/mean", limit if thr is True else float(thr)))
    if checkpoint_dir is not None:
        cbs.append(ModelCheckpointCallback(checkpoint_dir))
    return cbs
