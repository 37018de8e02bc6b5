"""Hyperparameter schedules (reference: trainer_callbacks/hyperparameter_scheduler.py:8-37, 76-103 and
utils/schedule_resolver.py:8-52): interpolation kinds, optional linear warm-up from end_value to start_value, and the
position -> progress-fraction rule (positions <= 1 are fractions of max_env_steps, larger ones absolute env steps)."""
from __future__ import annotations

import math
from typing import Optional


def linear(start_value: float, end_value: float, fraction: float) -> float:
    f = max(0.0, min(fraction, 1.0))
    return start_value + (end_value - start_value) * f


def cosine(start_value: float, end_value: float, fraction: float) -> float:
    f = max(0.0, min(fraction, 1.0))
    return end_value + (start_value - end_value) * (0.5 * (1 + math.cos(math.pi * f)))


def exponential(start_value: float, end_value: float, fraction: float) -> float:
    """Decay with rate 2, normalised so that fraction 0 gives start_value and fraction 1 gives end_value."""
    f = max(0.0, min(fraction, 1.0))
    rate = 2.0
    span = 1.0 - math.exp(-rate)
    return end_value + (start_value - end_value) * ((math.exp(-rate * f) - math.exp(-rate)) / span)


SCHEDULERS_MAP = {"linear": linear, "cosine": cosine, "exponential": exponential}


def position_to_env_steps(raw: Optional[float], *, param: str, default_to_max: bool, max_env_steps: Optional[float]) -> float:
    """utils/schedule_resolver.py:8-52 in env steps (the reference divides both ends and the counter by n_envs: same fraction)."""
    if raw is None:
        if default_to_max:
            if max_env_steps is None:
                raise ValueError(f"{param}_schedule requires config.max_env_steps or an explicit {param}_schedule_end.")
            return float(max_env_steps)
        return 0.0
    value = float(raw)
    if value < 0.0:
        raise ValueError(f"{param}_schedule start/end must be non-negative.")
    if value <= 1.0:
        if max_env_steps is None:
            raise ValueError(f"{param}_schedule uses fractional start/end but config.max_env_steps is not set.")
        return value * float(max_env_steps)
    return value


def progress_fraction(total_steps: float, start_step: float, end_step: float) -> float:
    if total_steps <= start_step:
        return 0.0
    if total_steps >= end_step or end_step == start_step:
        return 1.0
    return (total_steps - start_step) / (end_step - start_step)


def scheduled_value(schedule: str, start_value: float, end_value: float, fraction: float, warmup_fraction: float = 0.0) -> float:
    """HyperparameterSchedulerCallback.on_train_epoch_end (:76-96)."""
    if schedule not in SCHEDULERS_MAP:
        raise ValueError(f"invalid schedule: {schedule}")
    if not (0.0 <= warmup_fraction < 1.0):
        raise ValueError(f"warmup_fraction must be in [0, 1), got {warmup_fraction}")
    if warmup_fraction > 0.0 and fraction < warmup_fraction:
        return end_value + (start_value - end_value) * (fraction / warmup_fraction)
    if warmup_fraction > 0.0:
        fraction = (fraction - warmup_fraction) / (1.0 - warmup_fraction)
    return SCHEDULERS_MAP[schedule](start_value, end_value, fraction)
