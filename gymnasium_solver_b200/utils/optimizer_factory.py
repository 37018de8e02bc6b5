"""Optimizer construction (reference: utils/optimizer_factory.py:6-29): torch defaults, only lr is set."""
from __future__ import annotations

import torch


def build_optimizer(*, params, optimizer, lr: float, **extra) -> torch.optim.Optimizer:
    name = str(getattr(optimizer, "value", optimizer)).lower()
    cls = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW}[name]
    params = list(params)
    if name in ("adam", "adamw") and params and all(p.is_cuda for p in params):
        extra.setdefault("fused", True)      # same update rule as the default implementation, one kernel launch
    return cls(params, lr=lr, **extra)
