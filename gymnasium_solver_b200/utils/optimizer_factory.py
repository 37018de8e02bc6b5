"""Optimizer construction (reference: utils/optimizer_factory.py:6-29): torch defaults, only lr is set."""
from __future__ import annotations

import torch


def build_optimizer(*, params, optimizer, lr: float, **extra) -> torch.optim.Optimizer:
    name = str(getattr(optimizer, "value", optimizer)).lower()
    cls = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW}[name]
    return cls(params, lr=lr, **extra)
