"""Seeding helpers (reference: utils/random.py:12-41)."""
from __future__ import annotations

import random as _py_random
from typing import Optional

import numpy as np
import torch

_global_torch_generator: Optional[torch.Generator] = None


def get_global_torch_generator(seed: Optional[int] = None) -> torch.Generator:
    global _global_torch_generator
    if _global_torch_generator is None:
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        _global_torch_generator = g
    return _global_torch_generator


def set_random_seed(seed: int) -> None:
    _py_random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
