"""Host-side seeding (reference: utils/random.py:12-41).

Only the HOST generators live here: the python / numpy / torch streams that model initialisation and the ``torch`` minibatch shuffle
draw from.  Env resets and action sampling on the device use counter-based Philox streams keyed by (seed, global env id, step), owned by
the env handles and the collector (csrc/common.cuh), so they need no global state and no seeding call.
"""
from __future__ import annotations

import random as _stdlib_random
from typing import Dict, Optional

import numpy as np
import torch

_shared: Dict[str, torch.Generator] = {}      # the one process-wide shuffle generator, created on first use


def get_global_torch_generator(seed: Optional[int] = None) -> torch.Generator:
    """The process-wide ``torch.Generator`` of the minibatch shuffle.  The FIRST call may seed it; later calls return the same object
    and ignore ``seed`` (the reference's contract: one shared stream, seeded once)."""
    gen = _shared.get("shuffle")
    if gen is None:
        gen = _shared["shuffle"] = torch.Generator()
        if seed is not None:
            gen.manual_seed(int(seed))
    return gen


def set_random_seed(seed: int) -> None:
    """python ``random``, numpy's legacy global state and torch (CPU + every CUDA device)."""
    for seed_fn in (_stdlib_random.seed, np.random.seed, torch.manual_seed):
        seed_fn(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
