"""Minibatch loaders over a collected rollout (reference: utils/dataloaders.py:12-77, utils/datasets.py).

Compatibility surface for callers that drive ``training_step`` from a loader the way Lightning drove the reference
(``BaseAgent.train_dataloader``): ``num_passes`` independent permutations of the rollout from ``MultiPassRandomSampler``, cut into
consecutive minibatches, each gathered by ``collector.slice_trajectories``.  Same keyword surface, same errors, same index
batches from the same generator (pinned by tests/golden/dataloader.json, produced by running the reference's function).

Unlike the reference there is no ``torch.utils.data.DataLoader`` underneath: a pass is ONE int64 index tensor (no Python list of
``n_epochs * N * T`` ints, no worker processes to hand CUDA tensors to), sliced per minibatch.  The engine's own fit loop does
not use this module at all: its update kernels gather in place from the time-major buffer (BaseAgent.minibatches).
"""
from __future__ import annotations

from typing import Any, Callable, Iterator, Optional

import torch

from .samplers import MultiPassRandomSampler


class IndexDataset(torch.utils.data.Dataset):
    """Dataset of its own indices (reference utils/datasets.py)."""

    def __init__(self, length: int):
        self._len = int(length)

    def __len__(self) -> int:
        return self._len

    def __getitem__(self, idx: int) -> int:
        return idx


class IndexCollateLoader:
    """Iterable of minibatches; ``len()`` = minibatches per epoch; ``sampler`` / ``batch_size`` / ``dataset`` like a DataLoader."""

    def __init__(self, collector, trajectories_fn: Callable[[], Any], data_len: int, batch_size: int, sampler: MultiPassRandomSampler):
        self.collector, self._trajectories_fn = collector, trajectories_fn
        self.dataset, self.batch_size, self.sampler = IndexDataset(data_len), int(batch_size), sampler

    def __len__(self) -> int:
        return len(self.sampler) // self.batch_size

    def __iter__(self) -> Iterator[Any]:
        stream = self.sampler.passes().reshape(-1)              # (num_passes * data_len,) int64, on the sampler's device
        for lo in range(0, stream.numel(), self.batch_size):
            yield self.collector.slice_trajectories(self._trajectories_fn(), stream[lo:lo + self.batch_size])


def build_index_collate_loader_from_collector(*, collector, trajectories: Optional[Any] = None,
                                              trajectories_getter: Optional[Callable[[], Any]] = None, batch_size: int, num_passes: int,
                                              generator: Optional[torch.Generator] = None, num_workers: int = 0, pin_memory: bool = False,
                                              persistent_workers: bool = False, prefetch_factor: Optional[int] = None) -> IndexCollateLoader:
    """reference utils/dataloaders.py:20-77.  ``trajectories_getter`` is re-read for every minibatch, so the loader built once at the
    start of training serves every later rollout.  Worker / pinning options are accepted and must stay at their defaults: the
    rollout lives in HBM."""
    traj = trajectories_getter() if trajectories_getter is not None else trajectories
    if traj is None:
        raise ValueError("Either 'trajectories' or 'trajectories_getter' must be provided")
    if num_workers or pin_memory or persistent_workers or prefetch_factor is not None:
        raise ValueError("num_workers / pin_memory / persistent_workers / prefetch_factor do not apply: the rollout is device-resident")
    data_len = len(traj.observations)
    if data_len % int(batch_size) != 0:
        raise ValueError(f"Batch size must divide rollout size exactly: data_len={data_len}, batch_size={batch_size}. "
                         "Choose a batch_size (or fraction of n_envs*n_steps) that evenly divides the rollout.")
    device = generator.device if generator is not None else "cpu"
    sampler = MultiPassRandomSampler(data_len=data_len, num_passes=num_passes, generator=generator, device=device)
    getter = trajectories_getter if trajectories_getter is not None else (lambda: traj)
    return IndexCollateLoader(collector, getter, data_len, batch_size, sampler)


def build_dummy_loader(*, n_samples: int = 1, sample_dim: int = 1, batch_size: int = 1, num_workers: int = 0):
    """reference utils/dataloaders.py:12-17: a trivial loader for frameworks that insist on one (val / test stages)."""
    zeros = torch.zeros(n_samples, sample_dim)
    return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(zeros, zeros.clone()), batch_size=batch_size, num_workers=num_workers)
