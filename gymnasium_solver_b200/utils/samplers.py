"""Minibatch index generation for the update phase.

Mirrors the reference's utils/samplers.py:7-37 (MultiPassRandomSampler: ``num_passes`` independent permutations of the
rollout, concatenated) without ever materialising ``n_epochs * N * T`` Python ints: on the engine's fast path the
permutation is a keyed bijection evaluated inside the update kernel (csrc/common.cuh: feistel_permute); this module
holds the host-side mirror of that bijection (for tests / inspection) and a drop-in Sampler for the compat path.
"""
from __future__ import annotations

from typing import Iterator, Optional

import torch

_M32 = 0xFFFFFFFF


def _mix32(x: torch.Tensor) -> torch.Tensor:
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def feistel_permutation(length: int, key: int, device="cpu") -> torch.Tensor:
    """The bijection of [0, length) the update kernel evaluates per sample (int64 tensor perm[i] = pi_key(i)).

    4-round alternating Feistel network on exactly ceil(log2(length)) bits (halves may differ by one bit) plus cycle walking;
    same arithmetic as ``gs::feistel_permute`` in csrc/common.cuh (uint32 lanes emulated in int64).  A power-of-two length
    needs no cycle walking.
    """
    bits = 2
    while (1 << bits) < length:
        bits += 1
    lb = bits >> 1
    rb = bits - lb
    lmask, rmask = (1 << lb) - 1, (1 << rb) - 1
    k0, k1 = key & _M32, (key >> 32) & _M32
    x = torch.arange(length, dtype=torch.int64, device=device)
    pending = torch.ones(length, dtype=torch.bool, device=device)
    while bool(pending.any()):
        cur = x[pending]
        L, R = (cur >> rb) & lmask, cur & rmask
        L = L ^ (_mix32((R * 0x9E3779B1 + k0) & _M32) & lmask)
        R = R ^ (_mix32((L * 0x85EBCA6B + k1) & _M32) & rmask)
        L = L ^ (_mix32((R * 0xC2B2AE35 + (k0 ^ 0x27D4EB2F)) & _M32) & lmask)
        R = R ^ (_mix32((L * 0x165667B1 + (k1 ^ 0x9E3779B9)) & _M32) & rmask)
        x[pending] = (L << rb) | R
        pending = x >= length
    return x


class MultiPassRandomSampler(torch.utils.data.Sampler):
    """Drop-in for the reference sampler (utils/samplers.py:7-37): ``num_passes`` independent permutations.

    Same constructor / ``set_epoch`` / ``__len__``; permutations come from ``argsort(rand)`` on ``device`` and are yielded
    as one int64 tensor per pass through ``passes()`` (the engine consumes tensors; ``__iter__`` keeps the int protocol).
    """

    def __init__(self, data_len: int, num_passes: int, generator: Optional[torch.Generator] = None, device="cpu") -> None:
        if data_len <= 0:
            raise ValueError("data_len must be > 0")
        if num_passes <= 0:
            raise ValueError("num_passes must be > 0")
        self.data_len, self.num_passes = int(data_len), int(num_passes)
        self.device = torch.device(device)
        self.generator = generator or torch.Generator(device=self.device)
        self._base_seed = int(torch.initial_seed())

    def set_epoch(self, epoch: int) -> None:
        self.generator.manual_seed(self._base_seed + int(epoch))

    def passes(self) -> torch.Tensor:
        scores = torch.rand((self.num_passes, self.data_len), generator=self.generator, device=self.generator.device)
        return torch.argsort(scores, dim=1)

    def __iter__(self) -> Iterator[int]:
        return iter(self.passes().reshape(-1).tolist())

    def __len__(self) -> int:
        return self.data_len * self.num_passes
