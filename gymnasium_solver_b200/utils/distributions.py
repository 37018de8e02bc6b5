"""MaskedCategorical (reference: utils/distributions.py:8-82): a Categorical whose entropy ignores -inf logits."""
from __future__ import annotations

import torch
from torch.distributions import Categorical


class MaskedCategorical(Categorical):
    def __init__(self, logits=None, probs=None, validate_args=None):
        if logits is not None:
            self._mask = torch.isfinite(logits)
        elif probs is not None:
            self._mask = probs > 0
        else:
            raise ValueError("Either logits or probs must be specified")
        super().__init__(logits=logits, probs=probs, validate_args=validate_args)

    def entropy(self):
        p = self.probs
        log_p = torch.where(self._mask, torch.log(p + 1e-8), torch.zeros_like(p))
        return -(p * log_p).sum(dim=-1)

    def log_prob(self, value):
        out = super().log_prob(value)
        if __debug__:
            picked = self._mask[torch.arange(value.shape[0], device=value.device), value.long()]
            if not bool(picked.all()):
                raise ValueError(f"Sampled {int((~picked).sum())} invalid (masked) actions. This should not happen - check sampling logic.")
        return out
