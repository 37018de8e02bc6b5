"""REINFORCEAgent on the CUDA engine (reference: agents/reinforce/reinforce_agent.py:8-88).

loss = -(logp * target).mean() + ent_coef * (-entropy.mean()), target = returns or advantages (optionally batch
normalised), fused with the MLP forward / backward like PPO.  The reference reads ``config.normalize_advantages`` which
its REINFORCEConfig never declares; here the field exists and defaults to "off"."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _native as N
from ...utils.distributed import allreduce_moments
from ..base_agent import BaseAgent, EngineLoss


class REINFORCEAgent(BaseAgent):
    def __init__(self, config, **kw):
        super().__init__(config, **kw)
        self._mom = torch.zeros(6, dtype=torch.float64, device=self.device)   # [0:3] adv, [3:6] ret: {sum, sumsq, count}

    def losses_for_batch(self, batch, batch_idx):
        cfg = self.config
        if cfg.policy_targets not in ("returns", "advantages"):
            raise ValueError(f"Invalid policy targets: {cfg.policy_targets}")
        b = self._as_engine_batch(batch)
        model = self.policy_model
        mlp = N.mlp_struct(model)
        hp = N.GsReinforceHparams()
        hp.ent_coef = float(self.ent_coef)
        hp.policy_targets = 0 if cfg.policy_targets == "returns" else 1
        hp.normalize_returns = int(cfg.normalize_returns == "batch")
        hp.normalize_adv = int(getattr(cfg, "normalize_advantages", "off") == "batch")
        hp.track_activations = int(bool(getattr(cfg, "track_activations", True)))
        L = N.lib()
        with torch.cuda.device(self.device):
            st = N.stream()
            ret_mom = adv_mom = None             # one rank: the step takes the minibatch moments itself, in its gather pass
            if (hp.normalize_returns or hp.normalize_adv) and self.world_size > 1:
                # sharded minibatch: gather pass + local moments ([0:3] adv, [3:6] ret), all-reduce, then the rest of the step
                N.check(L.gs_batch_prepare(C.byref(mlp), C.byref(b.struct), int(hp.normalize_adv), int(hp.normalize_returns),
                                           N.ptr(self._mom), N.ptr(self._workspace), self._ws_bytes, st))
                allreduce_moments(self._mom, self.world_size)
                b.struct.prepared = 1
                adv_mom, ret_mom = self._mom[0:3], self._mom[3:6]
            N.check(L.gs_reinforce_step(C.byref(mlp), C.byref(b.struct), C.byref(hp), N.ptr(ret_mom), N.ptr(adv_mom),
                                        N.ptr(model.flat_grads), N.ptr(self._metrics_dev), N.ptr(self._workspace), self._ws_bytes, st))
        return dict(loss=EngineLoss(self._metrics_dev), early_stop_epoch=False)
