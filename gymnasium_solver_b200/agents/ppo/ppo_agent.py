"""PPOAgent on the CUDA engine (reference: agents/ppo/ppo_agent.py:9-152).

``losses_for_batch`` keeps its signature and return shape ({"loss", "early_stop_epoch"}) but is ONE fused kernel: gather
by env-major sample id, MLP forward, clipped surrogate + clipped value loss + entropy, every metric of the reference
(clip fractions, explained variance, KL / approx-KL, activation statistics) and the backward pass into the flat
gradient buffer."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _native as N
from ...utils.distributed import allreduce_moments
from ..base_agent import BaseAgent, EngineLoss


class PPOAgent(BaseAgent):
    def __init__(self, config, **kw):
        super().__init__(config, **kw)
        self.clip_range_vf = config.clip_range_vf
        self._adv_mom = torch.zeros(6, dtype=torch.float64, device=self.device)   # {sum, sumsq, count} of adv (and of ret, unused)

    def losses_for_batch(self, batch, batch_idx):
        cfg = self.config
        b = self._as_engine_batch(batch)
        model = self.policy_model
        mlp = N.mlp_struct(model)
        hp = N.GsPpoHparams()
        hp.clip_range, hp.clip_range_vf = float(self.clip_range), float(self.clip_range_vf)
        hp.vf_coef, hp.ent_coef = float(self.vf_coef), float(self.ent_coef)
        hp.normalize_adv = int(cfg.normalize_advantages == "batch")
        hp.track_activations = int(bool(getattr(cfg, "track_activations", True)))
        L = N.lib()
        with torch.cuda.device(self.device):
            st = N.stream()
            adv_mom = None                       # one rank: the step takes the minibatch moments itself, in its gather pass
            if hp.normalize_adv and self.world_size > 1:
                # sharded minibatch: gather pass + local moments, all-reduce them (statistics of the GLOBAL minibatch: the update is
                # W-invariant), then the rest of the step on the offsets the gather pass left in the workspace
                N.check(L.gs_batch_prepare(C.byref(mlp), C.byref(b.struct), 1, 0, N.ptr(self._adv_mom), N.ptr(self._workspace),
                                           self._ws_bytes, st))
                allreduce_moments(self._adv_mom, self.world_size)
                b.struct.prepared = 1
                adv_mom = self._adv_mom
            N.check(L.gs_ppo_step(C.byref(mlp), C.byref(b.struct), C.byref(hp), N.ptr(adv_mom), N.ptr(model.flat_grads),
                                  N.ptr(self._metrics_dev), N.ptr(self._workspace), self._ws_bytes, st))
        early_stop = False
        if cfg.target_kl is not None:       # the only per-minibatch host sync, and only when KL early stop is enabled
            early_stop = float(self._metrics_dev[N.M["opt/ppo/approx_kl"]].item()) > float(cfg.target_kl)
        return dict(loss=EngineLoss(self._metrics_dev), early_stop_epoch=early_stop)
