"""PPOAgent on the CUDA engine (reference: agents/ppo/ppo_agent.py:9-152).

``losses_for_batch`` keeps its signature and return shape ({"loss", "early_stop_epoch"}) but is ONE fused kernel: gather
by env-major sample id, MLP forward, clipped surrogate + clipped value loss + entropy, every metric of the reference
(clip fractions, explained variance, KL / approx-KL, activation statistics) and the backward pass into the flat
gradient buffer."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _native as N
from ..base_agent import BaseAgent, EngineLoss


class PPOAgent(BaseAgent):
    def __init__(self, config, **kw):
        super().__init__(config, **kw)
        self.clip_range_vf = config.clip_range_vf

    def _step_moments(self):
        return int(self.config.normalize_advantages == "batch"), 0

    def _launch_step(self, b, *, defer: bool, moments=None) -> N.GsFinish:
        cfg = self.config
        model = self.policy_model
        mlp = N.mlp_struct(model)
        hp = N.GsPpoHparams()
        hp.clip_range, hp.clip_range_vf = float(self.clip_range), float(self.clip_range_vf)
        hp.vf_coef, hp.ent_coef = float(self.vf_coef), float(self.ent_coef)
        hp.normalize_adv = int(cfg.normalize_advantages == "batch")
        hp.track_activations = int(bool(getattr(cfg, "track_activations", True)))
        b.struct.defer_reduce = 2 if defer else 0          # 2: the agent's workspace is kept clean (base_agent: torch.zeros)
        adv_mom = moments[0:3] if moments is not None else None     # None on one rank: taken in the step's own gather pass
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_ppo_step(C.byref(mlp), C.byref(b.struct), C.byref(hp), N.ptr(adv_mom), N.ptr(model.flat_grads),
                                        N.ptr(self._metrics_dev), N.ptr(self._workspace), self._ws_bytes, N.stream()))
        fin = N.GsFinish()
        fin.algo, fin.track_activations, fin.normalize_adv, fin.normalize_ret = 0, hp.track_activations, hp.normalize_adv, 0
        fin.vf_coef, fin.ent_coef = hp.vf_coef, hp.ent_coef
        return fin

    def losses_for_batch(self, batch, batch_idx):
        cfg = self.config
        b = self._as_engine_batch(batch)
        # sharded minibatch: statistics of the GLOBAL minibatch (gather pass + local moments, all-reduce) -> W-invariant update
        self._launch_step(b, defer=False, moments=self._global_moments(b))
        early_stop = False
        if cfg.target_kl is not None:       # the only per-minibatch host sync, and only when KL early stop is enabled
            early_stop = float(self._metrics_dev[N.M["opt/ppo/approx_kl"]].item()) > float(cfg.target_kl)
            self.metrics_recorder.record("train", {"opt/ppo/kl_stop_triggered": 1 if early_stop else 0})    # ppo_agent.py:140
        return dict(loss=EngineLoss(self._metrics_dev), early_stop_epoch=early_stop)

    losses_for_batch._engine_native = True   # the fused step tail (BaseAgent._fused_training_step) may stand in for it

    def pop_epoch_metrics(self):
        out = super().pop_epoch_metrics()
        out.setdefault("opt/ppo/kl_stop_triggered", 0.0)        # without target_kl no minibatch ever triggers (reference: constant 0)
        return out
