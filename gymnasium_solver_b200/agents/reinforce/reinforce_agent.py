"""REINFORCEAgent on the CUDA engine (reference: agents/reinforce/reinforce_agent.py:8-88).

loss = -(logp * target).mean() + ent_coef * (-entropy.mean()), target = returns or advantages (optionally batch
normalised), fused with the MLP forward / backward like PPO.  The reference reads ``config.normalize_advantages`` which
its REINFORCEConfig never declares; here the field exists and defaults to "off"."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _native as N
from ..base_agent import BaseAgent, EngineLoss


class REINFORCEAgent(BaseAgent):
    def _step_moments(self):
        cfg = self.config
        return int(getattr(cfg, "normalize_advantages", "off") == "batch"), int(cfg.normalize_returns == "batch")

    def _launch_step(self, b, *, defer: bool, moments=None) -> N.GsFinish:
        cfg = self.config
        if cfg.policy_targets not in ("returns", "advantages"):
            raise ValueError(f"Invalid policy targets: {cfg.policy_targets}")
        model = self.policy_model
        mlp = N.mlp_struct(model)
        hp = N.GsReinforceHparams()
        hp.ent_coef = float(self.ent_coef)
        hp.policy_targets = 0 if cfg.policy_targets == "returns" else 1
        hp.normalize_adv, hp.normalize_returns = self._step_moments()
        hp.track_activations = int(bool(getattr(cfg, "track_activations", True)))
        b.struct.defer_reduce = 2 if defer else 0          # 2: the agent's workspace is kept clean (base_agent: torch.zeros)
        adv_mom = moments[0:3] if moments is not None else None     # [0:3] adv, [3:6] ret; None: the step's own gather pass
        ret_mom = moments[3:6] if moments is not None else None
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_reinforce_step(C.byref(mlp), C.byref(b.struct), C.byref(hp), N.ptr(ret_mom), N.ptr(adv_mom),
                                              N.ptr(model.flat_grads), N.ptr(self._metrics_dev), N.ptr(self._workspace), self._ws_bytes,
                                              N.stream()))
        fin = N.GsFinish()
        fin.algo, fin.track_activations, fin.normalize_adv, fin.normalize_ret = 1, hp.track_activations, hp.normalize_adv, hp.normalize_returns
        fin.vf_coef, fin.ent_coef = 0.0, hp.ent_coef
        return fin

    def losses_for_batch(self, batch, batch_idx):
        b = self._as_engine_batch(batch)
        self._launch_step(b, defer=False, moments=self._global_moments(b))
        return dict(loss=EngineLoss(self._metrics_dev), early_stop_epoch=False)

    losses_for_batch._engine_native = True   # the fused step tail (BaseAgent._fused_training_step) may stand in for it
