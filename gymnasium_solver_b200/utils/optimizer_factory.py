"""Optimizer construction (reference: utils/optimizer_factory.py:6-29): torch defaults, only lr is set.

``EngineAdam`` is torch.optim.Adam's update rule on the model's FLAT parameter vector: one kernel (``gs_adam_step``), or no
kernel of its own at all when the agent fuses it into the step tail (``gs_update_finish``: partial reduction -> gradient
exchange -> clip -> Adam).  Its ``state_dict()`` has torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` /
``exp_avg_sq``), so checkpoints are interchangeable with the reference's."""
from __future__ import annotations

import torch

from .. import _native as N


class EngineAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        if getattr(model, "flat_params", None) is None:
            raise N.EngineError("EngineAdam needs a model with flat parameter storage (flatten_parameters_())")
        params = list(model.parameters())
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                        differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.flat_params, self.flat_grads = model.flat_params, model.flat_grads
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.flat_params.device)   # steps taken (device-resident)
        self._bind_state()

    def _bind_state(self) -> None:
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.exp_avg[off:off + n].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p)}
            off += n

    def adam_struct(self) -> N.GsAdam:
        g = self.param_groups[0]
        a = N.GsAdam()
        a.params_flat, a.exp_avg, a.exp_avg_sq = N.ptr(self.flat_params), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq)
        a.step_count = N.ptr(self.step_dev)
        a.lr, a.beta1, a.beta2, a.eps = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])
        return a

    @torch.no_grad()
    def step(self, closure=None):
        a = self.adam_struct()
        with torch.cuda.device(self.flat_params.device):
            N.check(N.lib().gs_adam_step(a.params_flat, N.ptr(self.flat_grads), a.exp_avg, a.exp_avg_sq, self.flat_params.numel(),
                                         a.step_count, a.lr, a.beta1, a.beta2, a.eps, N.stream()))
        return None

    def state_dict(self):
        step = float(self.step_dev.item())
        for st in self.state.values():
            st["step"] = torch.tensor(step)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)      # replaces the state tensors by copies: move them back into the flat buffers
        off, step = 0, 0.0
        for p in self.param_groups[0]["params"]:
            n, st = p.numel(), self.state.get(p, {})
            if "exp_avg" in st:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = float(st["step"])
            off += n
        self.step_dev.fill_(int(step))
        self._bind_state()


def build_optimizer(*, params, optimizer, lr: float, model=None, **extra) -> torch.optim.Optimizer:
    name = str(getattr(optimizer, "value", optimizer)).lower()
    if name == "adam" and model is not None and getattr(model, "flat_params", None) is not None and not extra:
        return EngineAdam(model, lr=lr)
    cls = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW}[name]
    params = list(params)
    if name in ("adam", "adamw") and params and all(p.is_cuda for p in params):
        extra.setdefault("fused", True)      # same update rule as the default implementation, one kernel launch
    return cls(params, lr=lr, **extra)
