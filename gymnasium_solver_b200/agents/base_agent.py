"""BaseAgent on the CUDA engine (reference: agents/base_agent.py).

Keeps the reference's agent surface — ``build_env`` / ``build_models`` / ``build_rollout_collector`` / ``losses_for_batch`` /
``training_step`` / ``_backpropagate_and_step`` / ``configure_optimizers`` / ``learn`` / ``save_checkpoint`` /
``load_checkpoint`` and the hook order of one "epoch" = one rollout + ``n_epochs`` passes over it (reference
agents/base_agent.py:253-366) — without PyTorch-Lightning: ``learn()`` drives the hooks itself.

Engine mapping of one epoch:
  collect                       1 fused collect launch + 1 target kernel                     (RolloutCollector.collect)
  for pass, minibatch           gs_ppo_step / gs_reinforce_step with defer_reduce (gather pass + forward + loss + backward, per-CTA
                                partial gradients) -> gs_update_finish (ONE launch: ordered reduction, gradient exchange over
                                NVLink peer memory, metrics, global-norm clip, Adam on the flat parameter vector)
                                [generic path for overridden losses / other optimizers / KL early stop: gs_*_step ->
                                 NCCL all-reduce -> gs_clip_grad_norm -> optimizer.step()]
Metrics stay on device as a running sum and reach the host once per epoch.
Data parallel: one process per GPU owns ``n_envs / world_size`` envs (global env ids keep RNG streams W-invariant); the
only data-path exchange is the gradient mean per minibatch (inside gs_update_finish) + ONE all-reduce per rollout of the
minibatch advantage moments of all its minibatches, so "batch" normalisation means the GLOBAL minibatch.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import sys
import threading
import time
from pathlib import Path
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from ..utils import distributed as D
from ..utils.distributed import PeerGroup, allreduce_moments, average_gradients, shard_spec
from .hyperparameter_mixin import HyperparameterMixin, RunConfigFile  # noqa: F401  (RunConfigFile: re-exported for launchers)
from ..utils.environment import build_env_from_config
from ..utils.optimizer_factory import EngineAdam, build_optimizer
from ..utils.policy_factory import build_policy_from_env_and_config
from ..utils.rollout_collector import DeviceTrajectory, RolloutCollector
from ..utils.schedules import position_to_env_steps, progress_fraction, scheduled_value
from ..utils.timings_tracker import TimingsTracker

STAGES = ["train", "val", "test"]
_SPLITMIX = 0x9E3779B97F4A7C15


def _mix64(x: int) -> int:
    x = (x + _SPLITMIX) & (2**64 - 1)
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return x ^ (x >> 31)


class EngineBatch:
    """One minibatch for the update kernels: a gs_batch_t over the rollout's time-major arrays plus what keeps it alive."""

    def __init__(self, struct: N.GsBatch, keep, n: int):
        self.struct, self.keep, self.n = struct, keep, int(n)
        self.moments = None          # device double[6] {adv, ret} x {sum, sumsq, count}, already reduced over ranks (prepared batch)

    def __len__(self):
        return self.n


class EngineLoss:
    """What ``losses_for_batch`` returns as "loss": the gradients already sit in ``policy_model.flat_grads`` (forward,
    loss and backward are one kernel), so ``backward()`` is a no-op; ``item()`` reads the scalar (one D2H copy)."""

    def __init__(self, metrics: torch.Tensor):
        self.metrics = metrics

    def backward(self):
        return None

    def detach(self):
        return self.metrics[N.M["opt/loss/total"]]

    def item(self) -> float:
        return float(self.metrics[N.M["opt/loss/total"]].item())

    __float__ = item


class MetricsRecorder:
    """Per-epoch metric aggregation (reference: utils/metrics_recorder.py + metrics_buffer.py): host scalars are averaged
    per epoch; the engine's per-minibatch metric vectors are summed ON DEVICE and folded in at epoch end."""

    def __init__(self):
        self._host: Dict[str, Dict[str, list]] = {}
        self.history: list[Dict[str, Any]] = []

    def record(self, stage: str, metrics: Dict[str, Any]) -> None:
        bucket = self._host.setdefault(stage, {})
        for k, v in metrics.items():
            if v is None:
                continue
            v = float(v.item()) if isinstance(v, torch.Tensor) else float(v)
            assert math.isfinite(v), f"metric {k} is not finite: {v}"
            bucket.setdefault(k, []).append(v)

    def epoch_means(self, stage: str, reset: bool = True) -> Dict[str, float]:
        bucket = self._host.get(stage, {})
        out = {k: sum(v) / len(v) for k, v in bucket.items() if v}
        if reset:
            self._host[stage] = {}
        return out

    # names the reference's callbacks call (utils/metrics_recorder.py)
    def reset_epoch(self, stage: str) -> None:
        self._host[stage] = {}

    def compute_epoch_means(self, stage: str) -> Dict[str, float]:
        return self.epoch_means(stage, reset=False)

    def update_history(self, metrics: Dict[str, Any]) -> None:
        self.history.append(dict(metrics))


class BaseAgent(HyperparameterMixin, nn.Module):
    def __init__(self, config, *, device=None, rank: Optional[int] = None, world_size: Optional[int] = None):
        super().__init__()
        self.config = config
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else int(world_size)
        if not torch.cuda.is_available():
            raise N.EngineError("the b200 engine needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.policy_model = None
        self._envs: Dict[str, Any] = {}
        self._rollout_collectors: Dict[str, RolloutCollector] = {}
        self._trajectories = None
        self._early_stop_epoch = False
        self._early_stop_reason = ""
        self._optimizer = None
        self.current_epoch = 0
        self.should_stop = False
        self.best_eval_reward = -float("inf")
        # asynchronous evaluation (reference agents/base_agent.py:38-40, 65-70)
        self._eval_models: Dict[str, Any] = {}
        self._async_eval_thread: Optional[threading.Thread] = None
        self._async_eval_metrics: Dict[str, Any] = {}
        self._async_eval_lock = threading.Lock()
        self._async_eval_shutdown = threading.Event()
        self._async_eval_pending_epoch: Optional[int] = None
        self._async_eval_running_epoch: Optional[int] = None
        self._async_eval_error: Optional[BaseException] = None
        self._eval_stream = None
        self._eval_snapshot = None          # weights of the epoch to evaluate next, written on the training stream
        self._eval_snapshot_event = None

        # schedulable hyperparameters (reference agents/base_agent.py:72-77)
        self.policy_lr = config.policy_lr
        self.clip_range = getattr(config, "clip_range", None)
        self.vf_coef = getattr(config, "vf_coef", None)
        self.ent_coef = config.ent_coef
        self.n_epochs = config.n_epochs
        self.metrics_recorder = MetricsRecorder()
        self.timings = TimingsTracker()

        self.shard = shard_spec(int(config.n_envs), int(config.batch_size), self.rank, self.world_size)
        self.local_n_envs, self.local_batch_size = self.shard.n_envs, self.shard.batch_size

        for stage in STAGES:
            self.build_env(stage)
        self.build_models()
        for stage in STAGES:
            self.build_rollout_collector(stage)

        # device scratch of the update path
        mlp = N.mlp_struct(self.policy_model)
        self._ws_bytes = N.lib().gs_update_workspace_bytes(C.byref(mlp), self.device.index, int(self.local_batch_size))
        if self._ws_bytes <= 0:
            raise N.EngineError(N.lib().gs_last_error().decode())
        self._workspace = torch.zeros(self._ws_bytes, dtype=torch.uint8, device=self.device)   # clean: steps run with defer_reduce = 2 (no memset node)
        self._metrics_dev = torch.zeros(N.N_METRICS, dtype=torch.float64, device=self.device)
        self._metrics_sum = torch.zeros(N.N_METRICS, dtype=torch.float64, device=self.device)
        self._metrics_n = 0
        self._mom_scratch = torch.zeros(6, dtype=torch.float64, device=self.device)
        # NVLink peer exchange for the gradient mean (gs_update_finish).  grad_allreduce: "peer" (required), "nccl" (generic path:
        # torch.distributed all-reduce between the step and clip / optimizer kernels), "auto" (peer, else nccl on every rank).
        self._peer = None
        self.grad_allreduce_mode = "none" if self.world_size == 1 else "nccl"
        mode = getattr(config, "grad_allreduce", "auto")
        if mode not in ("auto", "peer", "nccl"):
            raise ValueError(f"grad_allreduce must be auto, peer or nccl, got {mode!r}")
        if self.world_size > 1 and bool(getattr(config, "fused_update", True)) and mode in ("auto", "peer"):
            try:
                self._peer = PeerGroup(self.rank, self.world_size, int(self.policy_model.flat_params.numel()), self.device)
                self.grad_allreduce_mode = "peer"
            except N.EngineError as e:
                if mode == "peer":
                    raise
                if self.rank == 0:
                    print(f"[gymnasium_solver_b200] peer gradient exchange unavailable ({e}); using the NCCL all-reduce path", file=sys.stderr, flush=True)

    # ------------------------------------------------------------------------------------------------ construction
    def build_env(self, stage: str, **kwargs):
        cfg = self.config
        seed = {"train": cfg.seed_train, "val": cfg.seed_val, "test": cfg.seed_test}[stage]
        with torch.cuda.device(self.device):
            env = build_env_from_config(cfg, seed=seed, n_envs=self.local_n_envs, device=self.device,
                                        env_id_offset=self.shard.env_id_offset, **kwargs)
        self._envs[stage] = env
        return env

    def get_env(self, stage: str):
        return self._envs[stage]

    def build_models(self):
        """reference agents/base_agent.py:103-106 — then move to the GPU and flatten parameter / grad storage."""
        torch.manual_seed(int(self.config.seed))        # identical initial weights on every rank
        model = build_policy_from_env_and_config(self.get_env("train"), self.config)
        model.to(self.device)
        model.flatten_parameters_()
        self.policy_model = model

    def build_rollout_collector(self, stage: str):
        kw = self.config.get_rollout_collector_kwargs()
        # val / test collectors read the SAME parameter storage: evaluation sees the current weights without copies.  With
        # eval_async the val collector gets its own copy of the network (reference agents/base_agent.py:197-213): the background
        # evaluation reads a snapshot while the training stream keeps updating the live weights.
        model = self.policy_model
        if stage == "val" and bool(getattr(self.config, "eval_async", False)):
            model = build_policy_from_env_and_config(self.get_env("train"), self.config)
            model.to(self.device)
            model.flatten_parameters_()
            model.flat_params.copy_(self.policy_model.flat_params)
            self._eval_models[stage] = model
        col = RolloutCollector(self.get_env(stage), model, store_next_obs=bool(self.config.store_next_obs),
                               rng_seed=_mix64(int(self.config.seed) * 3 + STAGES.index(stage)) >> 1, **kw)
        self._rollout_collectors[stage] = col
        return col

    def get_rollout_collector(self, stage: str) -> RolloutCollector:
        return self._rollout_collectors[stage]

    # ------------------------------------------------------------------------------------------------ to override
    def losses_for_batch(self, batch, batch_idx):
        raise NotImplementedError("Subclasses must implement losses_for_batch")

    # ------------------------------------------------------------------------------------------------ optimisation
    def configure_optimizers(self):
        return build_optimizer(params=self.policy_model.parameters(), optimizer=self.config.optimizer, lr=self.policy_lr,
                               model=self.policy_model)

    def optimizers(self):
        if self._optimizer is None:
            self._optimizer = self.configure_optimizers()
        return self._optimizer

    # _change_optimizers_lr / set_hyperparameter / _read_hyperparameters_from_run / _log_hyperparameters: HyperparameterMixin

    def _as_engine_batch(self, batch) -> EngineBatch:
        """Accept the engine's own minibatch descriptor or a reference-style gathered RolloutTrajectory (flat tensors)."""
        if isinstance(batch, EngineBatch):
            return batch
        f = lambda x, dt: x.detach().to(device=self.device, dtype=dt).contiguous()
        obs = f(batch.observations, torch.float32)
        B = obs.shape[0]
        arrs = dict(obs=obs.reshape(1, B, -1), actions=f(batch.actions, torch.int32).reshape(1, B),
                    logprobs=f(batch.logprobs, torch.float32).reshape(1, B), values=f(batch.values, torch.float32).reshape(1, B),
                    adv=f(batch.advantages, torch.float32).reshape(1, B), ret=f(batch.returns, torch.float32).reshape(1, B))
        return self._make_batch(arrs, 1, B, n=B)

    def _make_batch(self, tm: dict, T: int, n_envs: int, *, n: int, idx=None, perm_key: int = 0, perm_offset: int = 0,
                    perm_len: int = 0, idx_map=None) -> EngineBatch:
        b = N.GsBatch()
        b.T, b.N, b.obs_dim = T, n_envs, tm["obs"].shape[-1]
        b.obs, b.actions, b.logp_old = N.ptr(tm["obs"]), N.ptr(tm["actions"]), N.ptr(tm["logprobs"])
        b.values_old, b.adv, b.ret = N.ptr(tm["values"]), N.ptr(tm["adv"]), N.ptr(tm["ret"])
        b.n, b.idx, b.perm_key, b.perm_offset, b.perm_len = n, N.ptr(idx), perm_key, perm_offset, perm_len
        b.idx_map = N.ptr(idx_map)
        b.packed = N.ptr(tm.get("packed"))
        return EngineBatch(b, (tm, idx, idx_map), n)

    # ---- engine step: subclasses launch their kernel here ------------------------------------------------------------
    def _step_moments(self):
        """(want_adv, want_ret): which minibatch moments the step's "batch" normalisation needs."""
        return 0, 0

    def _launch_step(self, b: EngineBatch, *, defer: bool, moments=None) -> N.GsFinish:
        raise NotImplementedError

    def _prepare(self, b: EngineBatch, moments: torch.Tensor) -> None:
        """Gather pass + local moments of a sharded minibatch (the caller all-reduces ``moments``)."""
        want_adv, want_ret = self._step_moments()
        mlp = N.mlp_struct(self.policy_model)
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_batch_prepare(C.byref(mlp), C.byref(b.struct), want_adv, want_ret, N.ptr(moments), N.ptr(self._workspace),
                                             self._ws_bytes, N.stream()))
        b.struct.prepared = 1
        b.moments = moments

    def _global_moments(self, b: EngineBatch):
        """Moments of the GLOBAL minibatch for a sharded step: those prepared up front, else prepare + all-reduce now."""
        if b.moments is not None:
            return b.moments
        if self.world_size <= 1 or not any(self._step_moments()):
            return None
        if b.moments is None:
            self._prepare(b, self._mom_scratch)
            allreduce_moments(self._mom_scratch, self.world_size, peer=self._peer)
        return b.moments

    def _fused_step_ok(self, batch) -> bool:
        own = getattr(type(self).losses_for_batch, "_engine_native", False)    # a user override takes the generic path
        return (own and isinstance(batch, EngineBatch) and bool(getattr(self.config, "fused_update", True))
                and getattr(self.config, "target_kl", None) is None and isinstance(self.optimizers(), EngineAdam)
                and (self.world_size == 1 or self._peer is not None))

    def _fused_training_step(self, b: EngineBatch) -> None:
        """losses_for_batch + _backpropagate_and_step as two launches: the update kernel (per-CTA partials stay in the
        workspace) and gs_update_finish (reduction, NVLink gradient mean, metrics, clip, Adam)."""
        fin = self._launch_step(b, defer=True, moments=self._global_moments(b))
        model, opt = self.policy_model, self.optimizers()
        mlp, adam = N.mlp_struct(model), opt.adam_struct()
        fin.max_grad_norm = float(self.config.max_grad_norm) if self.config.max_grad_norm is not None else 0.0
        peer = self._peer.handle if self._peer is not None else None
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_update_finish(C.byref(mlp), C.byref(b.struct), C.byref(fin), N.ptr(model.flat_grads), C.byref(adam), peer,
                                             N.ptr(self._metrics_dev), N.ptr(self._metrics_sum), N.ptr(self._workspace), self._ws_bytes,
                                             N.stream()))
        self._metrics_n += 1

    def training_step(self, batch, batch_idx):
        """reference agents/base_agent.py:330-366."""
        if self._early_stop_epoch:
            return None
        if self._fused_step_ok(batch):
            self._fused_training_step(batch)
            return None
        result = self.losses_for_batch(batch, batch_idx)
        if result["early_stop_epoch"]:
            self._early_stop_epoch = True
            return None
        self._backpropagate_and_step(result["loss"])
        return None

    def _backpropagate_and_step(self, losses):
        """reference agents/base_agent.py:591-621: grads -> group norms (pre-clip) -> global-norm clip -> optimizer step.
        The gradients were produced by the fused kernel; with several ranks they are averaged over NVLink first."""
        opt = self.optimizers()
        model = self.policy_model
        average_gradients(model.flat_grads, self.world_size)
        mlp = N.mlp_struct(model)
        max_norm = float(self.config.max_grad_norm) if self.config.max_grad_norm is not None else 0.0
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_clip_grad_norm(C.byref(mlp), N.ptr(model.flat_grads), max_norm, N.ptr(self._metrics_dev), N.stream()))
        self._metrics_sum += self._metrics_dev
        self._metrics_n += 1
        opt.step()

    def train_dataloader(self):
        """reference agents/base_agent.py:253-283: collect the first rollout and build the index-collate loader over
        ``self._trajectories`` (re-read for every minibatch, so the one loader serves every later rollout).  For callers that drive
        ``training_step(batch, batch_idx)`` themselves the way Lightning drove the reference; ``learn()`` / ``fit()`` do not go through
        it (their update kernels gather in place).  The local shard's batch size is used: every rank feeds its own loader."""
        from ..utils.dataloaders import build_index_collate_loader_from_collector
        from ..utils.random import get_global_torch_generator

        assert self.current_epoch == 0 or getattr(self, "_resume_from_epoch", None) is not None, \
            "train_dataloader should only be called once at the start of training"
        col = self.get_rollout_collector("train")
        self._trajectories = col.collect()
        self._train_dataloader = build_index_collate_loader_from_collector(
            collector=col, trajectories_getter=lambda: self._trajectories, batch_size=self.local_batch_size,
            num_passes=self.config.n_epochs, generator=get_global_torch_generator(self.config.seed))
        return self._train_dataloader

    # ------------------------------------------------------------------------------------------------ one epoch
    def minibatches(self, traj: DeviceTrajectory, epoch_key: int):
        """n_epochs passes of rollout/batch minibatches (reference: MultiPassRandomSampler, utils/samplers.py:29-34).
        "device" shuffle: pass p is the keyed bijection pi_{key(epoch,p)} evaluated inside the kernel; "torch" shuffle:
        argsort(rand) index tensors like the reference."""
        total = len(traj)
        B = self.local_batch_size
        if total % B != 0:
            raise ValueError(f"Batch size must divide rollout size exactly: data_len={total}, batch_size={B}.")
        n_mb = total // B
        tm, idx_map = traj.tm, traj.tm.get("idx_map")
        for p in range(int(self.n_epochs)):
            if self.config.minibatch_shuffle == "torch":
                g = getattr(self, "_shuffle_gen", None)
                if g is None:
                    g = self._shuffle_gen = torch.Generator(device=self.device).manual_seed(int(self.config.seed))
                order = torch.argsort(torch.rand(total, generator=g, device=self.device))
                for k in range(n_mb):
                    idx = order[k * B:(k + 1) * B].contiguous()
                    yield p, k, self._make_batch(tm, traj.T, traj.n_envs, n=B, idx=idx, idx_map=idx_map)
            else:
                key = _mix64(epoch_key * 1315423911 + p)
                for k in range(n_mb):
                    yield p, k, self._make_batch(tm, traj.T, traj.n_envs, n=B, perm_key=key, perm_offset=k * B, perm_len=total, idx_map=idx_map)

    def _pack_rollout(self, traj: DeviceTrajectory) -> None:
        """One 64-byte record per sample (gs_rollout_pack: the bf16x3 layer-1 operand row + the sample's scalars) so the tensor-core
        update kernel copies a minibatch sample straight into its operand tile with one aligned access.  The 64x64, 128x128 and (relu)
        256x256 networks have that kernel (observations of up to 7 features); the buffer is reused across rollouts."""
        if not self._tensor_path(traj.tm["obs"].shape[-1]) or "packed" in traj.tm:
            return
        total = traj.T * traj.n_envs
        buf = getattr(self, "_packed_records", None)
        if buf is None or buf.numel() != total * 16:
            buf = self._packed_records = torch.empty(total * 16, dtype=torch.float32, device=self.device)
        full = self._make_batch(traj.tm, traj.T, traj.n_envs, n=total)
        with torch.cuda.device(self.device):
            N.check(N.lib().gs_rollout_pack(C.byref(full.struct), N.ptr(buf), N.stream()))
        traj.tm["packed"] = buf

    def _tensor_path(self, obs_dim: Optional[int] = None) -> bool:
        """Does csrc/update_f16.cu / update_wide.cu serve this network (mirrors update_kernels.cu::tensor_path_for)?  Those kernels read per-minibatch
        offset buffers and rollout records; everything else runs the FMA-pipe kernel."""
        hd = tuple(getattr(self.config, "hidden_dims", ()))
        if hd not in ((64, 64), (128, 128), (256, 256)):
            return False
        if hd == (256, 256) and str(getattr(self.config, "activation", None) or "relu").lower() != "relu":
            return False                # csrc/update_wide.cu keeps relu' as a bit mask; tanh at 256 x 256 runs the FMA-pipe kernel
        if os.environ.get("GS_UPDATE_IMPL", "tc") == "simt":
            return False
        if obs_dim is None:
            shape = getattr(self.get_env("train").single_observation_space, "shape", None) or (1,)
            obs_dim = int(shape[-1])
        return obs_dim <= 7

    def _prepare_all(self, batches) -> None:
        """Sharded minibatches: run every gather pass of the rollout now and exchange ALL minibatch moments in ONE all-reduce
        (instead of one per minibatch on the critical path of every step)."""
        n, B = len(batches), self.local_batch_size
        mom = getattr(self, "_mom_all", None)
        if mom is None or mom.shape[0] != n:
            mom = self._mom_all = torch.zeros(n, 6, dtype=torch.float64, device=self.device)
        offs = None
        if self._tensor_path():     # the tensor-core kernel reads translated offsets
            offs = getattr(self, "_offs_all", None)
            if offs is None or offs.shape != (n, B):
                offs = self._offs_all = torch.empty(n, B, dtype=torch.int32, device=self.device)
        for k, b in enumerate(batches):
            if offs is not None:
                b.struct.offsets = N.ptr(offs[k])
            self._prepare(b, mom[k])
        allreduce_moments(mom, self.world_size, peer=self._peer)

    def train_on_rollout(self, traj: DeviceTrajectory) -> None:
        self._early_stop_epoch = False
        self._pack_rollout(traj)
        key = _mix64(int(self.config.seed) ^ (self.current_epoch * 0x100000001B3))
        batches = [b for _, _, b in self.minibatches(traj, key)]
        fused = bool(batches) and self._fused_step_ok(batches[0])
        tensor_path = self._tensor_path()
        if fused and self.world_size > 1 and any(self._step_moments()):
            per_pass = max(1, len(batches) // max(1, int(self.n_epochs)))
            if tensor_path and per_pass >= 2 and len(batches) >= 2 * per_pass:
                return self._train_pipelined_sharded(batches, per_pass)
            self._prepare_all(batches)
        elif fused and tensor_path and len(batches) > 1:
            return self._train_pipelined(batches)
        for k, batch in enumerate(batches):
            self.training_step(batch, k)

    def _train_pipelined_sharded(self, batches, per_pass: int) -> None:
        """Several ranks, tensor-core kernel: the gather passes of pass p+1 (sample-id translation + local minibatch moments) and
        the ONE all-reduce of that pass's moments run on a side stream while pass p is being trained, so only the first pass's
        preparation is exposed.  The gathers are queued one per training step (each runs in the shadow of a step tail, where one
        block works and the other SMs idle) rather than eight at once in front of an update kernel.  Offset buffers: a ring of
        two passes."""
        n, B = len(batches), self.local_batch_size
        n_chunks = (n + per_pass - 1) // per_pass
        st = getattr(self, "_pipe_sh", None)
        if st is None or st["mom"].shape[0] < n or st["offs"].shape != (2 * per_pass, B):
            st = self._pipe_sh = dict(side=torch.cuda.Stream(device=self.device),
                                      offs=torch.empty(2 * per_pass, B, dtype=torch.int32, device=self.device),
                                      mom=torch.zeros(n, 6, dtype=torch.float64, device=self.device),
                                      ready=[torch.cuda.Event() for _ in range(n_chunks)], done=[torch.cuda.Event() for _ in range(n_chunks)])
        side, main = st["side"], torch.cuda.current_stream(self.device)
        bounds = lambda c: (c * per_pass, min(n, (c + 1) * per_pass))

        def gather(c, k):                                           # minibatch k of pass c, on the side stream
            lo, hi = bounds(c)
            with torch.cuda.stream(side):
                if k == lo and c >= 2:
                    side.wait_event(st["done"][c - 2])              # pass c-2 has read this half of the offset ring
                b = batches[k]
                b.struct.offsets = N.ptr(st["offs"][(c & 1) * per_pass + (k - lo)])
                self._prepare(b, st["mom"][k])
                if k == hi - 1:
                    allreduce_moments(st["mom"][lo:hi], self.world_size, peer=self._peer)   # statistics of the GLOBAL minibatches of this pass
                    st["ready"][c].record(side)

        side.wait_stream(main)
        for k in range(*bounds(0)):
            gather(0, k)
        for c in range(n_chunks):
            lo, hi = bounds(c)
            nxt = list(range(*bounds(c + 1))) if c + 1 < n_chunks else []
            main.wait_event(st["ready"][c])
            for k in range(lo, hi):
                self.training_step(batches[k], k)
                # the gathers of the next pass are queued two per step from the pass's first step on, so they -- and the pass's moment
                # exchange, queued with the last of them -- are done by mid-pass on every rank: the exchange waits for the SLOWEST rank's
                # side stream, and with one gather per step it finished ~200 us after the pass boundary on 8 GPUs (measured: the per-pass
                # term of the 1 -> 8 GPU step-time difference)
                take = int(os.environ.get("GS_GATHERS_PER_STEP", "2"))
                for _ in range(take):
                    if nxt:
                        gather(c + 1, nxt.pop(0))
            while nxt:
                gather(c + 1, nxt.pop(0))
            st["done"][c].record(main)
        main.wait_stream(side)

    def _train_pipelined(self, batches) -> None:
        """One rank, tensor-core kernel: the gather pass of minibatch k+1 (sample-id translation + minibatch moments; independent
        of the weights) runs on a side stream while minibatch k is in its update kernel / step tail, so it leaves the
        critical path.  Two offset buffers; events order gather(k+2) after step(k) and step(k) after gather(k)."""
        n, B = len(batches), self.local_batch_size
        st = getattr(self, "_pipe", None)
        if st is None or st["mom"].shape[0] < n or st["offs"].shape[1] != B:
            st = self._pipe = dict(side=torch.cuda.Stream(device=self.device), offs=torch.empty(2, B, dtype=torch.int32, device=self.device),
                                   mom=torch.zeros(n, 6, dtype=torch.float64, device=self.device),
                                   gathered=[torch.cuda.Event() for _ in range(n)], stepped=[torch.cuda.Event() for _ in range(n)])
        side, main = st["side"], torch.cuda.current_stream(self.device)
        want = any(self._step_moments())

        def gather(k):
            b = batches[k]
            b.struct.offsets = N.ptr(st["offs"][k & 1])
            with torch.cuda.stream(side):
                if k >= 2:
                    side.wait_event(st["stepped"][k - 2])        # the update kernel of minibatch k-2 has read this offset buffer
                self._prepare(b, st["mom"][k])
                st["gathered"][k].record(side)
            if not want:
                b.moments = None

        side.wait_stream(main)                                      # the rollout's targets and packed records are complete
        gather(0)
        gather(1)
        for k, b in enumerate(batches):
            main.wait_event(st["gathered"][k])
            self.training_step(b, k)
            st["stepped"][k].record(main)
            if k + 2 < n:
                gather(k + 2)
        main.wait_stream(side)

    def train_one_rollout(self) -> DeviceTrajectory:
        """collect + targets + all minibatch passes: the unit the headline metric counts env-steps over."""
        col = self.get_rollout_collector("train")
        traj = col.collect()
        self._trajectories = traj
        self.train_on_rollout(traj)
        col.resolve_episodes_async()      # host-side episode bookkeeping while the passes run (no drain at the next collect)
        return traj

    def pop_epoch_metrics(self) -> Dict[str, float]:
        """Epoch means of the engine's per-minibatch metrics (one D2H copy)."""
        out = {}
        if self._metrics_n > 0:
            mean = (self._metrics_sum / self._metrics_n).cpu().numpy()
            out = {k: float(mean[i]) for i, k in enumerate(N.METRIC_KEYS)}
            self._metrics_sum.zero_()
            self._metrics_n = 0
        out.update(self.metrics_recorder.epoch_means("train"))
        return out

    # ------------------------------------------------------------------------------------------------ schedules
    def calc_training_progress(self) -> float:
        if self.config.max_env_steps is None:
            return 0.0
        total = self.get_rollout_collector("train").total_steps * self.world_size
        return max(0.0, min(total / float(self.config.max_env_steps), 1.0))

    def _apply_schedules(self) -> None:
        """reference trainer_callbacks/hyperparameter_scheduler.py:76-113 + utils/schedule_resolver.py (utils/schedules.py)."""
        cfg = self.config
        steps = self.get_rollout_collector("train").total_steps * self.world_size
        for param in ("policy_lr", "ent_coef", "vf_coef", "clip_range", "clip_range_vf"):
            kind = getattr(cfg, f"{param}_schedule", None)
            if not kind:
                continue
            v0, v1 = getattr(cfg, f"{param}_schedule_start_value"), getattr(cfg, f"{param}_schedule_end_value")
            s0 = position_to_env_steps(getattr(cfg, f"{param}_schedule_start", None), param=param, default_to_max=False,
                                       max_env_steps=cfg.max_env_steps)
            s1 = position_to_env_steps(getattr(cfg, f"{param}_schedule_end", None), param=param, default_to_max=True,
                                       max_env_steps=cfg.max_env_steps)
            frac = progress_fraction(float(steps), s0, s1)
            warm = float(getattr(cfg, f"{param}_schedule_warmup", 0.0) or 0.0)
            self.set_hyperparameter(param, scheduled_value(kind, float(v0), float(v1), frac, warm))

    # ------------------------------------------------------------------------------------------------ fit loop
    def learn(self, *, log_fn=None, max_epochs: Optional[int] = None) -> Dict[str, Any]:
        """Built-in fit loop with the reference's hook order."""
        cfg = self.config
        col = self.get_rollout_collector("train")
        t0 = time.time()
        self.on_fit_start()
        history = []
        max_epochs = max_epochs if max_epochs is not None else cfg.max_epochs
        reason = ""
        try:
            while True:
                # on_train_epoch_start: live hyper-parameters (:297-302), then the budget check before collecting (:306-320)
                self._read_hyperparameters_from_run()
                self._log_hyperparameters()
                if cfg.max_env_steps is not None:
                    cur = col.total_steps * self.world_size
                    nxt = int(cfg.n_envs) * int(cfg.n_steps)
                    if cur + nxt > cfg.max_env_steps:
                        reason = f"'train/cnt/total_env_steps': {cur} + {nxt} would exceed {int(cfg.max_env_steps)}."
                        break
                if max_epochs is not None and self.current_epoch >= max_epochs:
                    reason = f"max_epochs={max_epochs} reached."
                    break
                self.train_one_rollout()
                # on_train_epoch_end: metrics, schedules, early stopping
                row = {"epoch": self.current_epoch, "time_s": time.time() - t0}
                row.update({f"train/{k}": v for k, v in col.get_metrics().items() if k != "action_dist"})
                row.update({f"train/{k}": v for k, v in self.pop_epoch_metrics().items()})
                row["train/cnt/total_env_steps"] = col.total_steps * self.world_size
                self._apply_schedules()
                thr = cfg.early_stop_on_train_threshold
                if thr and "train/roll/ep_rew/mean" in row:
                    limit = self.get_env("train").get_return_threshold() if thr is True else float(thr)
                    if row["train/roll/ep_rew/mean"] >= limit:
                        reason = f"train/roll/ep_rew/mean {row['train/roll/ep_rew/mean']:.2f} >= {limit}"
                        self.should_stop = True
                if cfg.eval_freq_epochs and (self.current_epoch + 1) % int(cfg.eval_freq_epochs) == 0 \
                        and self.current_epoch + 1 >= int(cfg.eval_warmup_epochs):
                    self.metrics_recorder.reset_epoch("val")
                    ev = self.validation_epoch()        # synchronous, or the latest finished background evaluation (eval_async)
                    row.update({f"val/{k}": v for k, v in ev.items()})
                    mean = ev.get("roll/ep_rew/mean")
                    if mean is not None:
                        thr = cfg.early_stop_on_eval_threshold
                        if thr:
                            limit = cfg.reward_threshold if cfg.reward_threshold is not None else self.get_env("val").get_return_threshold()
                            limit = limit if thr is True else float(thr)
                            if mean >= limit:
                                reason = f"val/roll/ep_rew/mean {mean:.2f} >= {limit}"
                                self.should_stop = True
                history.append(row)
                if log_fn is not None and self.rank == 0:
                    log_fn(row)
                self.current_epoch += 1
                # every rank evaluates its own env shard: the stop decision is taken together (any rank over the threshold stops all),
                # the budget / max_epochs exits above are functions of the global step count and already identical everywhere
                if self.world_size > 1 and (cfg.early_stop_on_train_threshold or cfg.early_stop_on_eval_threshold):
                    self.should_stop = D.agree_any(self.should_stop, self.device, self.world_size)
                    if self.should_stop and not reason:
                        reason = "another rank crossed the early-stopping threshold"
                if self.should_stop:
                    break
        except BaseException as exc:
            self.on_exception(None, self, exc)
            raise
        self._early_stop_reason = reason
        self.on_fit_end()
        return {"history": history, "stop_reason": reason, "epochs": self.current_epoch, "elapsed_s": time.time() - t0,
                "total_env_steps": col.total_steps * self.world_size, "best_eval_reward": self.best_eval_reward}

    # ------------------------------------------------------------------------------------------------ trainer-shell hooks
    _fit_t0 = 0.0
    trainer = None

    def on_fit_start(self) -> None:
        self._fit_t0 = time.time()
        self._async_eval_shutdown.clear()          # a new fit may evaluate in the background again (set by the previous on_fit_end)
        self.timings.start("on_fit_start", values=self._host_counters("train"))

    def _host_counters(self, stage: str) -> Dict[str, float]:
        """The collector's step counters as plain host ints (no device sync, unlike get_metrics()): baselines of the time markers."""
        col = self.get_rollout_collector(stage)
        return {"cnt/total_env_steps": col.total_steps, "cnt/total_vec_steps": col.total_vec_steps, "cnt/epoch": self.current_epoch}

    def on_train_epoch_start(self) -> bool:
        """reference agents/base_agent.py:284-328: env-step budget check, then collect this epoch's rollout.  False = stop."""
        cfg, col = self.config, self.get_rollout_collector("train")
        self.timings.start("on_train_epoch_start", values=self._host_counters("train"))
        self._read_hyperparameters_from_run()      # the user may have edited the run's config.json (reference :297-299)
        self._log_hyperparameters()
        if cfg.max_env_steps is not None:
            cur, nxt = col.total_steps * self.world_size, int(cfg.n_envs) * int(cfg.n_steps)
            if cur + nxt > cfg.max_env_steps:
                self.set_early_stop_reason(f"'train/cnt/total_env_steps': {cur} + {nxt} would exceed {int(cfg.max_env_steps)}.")
                return False
        self._trajectories = col.collect()
        return True

    def validation_epoch(self) -> Dict[str, Any]:
        """reference agents/base_agent.py:377-496.  Synchronous: evaluate now, record under "val" for the dispatch callback.
        ``eval_async``: start (or queue) a background evaluation of the current weights and record the most recent finished one,
        which may belong to an earlier epoch (``eval/model_epoch`` says which)."""
        self.timings.start("on_validation_epoch_start", values={"cnt/epoch": self.current_epoch})
        if bool(getattr(self.config, "eval_async", False)):
            self._raise_async_eval_error()
            self._launch_async_eval()
            with self._async_eval_lock:
                ev = dict(self._async_eval_metrics)
        else:
            ev = self.evaluate("val")
        scalars = {}
        for k, v in ev.items():
            try:
                scalars[k] = float(v)              # numpy / torch scalars included; arrays and dicts are not loggable
            except (TypeError, ValueError):
                pass
        if scalars:
            self.metrics_recorder.record("val", scalars)
        mean = ev.get("roll/ep_rew/mean")
        if mean is not None:
            self.best_eval_reward = max(self.best_eval_reward, mean)
        return ev

    # ------------------------------------------------------------------------------------------------ asynchronous evaluation
    def _launch_async_eval(self, eval_epoch: Optional[int] = None) -> None:
        """reference agents/base_agent.py:387-463.  Called on the training thread.  The weights of ``eval_epoch`` are snapshotted with
        one device copy ON THE TRAINING STREAM (ordered after the last optimizer step, no host sync); a daemon thread evaluates
        the snapshot on its own CUDA stream with the val collector / env / network copy.  While an evaluation is running the
        newest request is kept as pending (its snapshot replaces an older pending one) and starts when the running one ends."""
        if self._async_eval_shutdown.is_set():
            return
        if "val" not in self._eval_models:
            raise N.EngineError("eval_async needs the val collector's own network copy: set config.eval_async before constructing the agent")
        eval_epoch = int(self.current_epoch) if eval_epoch is None else int(eval_epoch)
        with self._async_eval_lock:
            # under the lock the eval thread is not reading the snapshot (it copies it out and synchronises under the same lock)
            if self._eval_snapshot is None:
                self._eval_snapshot = torch.empty_like(self.policy_model.flat_params)
                self._eval_snapshot_event = torch.cuda.Event()
                self._eval_stream = torch.cuda.Stream(device=self.device)
            self._eval_snapshot.copy_(self.policy_model.flat_params)
            self._eval_snapshot_event.record()
            if self._async_eval_thread is not None and self._async_eval_thread.is_alive() and self._async_eval_running_epoch is not None:
                self._async_eval_pending_epoch = eval_epoch
                return
            self._async_eval_running_epoch = eval_epoch
            self._async_eval_pending_epoch = None
        self._async_eval_thread = threading.Thread(target=self._run_async_eval, name="gs-async-eval", daemon=True)
        self._async_eval_thread.start()

    def _run_async_eval(self) -> None:
        try:
            with torch.cuda.device(self.device), torch.cuda.stream(self._eval_stream):
                while not self._async_eval_shutdown.is_set():
                    with self._async_eval_lock:
                        epoch = self._async_eval_running_epoch
                        self._eval_stream.wait_event(self._eval_snapshot_event)
                        self._eval_models["val"].flat_params.copy_(self._eval_snapshot)
                        self._eval_stream.synchronize()
                    t0 = time.time()
                    metrics = self.evaluate("val")
                    dt = time.time() - t0
                    with self._async_eval_lock:
                        self._async_eval_metrics = {**metrics, "cnt/epoch": int(epoch), "eval/model_epoch": int(epoch),
                                                    "epoch_fps": metrics.get("cnt/total_vec_steps", 0) / dt if dt > 0 else 0.0}
                        self._async_eval_running_epoch = self._async_eval_pending_epoch
                        self._async_eval_pending_epoch = None
                        if self._async_eval_running_epoch is None:
                            return
        except BaseException as e:                      # fail fast: re-raised on the training thread at the next validation epoch
            with self._async_eval_lock:
                self._async_eval_error = e
                self._async_eval_running_epoch = None
                self._async_eval_pending_epoch = None

    def _raise_async_eval_error(self) -> None:
        with self._async_eval_lock:
            err, self._async_eval_error = self._async_eval_error, None
        if err is not None:
            raise N.EngineError(f"asynchronous evaluation failed: {err!r}") from err

    def wait_async_eval(self, timeout: Optional[float] = None) -> Dict[str, Any]:
        """Block until the running (and any pending) background evaluation has finished; returns its metrics."""
        th = self._async_eval_thread
        if th is not None:
            th.join(timeout)
        self._raise_async_eval_error()
        with self._async_eval_lock:
            return dict(self._async_eval_metrics)

    def get_async_eval_metric(self, metric_key: str) -> Optional[float]:
        """reference agents/base_agent.py:649-652."""
        with self._async_eval_lock:
            return self._async_eval_metrics.get(metric_key)

    def _cleanup_async_eval(self) -> None:
        """reference agents/base_agent.py:498-507 (there the join gives up after 5 s; here a running evaluation is allowed to end, the
        pending one is dropped)."""
        if not bool(getattr(self.config, "eval_async", False)):
            return
        with self._async_eval_lock:
            self._async_eval_pending_epoch = None
        self._async_eval_shutdown.set()
        th = self._async_eval_thread
        if th is not None and th.is_alive():
            th.join()
        self._async_eval_thread = None
        self._raise_async_eval_error()

    def on_fit_end(self) -> None:
        self._cleanup_async_eval()

    def on_exception(self, trainer, pl_module, exception) -> None:
        """reference agents/base_agent.py:509-511: training failed -- stop and join the background evaluation; the training error is the
        one that propagates, an evaluation error on top of it is dropped."""
        try:
            self._cleanup_async_eval()
        except N.EngineError:
            pass

    def log_dict(self, metrics: Dict[str, Any]) -> None:
        if self.trainer is not None:
            self.trainer.log_dict(metrics)
        self.metrics_recorder.update_history(metrics)

    def set_early_stop_reason(self, reason: str) -> None:
        self._early_stop_reason = reason

    def fit(self, *, checkpoint_dir=None, callbacks=None, loggers=(), max_epochs: Optional[int] = None) -> Dict[str, Any]:
        """``learn()`` through the trainer shell (gymnasium_solver_b200/trainer.py): the reference's callback protocol."""
        from ..trainer import Trainer, build_callbacks

        cbs = list(callbacks) if callbacks is not None else build_callbacks(self, checkpoint_dir=checkpoint_dir)
        return Trainer(callbacks=cbs, loggers=loggers, max_epochs=max_epochs).fit(self)

    def evaluate(self, stage: str = "val") -> Dict[str, Any]:
        col = self.get_rollout_collector(stage)
        per_rank = max(1, int(self.config.eval_episodes) // self.world_size)
        return col.evaluate_episodes(n_episodes=per_rank, deterministic=bool(self.config.eval_deterministic))

    # ------------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, checkpoint_dir) -> None:
        """model.pt / optimizer.pt / state.json in the reference's format (agents/base_agent.py:658-732: state-dict, LIST of optimizer
        state-dicts, state.json with epoch / total_env_steps / total_vec_steps / run_id / config / best_train_reward / best_val_reward /
        rng_states), so either side loads the other's checkpoints -- plus what the reference cannot resume (TODO.md:29): an exact
        snapshot of this rank's device envs (physics state, episode accumulators, autoreset flags, reset-stream counters, count-bonus
        tables), the collector's current observations and running statistics.  Resuming from it continues the run bit for bit.
        Rank 0 writes the shared files, every rank its own env shard."""
        import dataclasses
        import random

        d = Path(checkpoint_dir)
        d.mkdir(parents=True, exist_ok=True)
        col = self.get_rollout_collector("train")
        col._resolve_pending_episodes()
        if self.rank == 0:
            finite = lambda x: float(x) if math.isfinite(x) else None
            torch.save(self.policy_model.state_dict(), d / "model.pt")
            torch.save([self.optimizers().state_dict()], d / "optimizer.pt")
            config = dataclasses.asdict(self.config) if dataclasses.is_dataclass(self.config) else dict(vars(self.config))
            config["algo_id"] = self.config.algo_id
            np_state = np.random.get_state()
            # total_env_steps is the GLOBAL count (what the budget, the schedules and the logged metric use); per_rank_* restore this
            # rank layout's collector counters exactly
            state = {"epoch": int(self.current_epoch), "total_env_steps": col.total_steps * self.world_size, "total_vec_steps": col.total_vec_steps,
                     "per_rank_env_steps": col.total_steps,
                     "run_id": getattr(getattr(self, "run", None), "run_id", None), "config": config,
                     "best_train_reward": finite(col._best_episode_reward),
                     "best_val_reward": finite(self.get_rollout_collector("val")._best_episode_reward),
                     "rng_states": {"torch": torch.get_rng_state().tolist(),
                                    "torch_cuda": [t.tolist() for t in torch.cuda.get_rng_state_all()],
                                    "numpy": {"state_type": np_state[0], "state_keys": np_state[1].tolist(), "state_pos": int(np_state[2]),
                                              "state_has_gauss": int(np_state[3]), "state_cached_gaussian": float(np_state[4])},
                                    "random": random.getstate()},
                     # engine additions
                     "best_eval_reward": finite(self.best_eval_reward), "algo_id": self.config.algo_id, "env_id": self.config.env_id,
                     "rng_seed": col.rng_seed, "world_size": self.world_size}
            (d / "state.json").write_text(json.dumps(state, indent=2, default=str))
        env = self.get_env("train")
        torch.save({"snapshot": env.snapshot().cpu(), "obs": None if col.obs is None else col.obs.cpu(),
                    "stats": None if col._stats_dev is None else col._stats_dev.cpu(), "n_envs": env.num_envs,
                    "env_id_offset": self.shard.env_id_offset,
                    # stamp: a shard is only an exact resume point next to the weights of the same save
                    "epoch": int(self.current_epoch), "total_vec_steps": int(col.total_vec_steps), "world_size": self.world_size},
                   d / f"env_state.rank{self.rank}.pt")

    def _load_model_state(self, state_dict, strict: bool) -> None:
        """reference agents/base_agent.py:754-781: strict load, or (transfer learning) only the tensors whose key and shape match."""
        if strict:
            self.policy_model.load_state_dict(state_dict, strict=True)
            return
        own = self.policy_model.state_dict()
        keep = {k: v for k, v in state_dict.items() if k in own and v.shape == own[k].shape}
        res = self.policy_model.load_state_dict(keep, strict=False)
        skipped = sum(1 for k, v in state_dict.items() if k in own and v.shape != own[k].shape)
        if self.rank == 0:
            print(f"Partial weight loading: {len(keep)} params loaded, {skipped} skipped (size mismatch), {len(res.missing_keys)} missing"
                  if keep else "Warning: No compatible weights found for transfer learning")

    def load_checkpoint(self, checkpoint_dir, resume_training: bool = True, strict: bool = True, load_optimizer_only: bool = False) -> None:
        """reference agents/base_agent.py:734-885 (same arguments): weights always; with ``resume_training`` the optimizer state (a list or
        a single state-dict), then -- unless ``load_optimizer_only`` -- host RNG states, step counters and best rewards; and, when the
        checkpoint carries this rank's env shard for the same world size, the exact device-env snapshot."""
        import random

        d = Path(checkpoint_dir)
        if (d / "model.pt").exists():
            self._load_model_state(torch.load(d / "model.pt", map_location=self.device, weights_only=True), strict)
        elif (d / "policy.ckpt").exists():                       # the reference's old single-file format
            ck = torch.load(d / "policy.ckpt", map_location=self.device, weights_only=False)
            self._load_model_state(ck.get("model_state_dict", ck), strict)
        else:
            raise FileNotFoundError(f"Model checkpoint not found at {d / 'model.pt'} or {d / 'policy.ckpt'}")
        if not resume_training or not (d / "state.json").exists():
            return
        state = json.loads((d / "state.json").read_text())
        if (d / "optimizer.pt").exists():
            opt_states = torch.load(d / "optimizer.pt", map_location=self.device, weights_only=False)
            self.optimizers().load_state_dict(opt_states[0] if isinstance(opt_states, (list, tuple)) else opt_states)
        if load_optimizer_only:
            return
        rs = state.get("rng_states")
        if rs:
            torch.set_rng_state(torch.tensor(rs["torch"], dtype=torch.uint8))
            if rs.get("torch_cuda") and len(rs["torch_cuda"]) == torch.cuda.device_count():
                torch.cuda.set_rng_state_all([torch.tensor(t, dtype=torch.uint8) for t in rs["torch_cuda"]])
            ns = rs["numpy"]
            np.random.set_state((ns["state_type"], np.array(ns["state_keys"], dtype=np.uint32), ns["state_pos"], ns["state_has_gauss"],
                                 ns["state_cached_gaussian"]))
            random.setstate((rs["random"][0], tuple(rs["random"][1]), rs["random"][2]))
        col = self.get_rollout_collector("train")
        self.current_epoch = int(state.get("epoch", 0))
        # the collector counts this rank's steps; state.json holds the global count (the reference's field, and what a checkpoint
        # written under another world size must be read as)
        col.total_steps = int(state.get("total_env_steps", state.get("total_timesteps", 0))) // self.world_size
        col.total_vec_steps = int(state.get("total_vec_steps", 0))
        best_train = state.get("best_train_reward", state.get("best_episode_reward"))
        if best_train is not None:
            col._best_episode_reward = float(best_train)
        if state.get("best_val_reward") is not None:
            self.get_rollout_collector("val")._best_episode_reward = float(state["best_val_reward"])
        if state.get("best_eval_reward") is not None:
            self.best_eval_reward = float(state["best_eval_reward"])
        shard = d / f"env_state.rank{self.rank}.pt"
        if shard.exists() and int(state.get("world_size", 1)) == self.world_size:      # same sharding: continue bit for bit
            es = torch.load(shard, map_location="cpu")
            stamped = "epoch" in es
            if stamped and (int(es["epoch"]) != int(state.get("epoch", -1)) or int(es["total_vec_steps"]) != int(state.get("total_vec_steps", -1))):
                # a shard of another save (ranks once disagreed on when to checkpoint): weights and env state would not belong together
                if self.rank == 0:
                    print(f"Ignoring {shard.name}: written at epoch {es['epoch']}, state.json is epoch {state.get('epoch')}")
                es = {"obs": None}
            if es["obs"] is not None:
                col._sync_device_and_prepare_buffers()
                self.get_env("train").restore(es["snapshot"])
                col.obs.copy_(es["obs"].to(self.device))
                if es.get("stats") is not None:
                    col._stats_dev.copy_(es["stats"].to(self.device))
                    col._stats_synced = np.zeros_like(col._stats_synced)
                    col._flush_stats()
