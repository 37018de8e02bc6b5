#!/usr/bin/env python
"""Headline benchmark: PPO env-steps/sec (collect + GAE + update) on CartPole-v1, BASELINE.json config C2
(65,536 GPU-resident envs per GPU, n_steps=128, 64x64 MLP, PPO defaults: 10 passes, 8 minibatches per pass).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one training iteration over one rollout: fused collect (128 vector steps of 65,536 envs), GAE, then
n_epochs x n_minibatches fused forward/loss/backward kernels each followed by the gradient all-reduce (N > 1), global-norm
clip and the torch Adam step.  Weak scaling: every rank owns 65,536 envs.  Rank 0 prints ONE JSON line.

`value`   device-timed, no host round trips inside the timed region (inputs — env state, weights — resident in HBM).
`e2e`     the same iteration through the public API (agent.train_one_rollout + metrics) with, every step, the
          hyper-parameter block copied host->device from pinned memory and the epoch metrics, episode statistics and a
          weight snapshot copied device->host into pinned memory, then a host synchronisation.
`--impl reference`  the reference's CPU path restated by oracle/ (C env step loop + torch CPU policy + numpy GAE + torch
          autograd PPO loss + clip + Adam) on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_PASS = {(64, 64): 27264, (256, 256): 403968}   # 3 x forward flops
GAE_BYTES_PER_ELEM = 18                                           # v 4 + r 4 + done 1 + timeout 1 read, adv 4 + ret 4 written (the all-zero bootstrap array is not read)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_sustained": 1400.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons = index, period, [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} \
            if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
            {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------- b200 arm
OTHER_CONFIGS = {"c3": ("CartPole-v1", "reinforce_b200"), "c4_acrobot": ("Acrobot-v1", "ppo_b200"), "c4_mcar": ("MountainCar-v0", "ppo_b200")}
OTHER_WORKLOADS = {"c3": "CartPole-v1:reinforce, policy_targets=returns + MC baseline, 262,144 envs total, n_steps=128, 64x64 MLP (BASELINE.json configs[2])",
                   "c4_acrobot": "Acrobot-v1:ppo, 1,048,576 envs total sharded over the ranks, n_steps=128, 128x128 MLP (BASELINE.json configs[3])",
                   "c4_mcar": "MountainCar-v0:ppo + StateCountBonus 50x50, 1,048,576 envs total sharded over the ranks, n_steps=128, 256x256 MLP (BASELINE.json configs[3])"}


def build_agent_for_bench(args, rank, world):
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config

    which = getattr(args, "config", "c2")
    if which != "c2":
        # BASELINE.json configs[2] / configs[3] exactly as their *_b200 YAML variants state them; the env count is the TOTAL over the
        # ranks (strong scaling: "1M envs sharded across 1/2/4/8 B200"), so every rank owns n_envs / world envs and batch / world samples
        env_id, variant = OTHER_CONFIGS[which]
        cfg = load_config(env_id, variant)
        cfg.eval_freq_epochs = None
        cfg.max_env_steps = None
        cfg.validate()
        return build_agent(cfg, rank=rank, world_size=world), cfg
    cfg = load_config("CartPole-v1", "ppo_b200")
    cfg.n_envs = args.n_envs * world
    cfg.n_steps = args.n_steps
    cfg.batch_size = args.batch_size * world
    cfg.n_epochs = args.n_epochs
    cfg.model_id = args.model_id
    from gymnasium_solver_b200.utils.model_registry import resolve_model_spec
    cfg._hidden_dims = resolve_model_spec(args.model_id).hidden_dims
    cfg.eval_freq_epochs = None
    cfg.max_env_steps = None
    cfg.track_activations = bool(args.track_activations)
    cfg.validate()
    return build_agent(cfg, rank=rank, world_size=world), cfg


def launches_per_step(cfg, world):
    """Engine kernels launched per training iteration (ours; torch's optimizer / bookkeeping kernels are not counted)."""
    n_mb = (int(cfg.n_envs) * int(cfg.n_steps)) // int(cfg.batch_size) * int(cfg.n_epochs)
    wide = tuple(cfg.hidden_dims) == (256, 256)
    tensor_path = tuple(cfg.hidden_dims) in ((64, 64), (128, 128), (256, 256))
    per_rollout = 1 + 2 + 1 + 2 + (1 if tensor_path else 0)   # collect, obs/reward moments, gae, adv/ret moments, rollout_pack
    # tensor path: gather pass (offsets + batch moments; all of a rollout's up front when sharded), update kernel,
    # gs_update_finish (reduction + NVLink gradient mean + metrics + clip + Adam).  FMA-pipe path: batch moments, update, finish.
    # per minibatch: gather pass + update kernel + step tail (one launch; three -- reduce, receive, apply -- above 8,192 parameters);
    # 256 x 256: + stage_w2_kernel and wgrad_wide_kernel (csrc/update_wide.cu); several ranks: + one moment-exchange kernel per pass
    hd = tuple(cfg.hidden_dims)
    per_mb = 3 + (2 if hd in ((128, 128), (256, 256)) else 0) + (2 if wide else 0)
    n_pass = int(cfg.n_epochs) if world > 1 else 0
    return per_rollout + n_mb * per_mb + n_pass


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dist = torch.distributed
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    agent, cfg = build_agent_for_bench(args, rank, world)
    col = agent.get_rollout_collector("train")
    steps_per_iter_local = agent.local_n_envs * int(cfg.n_steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also first-touch of every buffer) ----------------------------------------------------------------
    for _ in range(args.warmup):
        agent.train_one_rollout()
    agent.pop_epoch_metrics()
    barrier()

    # ---- timed region: device time, no host round trips ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        agent.train_one_rollout()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = steps_per_iter_local * world * args.steps / (ms_total * 1e-3)
    agent.pop_epoch_metrics()
    if os.environ.get("GS_DEV_FINISH_TRACE"):     # development builds (-DGS_FINISH_TRACE): per-phase time of the step tail's last block
        import ctypes
        from gymnasium_solver_b200 import _native as _N
        buf = (ctypes.c_ulonglong * 8)()
        ctypes.CDLL(_N.LIB_PATH).gs_debug_finish_trace(buf)
        n_calls = max(1, int(buf[7]))
        print(f"[finish trace] rank {rank}: calls {n_calls}, ns per call: phaseA-all-blocks {buf[0] / n_calls:.0f}, flags {buf[1] / n_calls:.0f}, "
              f"slot-sum {buf[2] / n_calls:.0f}, metrics+norms+adam {buf[3] / n_calls:.0f}", file=sys.stderr, flush=True)

    # ---- e2e: public API + per-step H2D of the hyper-parameter block and D2H of metrics / episode stats / weights ------
    num = lambda name: float(getattr(cfg, name, 0.0) or 0.0) if not isinstance(getattr(cfg, name, 0.0), dict) else 0.0
    hp_host = torch.tensor([num("policy_lr"), num("clip_range"), num("clip_range_vf"), num("vf_coef"), num("ent_coef"), 0, 0, 0],
                           dtype=torch.float32).pin_memory()
    hp_dev = torch.zeros(8, dtype=torch.float32, device=dev)
    P = agent.policy_model.flat_params.numel()
    weights_host = torch.empty(P, dtype=torch.float32).pin_memory()
    metrics_host = torch.empty(40, dtype=torch.float64).pin_memory()
    h2d_bytes = hp_host.numel() * 4
    d2h_bytes = P * 4 + 40 * 8 + 18 * 8          # weights + metric vector + collector running stats (6x3 doubles)
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        hp_host[5] = float(i)
        hp_dev.copy_(hp_host, non_blocking=True)
        agent.train_one_rollout()
        epoch_metrics = agent.pop_epoch_metrics()                 # D2H: 40 doubles
        roll_metrics = col.get_metrics()                          # D2H: running stats + episode window
        weights_host.copy_(agent.policy_model.flat_params, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = steps_per_iter_local * world * e2e_steps / float(e2e_s.item())

    # ---- roofline of the dominant kernel (the fused update kernel), timed alone on its own stream ----------------------
    # (rank 0 only, with the collectives switched off: the microbench must not enter all-reduces the other ranks never join)
    roof = None
    c2 = getattr(args, "config", "c2") == "c2"
    if rank == 0 and c2:
        agent.world_size = 1
        roof = kernel_rooflines(agent, cfg, dev)
        agent.world_size = world
    cpu_base = None
    if rank == 0 and world == 1 and c2 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(args, threads=1, budget_s=20.0)

    if rank == 0:
        peaks = _peaks()
        line = {
            "metric": "PPO env-steps/sec (collect+GAE+update)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak" if c2 else "strong", "vs_baseline": None, "dtype": "f32 (policy/update), f64 (env physics)", "data": "synthetic",
            "config": {"workload": "CartPole-v1:ppo, 65,536 GPU-resident envs per GPU, n_steps=128, 64x64 MLP (BASELINE.json configs[1])" if c2
                                   else OTHER_WORKLOADS[args.config],
                       "n_envs_per_gpu": agent.local_n_envs, "n_envs_total": int(cfg.n_envs), "n_steps": int(cfg.n_steps), "n_epochs": int(agent.n_epochs),
                       "batch_size_total": int(cfg.batch_size), "minibatches_per_step": (int(cfg.n_envs) * int(cfg.n_steps)) // int(cfg.batch_size) * int(cfg.n_epochs),
                       "model_id": args.model_id if c2 else str(tuple(cfg.hidden_dims)), "parallelism": f"dp{world}: envs sharded; gradient mean per minibatch over NVLink peer memory inside gs_update_finish"
                                      if agent._peer is not None else f"dp{world}: envs sharded" + ("; NCCL grad all-reduce per minibatch" if world > 1 else ""),
                       "grad_allreduce": agent.grad_allreduce_mode,
                       "l2": "rollout working set (>=300 MB per GPU) exceeds the 126 MB L2; no explicit flush",
                       "last_policy_loss": epoch_metrics.get("opt/loss/policy"), "last_ep_rew_mean": roll_metrics.get("roll/ep_rew/mean")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "api": "PPOAgent.train_one_rollout() + pop_epoch_metrics() + RolloutCollector.get_metrics()"},
            "gpu_launches": launches_per_step(cfg, world) * args.steps,
            "roofline": roof["update"] if roof else None,
            "roofline_gae": roof["gae"] if roof else None,
            "roofline_collect": roof["collect"] if roof else None,
            "cpu_baseline": cpu_base,
            "peaks_source": peaks["source"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_rooflines(agent, cfg, dev):
    """Per-kernel achieved rates, each kernel timed alone with CUDA events on the launching stream."""
    from gymnasium_solver_b200 import _native as N

    peaks = _peaks()
    col = agent.get_rollout_collector("train")
    traj = col.collect()
    torch.cuda.synchronize()
    out = {}
    reps = 20

    def timed(fn, reps=reps):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3

    # the update kernel ALONE: every minibatch is prepared first (gather pass: sample offsets + minibatch moments into its own
    # buffers), so a timed call launches only update_f16_kernel (deferred reduction); a different minibatch every call
    agent._pack_rollout(traj)          # as train_on_rollout does: 64-byte sample records for the tensor-core kernel's gather
    batches = [b for _, _, b in agent.minibatches(traj, 12345)][:8]
    hd = tuple(cfg.hidden_dims)
    tensor_path = hd in ((64, 64), (128, 128), (256, 256)) and os.environ.get("GS_UPDATE_IMPL", "tc") != "simt"
    if tensor_path:
        offs = torch.empty(len(batches), agent.local_batch_size, dtype=torch.int32, device=dev)
        mom = torch.zeros(len(batches), 6, dtype=torch.float64, device=dev)
        for k, bb in enumerate(batches):
            bb.struct.offsets = N.ptr(offs[k])
            agent._prepare(bb, mom[k])
    it = [0]

    def one_update():
        bb = batches[it[0] % len(batches)]
        agent._launch_step(bb, defer=True, moments=bb.moments)
        it[0] += 1

    t = timed(one_update)
    env = agent.get_env("train")
    obs_dim, n_act = int(traj.tm["obs"].shape[-1]), int((getattr(env, "single_action_space", None) or env.action_space).n)
    flop_per_sample = 6 * (obs_dim * hd[0] + hd[0] * hd[-1] + (n_act + 1) * hd[-1])      # 3 x forward flops
    flops = FLOP_PER_SAMPLE_PASS.get(hd, flop_per_sample) * agent.local_batch_size if (obs_dim, n_act) == (4, 2) else flop_per_sample * agent.local_batch_size
    fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    tp = os.path.join(ROOT, "profiles", "traffic.json")        # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu captures
    traffic_table = json.load(open(tp)) if os.path.exists(tp) else {}
    full_size = (int(cfg.n_steps), agent.local_n_envs, agent.local_batch_size) == (128, 65536, 1048576)   # the captures' launch sizes
    tr = lambda name: traffic_table.get(name, {}).get("dram_bytes_per_launch") if full_size else None
    traffic = tr(("update_wide_kernel" if hd == (256, 256) else "update_f16_kernel") if tensor_path else "update_kernel")
    if tensor_path:
        # tensor-pipe floor of a 128-sample tile from the measured per-instruction costs (probes/bf16_rate.cu, cycles per tcgen05.mma K=16:
        # M=128 N=16: 40, N=64: 50, N=128: 66; M=64 N=16: 25, N=64: 34, N=80: 42): 64x64: 93 MMAs = 3,464 cycles; 128x128: 141 MMAs
        tiles_per_sm = -(-(agent.local_batch_size // 128) // 148)
        tile_cycles = 3464.0 if hd == (64, 64) else (3 * 66 + 24 * 66 + 16 * 40 + 66 + 16 * 40 + 24 * 66 + 16 * 72 + 8 * 66 + 16 * 40)
        kname = f"update_f16_kernel<{hd[0]}, PPO> (gs_ppo_step: tcgen05 kind::f16, fp16x3 split, TMEM accumulators)"
        if hd == (256, 256):
            # update_wide_kernel: 2 x 48 N=256 MMAs (128 cycles each) + bias / layer 1 / dh2 (4 x 128) + heads and three 16-column weight-gradient
            # groups (4 x 32 x 40); wgrad_wide_kernel: 48 N=256 MMAs per tile
            tile_cycles = 96 * 128 + 4 * 128 + 128 * 40 + 48 * 128
            kname = "update_wide_kernel + wgrad_wide_kernel (gs_ppo_step for 256x256: tcgen05 kind::f16, fp16x3 split, W2 streamed by cp.async.bulk)"
        mma_floor_s = tiles_per_sm * tile_cycles / (peaks["sm_max_mhz"] * 1e6)
        out["update"] = {"kernel": kname, "bound": "tensor",
                         "achieved": flops / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": flops / t / 1e12 / peaks["bf16_tflops"],
                         "traffic": traffic,
                         "note": "achieved = algorithmic fp32 FLOP (fwd + dgrad + wgrad of the MLP, 27,264 per sample for 64x64) / average launch time of "
                                 "update_f16_kernel alone (CUDA events on the launching stream, 20 launches on 8 different prepared minibatches); "
                                 "peak = measured dense bf16 (burst: the kernel is timed alone); traffic = ncu dram bytes of one launch (profiles/traffic.json; "
                                 "algorithmic 71 MB). fp32 parity (1e-4) costs 3 fp16 MMAs per product (x3 the algorithmic flops on the pipe), and the K=16 MMAs of a "
                                 "64-wide network (N <= 80) are bound by a ~25-50 cycle per-instruction floor, not by math: tensor-pipe floor of this launch "
                                 f"{mma_floor_s * 1e3:.3f} ms (tiles per SM x measured cycles per tile); fp32 FMA peak {fp32_peak:.1f} TF/s",
                         "algorithmic_flop_per_launch": flops, "avg_launch_s": t, "tensor_pipe_floor_s": mma_floor_s,
                         "tensor_pipe_floor_frac": mma_floor_s / t, "frac_of_fp32_fma_peak": flops / t / 1e12 / fp32_peak}
    else:
        out["update"] = {"kernel": f"update_kernel<{hd}> (gs_ppo_step, fp32 FMA pipe)", "bound": "tensor", "achieved": flops / t / 1e12,
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": flops / t / 1e12 / peaks["bf16_tflops"], "traffic": traffic,
                         "note": f"fp32 FMA-pipe kernel: fp32 FMA peak {fp32_peak:.1f} TF/s = 148 SM x 128 FMA x 2 x sm_max_mhz, of which "
                                 f"{flops / t / 1e12 / fp32_peak:.3f}",
                         "algorithmic_flop_per_launch": flops, "avg_launch_s": t, "frac_of_fp32_fma_peak": flops / t / 1e12 / fp32_peak}
    # GAE kernel
    b = col._buffer
    T, n = int(cfg.n_steps), agent.local_n_envs
    adv, ret = torch.empty_like(b.rewards_buf), torch.empty_like(b.rewards_buf)

    def one_gae():
        N.check(N.lib().gs_gae_zero_boot(N.ptr(b.values_buf), N.ptr(b.rewards_buf), N.ptr(b.dones_buf), N.ptr(b.timeouts_buf), N.ptr(col._last_values),
                                         T, n, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.stream()))

    t = timed(one_gae)
    gbytes = GAE_BYTES_PER_ELEM * T * n
    out["gae"] = {"kernel": "gae_kernel<zero bootstrap> (the collector's call: gs_gae_zero_boot)", "bound": "hbm", "achieved": gbytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "frac": gbytes / t / 1e9 / peaks["hbm_gbs"], "traffic": tr("gae_kernel"), "algorithmic_bytes_per_launch": gbytes, "avg_launch_s": t}

    # fused collect kernel: 34 B/sample store traffic is the algorithmic HBM figure; it is compute (fp32 MLP + fp64 physics) bound
    def one_collect():
        col.collect()

    t = timed(one_collect, reps=5)
    cbytes = (34 + 12) * T * n
    out["collect"] = {"kernel": "collect_f16_kernel<64, CartPole> (tcgen05 forward) + GAE + stats" if hd == (64, 64) else "collect kernel + GAE + stats", "bound": "hbm", "achieved": cbytes / t / 1e9,
                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": cbytes / t / 1e9 / peaks["hbm_gbs"], "traffic": tr("collect_f16_kernel" if hd in ((64, 64), (128, 128)) else "collect_kernel"),
                      "env_steps_per_s": T * n / t, "avg_call_s": t,
                      "note": "whole RolloutCollector.collect() call (collect kernel + GAE + moments); compute-bound, reported for reference"}
    return out


# ----------------------------------------------------------------------------------------------------------- CPU baseline
def cpu_iteration(env, params, opt, n_envs, T, n_epochs, batch, obs, gen, hp):
    """One reference-shaped training iteration on the CPU: per-step policy_act + env.step, numpy GAE, torch PPO update."""
    from oracle import policy as P
    from oracle import returns as R

    D = obs.shape[1]
    obs_buf = np.zeros((T, n_envs, D), np.float32)
    act_buf = np.zeros((T, n_envs), np.int64)
    logp_buf = np.zeros((T, n_envs), np.float32)
    val_buf = np.zeros((T, n_envs), np.float32)
    rew_buf = np.zeros((T, n_envs), np.float32)
    done_buf = np.zeros((T, n_envs), bool)
    to_buf = np.zeros((T, n_envs), bool)
    with torch.no_grad():
        for t in range(T):
            u = torch.rand(n_envs, generator=gen)
            a, lp, v, _ = P.act(params, torch.from_numpy(obs), uniforms=u)
            obs_buf[t], act_buf[t], logp_buf[t], val_buf[t] = obs, a.numpy(), lp.numpy(), v.numpy()
            obs, r, term, trunc, _ = env.step(a.numpy().astype(np.int32))
            rew_buf[t], done_buf[t], to_buf[t] = r, term | trunc, trunc
        _, last_v = P.forward(params, torch.from_numpy(obs))
    adv, ret = R.gae(val_buf, rew_buf, done_buf, to_buf, last_v.numpy(), np.zeros_like(val_buf), 0.99, 0.95)
    flat = lambda x: torch.from_numpy(np.ascontiguousarray(x.swapaxes(0, 1).reshape(n_envs * T, *x.shape[2:])))
    data = [flat(obs_buf), flat(act_buf), flat(logp_buf), flat(val_buf), flat(adv), flat(ret)]
    total = n_envs * T
    for _ in range(n_epochs):
        order = torch.argsort(torch.rand(total, generator=gen))
        for k in range(total // batch):
            idx = order[k * batch:(k + 1) * batch]
            opt.zero_grad()
            loss, _ = P.ppo_loss(params, *(d[idx] for d in data), **hp)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(params.values()), 0.5)
            opt.step()
    return obs


def cpu_baseline(args, threads: int, budget_s: float, steps: int | None = None, warmup: int = 0):
    """The oracle port of the reference's CPU path on a bounded sample of the workload (same T, n_epochs, model;
    fewer envs; minibatch count per pass preserved)."""
    from oracle import envs as OE
    from oracle import policy as P

    torch.set_num_threads(max(1, threads))
    n_envs = args.cpu_envs
    T, n_epochs = args.n_steps, args.n_epochs
    n_mb = max(1, (args.n_envs * args.n_steps) // args.batch_size)
    batch = n_envs * T // n_mb
    env = OE.OracleVecEnv("CartPole-v1", n_envs, seed=42, threads=threads)
    obs, _ = env.reset()
    hd = {"mlp_64x64": (64, 64), "mlp_medium": (256, 256), "mlp_small": (128, 128), "mlp_tiny": (64,)}[args.model_id]
    params = {k: v.requires_grad_(True) for k, v in P.init_params(4, hd, 2, seed=0).items()}
    opt = torch.optim.Adam(list(params.values()), lr=3e-4)
    gen = torch.Generator().manual_seed(0)
    hp = dict(clip_range=0.2, clip_range_vf=0.2, vf_coef=0.5, ent_coef=0.0, normalize_adv=True)
    for _ in range(warmup):
        obs = cpu_iteration(env, params, opt, n_envs, T, n_epochs, batch, obs, gen, hp)
    t0, done_iters = time.perf_counter(), 0
    while True:
        obs = cpu_iteration(env, params, opt, n_envs, T, n_epochs, batch, obs, gen, hp)
        done_iters += 1
        el = time.perf_counter() - t0
        if (steps is not None and done_iters >= steps) or (steps is None and el >= budget_s):
            break
    return {"value": n_envs * T * done_iters / el, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{done_iters} iteration(s) of {n_envs} envs x {T} steps, {n_epochs} passes x {n_mb} minibatches of {batch}, "
                      f"oracle C env loop + torch CPU policy/loss + numpy GAE, {el:.1f} s",
            "host_cores_available": os.cpu_count(), "torch_threads": torch.get_num_threads(), "seconds": el}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on all host threads; rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    threads = os.cpu_count() or 1
    args.cpu_envs = max(args.cpu_envs, 2048)
    base = cpu_baseline(args, threads=threads, budget_s=0.0, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    per_step_ms = base["seconds"] / max(1, args.steps) * 1e3
    line = {"impl": "reference", "metric": "PPO env-steps/sec (collect+GAE+update)", "value": base["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (policy/update), f64 (env physics)", "data": "synthetic",
            "config": {"workload": "CartPole-v1:ppo, n_steps=128, 64x64 MLP, 10 passes x 8 minibatches (BASELINE.json configs[1]); "
                                   f"bounded sample of {args.cpu_envs} envs per step on the host CPU", "model_id": args.model_id,
                       "n_steps": args.n_steps, "n_epochs": args.n_epochs},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4_acrobot", "c4_mcar"],
                    help="c2 = BASELINE.json configs[1] (the metric's configuration, default); the others are configs[2] / configs[3] at full size, strong-scaled")
    ap.add_argument("--n-envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--batch-size", type=int, default=1048576, help="minibatch per GPU")
    ap.add_argument("--n-epochs", type=int, default=10)
    ap.add_argument("--model-id", default="mlp_64x64")
    ap.add_argument("--track-activations", type=int, default=1)
    ap.add_argument("--cpu-envs", type=int, default=512, help="envs of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args(argv)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()
