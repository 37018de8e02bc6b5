#!/usr/bin/env python
"""C5 of BASELINE.json: GAE, MC-returns and update-step microbench over n_envs = 2^10 .. 2^22 (T = 128) on one GPU.

Each kernel is timed alone with CUDA events on the launching stream (3 warm-up calls, then `--reps` calls).  Working sets
below the 126 MB L2 stay L2-resident between calls (stated in the table: those rows measure L2, not HBM).  Prints one
markdown table.  Inputs follow SURVEY 8(d): values ~ N(0,1), rewards = 1, dones ~ Bernoulli(0.02) of which 10 % are timeouts,
bootstrapped = 0; the update step runs on a synthetic rollout of the same shape (64x64 MLP, D=4, A=2, minibatch = min(1M, T*N)).
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gymnasium_solver_b200 import _native as N


def timed(fn, reps):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=22)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hbm = 6557.1
    try:
        hbm = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    T, D, A, H = 128, 4, 2, 64
    L = N.lib()
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)
    p = dict(w1=r(H, D) * 0.5, b1=r(H) * 0.1, w2=r(H, H) * 0.15, b2=r(H) * 0.1, wp=r(A, H) * 0.1, bp=r(A) * 0.1, wv=r(1, H) * 0.1, bv=r(1) * 0.1)
    mlp = N.mlp_struct_from_params(p, "relu")
    P = L.gs_mlp_param_count(C.byref(mlp))
    print(f"| n_envs | elements | GAE us | GAE GB/s (22 B/elem) | of {hbm:.0f} GB/s | MC us | MC GB/s (10 B/elem) | working set | update us | minibatch | update TFLOP/s (27,264 flop/sample) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg
        values, boot = r(T, n), torch.zeros(T, n, device=dev)
        rewards = torch.ones(T, n, device=dev)
        u = torch.rand(T, n, generator=g, device=dev)
        dones = (u < 0.02).to(torch.uint8)
        timeouts = (u < 0.002).to(torch.uint8)
        last_v = r(n)
        adv, ret = torch.empty_like(values), torch.empty_like(values)
        st = N.stream()
        t_gae = timed(lambda: N.check(L.gs_gae(N.ptr(values), N.ptr(rewards), N.ptr(dones), N.ptr(timeouts), N.ptr(last_v), N.ptr(boot), T, n,
                                               0.99, 0.95, N.ptr(adv), N.ptr(ret), st)), args.reps)
        t_mc = timed(lambda: N.check(L.gs_mc_returns(N.ptr(rewards), N.ptr(dones), N.ptr(timeouts), T, n, 0.99, 0, N.ptr(ret), None, st)), args.reps)
        # update step on a rollout of this shape
        obs = r(T, n, D) * 0.5
        actions = torch.randint(0, A, (T, n), generator=g, device=dev, dtype=torch.int32)
        logp = r(T, n) * 0.1 - 0.7
        total = T * n
        B = min(1 << 20, total)
        b = N.GsBatch()
        b.T, b.N, b.obs_dim = T, n, D
        b.obs, b.actions, b.logp_old, b.values_old, b.adv, b.ret = (N.ptr(x) for x in (obs, actions, logp, values, adv, ret))
        b.n, b.idx, b.perm_key, b.perm_offset, b.perm_len, b.idx_map = B, None, 77, 0, total, None
        hp = N.GsPpoHparams()
        hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef, hp.normalize_adv, hp.track_activations = 0.2, 0.2, 0.5, 0.0, 1, 1
        wsb = L.gs_update_workspace_bytes(C.byref(mlp), 0, B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        grads = torch.empty(P, device=dev)
        met = torch.zeros(N.N_METRICS, dtype=torch.float64, device=dev)
        t_up = timed(lambda: N.check(L.gs_ppo_step(C.byref(mlp), C.byref(b), C.byref(hp), None, N.ptr(grads), N.ptr(met), N.ptr(ws), wsb, st)), args.reps)
        ws_mb = 22 * total / 1e6
        print(f"| {n} | {total} | {t_gae * 1e6:.1f} | {22 * total / t_gae / 1e9:.0f} | {22 * total / t_gae / 1e9 / hbm:.2f} | {t_mc * 1e6:.1f} | "
              f"{10 * total / t_mc / 1e9:.0f} | {ws_mb:.0f} MB{' (fits L2)' if ws_mb < 126 else ''} | {t_up * 1e6:.1f} | {B} | {27264 * B / t_up / 1e12:.1f} |")
        del values, boot, rewards, u, dones, timeouts, adv, ret, obs, actions, logp, ws


def env_sweep(reps=20):
    """Unfused `gs_env_step` (state resident in HBM, episode-statistics outputs not requested): algorithmic bytes per env-step
    from SURVEY 8(d); the kernel also carries the RecordEpisodeStatistics accumulators (ep_return fp64, ep_length) and the
    autoreset flag, 26 B per env-step that the algorithmic figure does not count."""
    dev = torch.device("cuda", 0)
    L = N.lib()
    hbm = 6557.1
    print()
    print("| env | n_envs | gs_env_step us | GB/s (algorithmic B/env-step) | of 6557 GB/s |")
    print("|---|---|---|---|---|")
    for env_id, bytes_per_step in (("CartPole-v1", 98), ("MountainCar-v0", 58), ("Acrobot-v1", 106)):
        for n in (1 << 16, 1 << 20, 1 << 22):
            h = C.c_void_p()
            N.check(L.gs_env_create(N.ENV_KINDS[env_id], n, 0, 42, 0, 0, C.byref(h)))
            D = L.gs_env_obs_dim(N.ENV_KINDS[env_id])
            obs = torch.empty(n, D, device=dev)
            rew = torch.empty(n, device=dev)
            term, trunc = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
            acts = torch.randint(0, 2, (n,), device=dev, dtype=torch.int32)
            st = N.stream()
            N.check(L.gs_env_reset(h, N.ptr(obs), st))
            t = timed(lambda: N.check(L.gs_env_step(h, N.ptr(acts), N.ptr(obs), N.ptr(rew), N.ptr(term), N.ptr(trunc), None, None, st)), reps)
            print(f"| {env_id} | {n} | {t * 1e6:.1f} | {bytes_per_step * n / t / 1e9:.0f} ({bytes_per_step}) | {bytes_per_step * n / t / 1e9 / hbm:.2f} |")
            L.gs_env_destroy(h)


if __name__ == "__main__":
    main()
    env_sweep()
