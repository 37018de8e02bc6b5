"""Thin test-side helpers that call the C-ABI (through gymnasium_solver_b200._native) on torch CUDA tensors."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from gymnasium_solver_b200 import _native as N

DEV = "cuda"


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)) if not isinstance(x, torch.Tensor) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def sync():
    torch.cuda.synchronize()


# ---- returns ------------------------------------------------------------------------------------------------------
def gae(values, rewards, dones, timeouts, last_values, boot, gamma, lam):
    """boot: None (no bootstrap values), a (T, N) array, or the string "zero" (gs_gae_zero_boot: an all-zero array that is not read)."""
    v, r = cu(values, torch.float32), cu(rewards, torch.float32)
    d, to = cu(np.asarray(dones, dtype=np.uint8)), cu(np.asarray(timeouts, dtype=np.uint8))
    lv = cu(last_values, torch.float32)
    T, Nn = r.shape
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    if isinstance(boot, str):
        assert boot == "zero"
        N.check(N.lib().gs_gae_zero_boot(N.ptr(v), N.ptr(r), N.ptr(d), N.ptr(to), N.ptr(lv), T, Nn, gamma, lam, N.ptr(adv), N.ptr(ret), N.stream()))
    else:
        b = None if boot is None else cu(boot, torch.float32)
        N.check(N.lib().gs_gae(N.ptr(v), N.ptr(r), N.ptr(d), N.ptr(to), N.ptr(lv), N.ptr(b), T, Nn, gamma, lam, N.ptr(adv), N.ptr(ret), N.stream()))
    sync()
    return adv.cpu().numpy(), ret.cpu().numpy()


def mc_returns(rewards, dones, timeouts, gamma, episode_mode=False):
    r = cu(rewards, torch.float32)
    d = cu(np.asarray(dones, dtype=np.uint8))
    to = None if timeouts is None else cu(np.asarray(timeouts, dtype=np.uint8))
    T, Nn = r.shape
    ret = torch.empty_like(r)
    lt = torch.empty(Nn, dtype=torch.int32, device=DEV)
    N.check(N.lib().gs_mc_returns(N.ptr(r), N.ptr(d), N.ptr(to), T, Nn, gamma, int(episode_mode), N.ptr(ret), N.ptr(lt), N.stream()))
    sync()
    return ret.cpu().numpy(), lt.cpu().numpy()


def valid_index_map(last_terminal, T):
    lt = cu(last_terminal, torch.int32)
    Nn = lt.numel()
    mask = torch.empty(Nn * T, dtype=torch.uint8, device=DEV)
    imap = torch.empty(Nn * T, dtype=torch.int64, device=DEV)
    nv = torch.zeros(1, dtype=torch.int64, device=DEV)
    wsb = N.lib().gs_valid_index_map_workspace_bytes(Nn)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    N.check(N.lib().gs_valid_index_map(N.ptr(lt), T, Nn, N.ptr(mask), N.ptr(imap), N.ptr(nv), N.ptr(ws), wsb, N.stream()))
    sync()
    return mask.cpu().numpy().astype(bool), imap.cpu().numpy(), int(nv.item())


def moments(x, last_terminal=None):
    xt = cu(x, torch.float32)
    T, Nn = xt.shape
    lt = None if last_terminal is None else cu(last_terminal, torch.int32)
    out = torch.zeros(3, dtype=torch.float64, device=DEV)
    N.check(N.lib().gs_moments(N.ptr(xt), N.ptr(lt), T, Nn, N.ptr(out), N.stream()))
    sync()
    return out.cpu().numpy()


def moments_valid(x, last_terminal):
    """gs_valid_index_map's n_valid + gs_moments_valid: masked moments, or all elements when the rollout has no valid entry."""
    xt = cu(x, torch.float32)
    T, Nn = xt.shape
    lt = cu(last_terminal, torch.int32)
    mask = torch.empty(Nn * T, dtype=torch.uint8, device=DEV)
    imap = torch.empty(Nn * T, dtype=torch.int64, device=DEV)
    nv = torch.zeros(1, dtype=torch.int64, device=DEV)
    wsb = N.lib().gs_valid_index_map_workspace_bytes(Nn)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    N.check(N.lib().gs_valid_index_map(N.ptr(lt), T, Nn, N.ptr(mask), N.ptr(imap), N.ptr(nv), N.ptr(ws), wsb, N.stream()))
    out = torch.zeros(3, dtype=torch.float64, device=DEV)
    N.check(N.lib().gs_moments_valid(N.ptr(xt), N.ptr(lt), N.ptr(nv), T, Nn, N.ptr(out), N.stream()))
    sync()
    return out.cpu().numpy()


def normalize(x, eps=1e-8, shift_only=False):
    xt = cu(x, torch.float32)
    T, Nn = xt.shape
    mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    N.check(N.lib().gs_moments(N.ptr(xt), None, T, Nn, N.ptr(mom), N.stream()))
    y = torch.empty_like(xt)
    if shift_only:
        N.check(N.lib().gs_shift_by_mean(N.ptr(xt), xt.numel(), N.ptr(mom), N.ptr(y), N.stream()))
    else:
        N.check(N.lib().gs_normalize(N.ptr(xt), xt.numel(), N.ptr(mom), eps, N.ptr(y), N.stream()))
    sync()
    return y.cpu().numpy()


# ---- environments ----------------------------------------------------------------------------------------------------
class DevEnv:
    def __init__(self, env_id, n, seed=0, env_id_offset=0, max_episode_steps=0, wrappers=()):
        self.kind = N.ENV_KINDS[env_id]
        self.n = n
        h = C.c_void_p()
        N.check(N.lib().gs_env_create(self.kind, n, env_id_offset, seed, int(max_episode_steps or 0), 0, C.byref(h)))
        self.h = h
        self.D = N.lib().gs_env_obs_dim(self.kind)
        self.S = N.lib().gs_env_state_dim(self.kind)
        for kind, params in wrappers:
            arr = (C.c_double * len(params))(*params)
            N.check(N.lib().gs_wrapper_attach(self.h, kind, arr, len(params)))

    def close(self):
        if self.h:
            N.lib().gs_env_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        obs = torch.empty(self.n, self.D, dtype=torch.float32, device=DEV)
        N.check(N.lib().gs_env_reset(self.h, N.ptr(obs), N.stream()))
        sync()
        return obs.cpu().numpy()

    def set_state(self, state, elapsed=None):
        s = cu(state, torch.float64)
        e = None if elapsed is None else cu(elapsed, torch.int32)
        N.check(N.lib().gs_env_set_state(self.h, N.ptr(s), N.ptr(e), N.stream()))
        sync()

    def get_state(self):
        s = torch.empty(self.S, self.n, dtype=torch.float64, device=DEV)
        e = torch.empty(self.n, dtype=torch.int32, device=DEV)
        N.check(N.lib().gs_env_get_state(self.h, N.ptr(s), N.ptr(e), N.stream()))
        sync()
        return s.cpu().numpy(), e.cpu().numpy()

    def step(self, actions):
        a = cu(np.asarray(actions, dtype=np.int32))
        obs = torch.empty(self.n, self.D, dtype=torch.float32, device=DEV)
        rew = torch.empty(self.n, dtype=torch.float32, device=DEV)
        term = torch.empty(self.n, dtype=torch.uint8, device=DEV)
        trunc = torch.empty(self.n, dtype=torch.uint8, device=DEV)
        ep_r = torch.empty(self.n, dtype=torch.float64, device=DEV)
        ep_l = torch.empty(self.n, dtype=torch.int32, device=DEV)
        N.check(N.lib().gs_env_step(self.h, N.ptr(a), N.ptr(obs), N.ptr(rew), N.ptr(term), N.ptr(trunc), N.ptr(ep_r), N.ptr(ep_l), N.stream()))
        sync()
        return (obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy().astype(bool), trunc.cpu().numpy().astype(bool),
                ep_r.cpu().numpy(), ep_l.cpu().numpy())


# ---- policy ------------------------------------------------------------------------------------------------------------
def dev_params(p):
    return {k: v.detach().to(DEV).contiguous() for k, v in p.items()}


def policy_act(p_dev, obs, *, deterministic=False, uniforms=None, seed=0, offset=0, row_offset=0, activation="relu"):
    m = N.mlp_struct_from_params(p_dev, activation)
    o = cu(obs, torch.float32)
    n = o.shape[0]
    a = torch.empty(n, dtype=torch.int32, device=DEV)
    lp = torch.empty(n, dtype=torch.float32, device=DEV)
    v = torch.empty(n, dtype=torch.float32, device=DEV)
    lg = torch.empty(n, m.n_actions, dtype=torch.float32, device=DEV)
    u = None if uniforms is None else cu(uniforms, torch.float32)
    N.check(N.lib().gs_policy_act(C.byref(m), N.ptr(o), n, seed, offset, row_offset, int(deterministic), N.ptr(u), N.ptr(a), N.ptr(lp),
                                  N.ptr(v), N.ptr(lg), N.stream()))
    sync()
    return a.cpu().numpy(), lp.cpu().numpy(), v.cpu().numpy(), lg.cpu().numpy()


def policy_values(p_dev, obs, activation="relu"):
    m = N.mlp_struct_from_params(p_dev, activation)
    o = cu(obs, torch.float32)
    v = torch.empty(o.shape[0], dtype=torch.float32, device=DEV)
    N.check(N.lib().gs_policy_values(C.byref(m), N.ptr(o), o.shape[0], N.ptr(v), N.stream()))
    sync()
    return v.cpu().numpy()


# ---- update ------------------------------------------------------------------------------------------------------------
def make_batch(T, Nn, obs_tm, actions_tm, logp_tm, values_tm, adv_tm, ret_tm, *, n=None, idx=None, perm_key=0, perm_offset=0,
               perm_len=0, idx_map=None):
    """gs_batch_t over time-major device arrays; returns (struct, keepalive list)."""
    b = N.GsBatch()
    keep = [obs_tm, actions_tm, logp_tm, values_tm, adv_tm, ret_tm]
    b.T, b.N, b.obs_dim = T, Nn, obs_tm.shape[-1]
    b.obs, b.actions, b.logp_old = N.ptr(obs_tm), N.ptr(actions_tm), N.ptr(logp_tm)
    b.values_old = N.ptr(values_tm)
    b.adv, b.ret = N.ptr(adv_tm), N.ptr(ret_tm)
    if idx is not None:
        it = cu(np.asarray(idx, dtype=np.int64))
        keep.append(it)
        b.idx, b.n = N.ptr(it), it.numel()
    else:
        b.idx, b.n = None, int(n if n is not None else T * Nn)
    b.perm_key, b.perm_offset, b.perm_len = perm_key, perm_offset, perm_len
    if idx_map is not None:
        im = cu(np.asarray(idx_map, dtype=np.int64))
        keep.append(im)
        b.idx_map = N.ptr(im)
    return b, keep


def pack_rollout(batch, keep):
    """gs_rollout_pack over the batch's arrays; sets batch.packed (and keeps the buffer alive)."""
    total = int(batch.T) * int(batch.N)
    buf = torch.full((total, 16), float("nan"), dtype=torch.float32, device=DEV)
    N.check(N.lib().gs_rollout_pack(C.byref(batch), N.ptr(buf), N.stream()))
    sync()
    batch.packed = N.ptr(buf)
    keep.append(buf)
    return buf


def flat_to_time_major(x_flat, T, Nn):
    """(N*T, ...) env-major -> (T, N, ...) time-major numpy."""
    x = np.asarray(x_flat)
    return np.ascontiguousarray(x.reshape(Nn, T, *x.shape[1:]).swapaxes(0, 1))


def update_step(algo, p_dev, batch, hp, *, activation="relu", max_norm=None, internal_moments=False):
    m = N.mlp_struct_from_params(p_dev, activation)
    P = N.lib().gs_mlp_param_count(C.byref(m))
    wsb = N.lib().gs_update_workspace_bytes(C.byref(m), 0, int(batch.n))
    assert P > 0 and wsb > 0, N.lib().gs_last_error()
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    grads = torch.full((P,), float("nan"), dtype=torch.float32, device=DEV)
    metrics = torch.zeros(N.N_METRICS, dtype=torch.float64, device=DEV)
    adv_mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    ret_mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    L = N.lib()
    if internal_moments:                      # NULL moments: the step takes them over the minibatch itself
        if algo == "ppo":
            N.check(L.gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), None, N.ptr(grads), N.ptr(metrics), N.ptr(ws), wsb, N.stream()))
        else:
            N.check(L.gs_reinforce_step(C.byref(m), C.byref(batch), C.byref(hp), None, None, N.ptr(grads), N.ptr(metrics), N.ptr(ws), wsb, N.stream()))
    elif algo == "ppo":
        if hp.normalize_adv:
            N.check(L.gs_batch_moments(C.byref(batch), batch.adv, N.ptr(adv_mom), N.stream()))
        N.check(L.gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(adv_mom), N.ptr(grads), N.ptr(metrics), N.ptr(ws), wsb, N.stream()))
    else:
        if hp.normalize_adv:
            N.check(L.gs_batch_moments(C.byref(batch), batch.adv, N.ptr(adv_mom), N.stream()))
        if hp.normalize_returns:
            N.check(L.gs_batch_moments(C.byref(batch), batch.ret, N.ptr(ret_mom), N.stream()))
        N.check(L.gs_reinforce_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(ret_mom), N.ptr(adv_mom), N.ptr(grads), N.ptr(metrics),
                                    N.ptr(ws), wsb, N.stream()))
    sync()
    g_raw = grads.cpu().numpy().copy()
    if max_norm is not None:
        N.check(L.gs_clip_grad_norm(C.byref(m), N.ptr(grads), max_norm, N.ptr(metrics), N.stream()))
        sync()
    mv = metrics.cpu().numpy()
    return g_raw, grads.cpu().numpy(), {k: mv[i] for i, k in enumerate(N.METRIC_KEYS)}
