"""ctypes binding of include/gs_engine.h — the only door from Python into the CUDA engine.

There is NO CPU or PyTorch fallback: if ``csrc/libgs_engine.so`` is missing or a call fails, an exception is raised.
PyTorch is used for device memory (tensor handles), streams and the optimizer step only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgs_engine.so")

ENV_KINDS = {"CartPole-v1": 0, "Acrobot-v1": 1, "MountainCar-v0": 2}
WRAPPER_KINDS = {"MountainCarV0_StateCountBonus": 1, "CartPoleV1_RewardShaper": 2, "MountainCarV0_RewardShaper": 3, "ScriptedReplay": 4}
ACTIVATIONS = {"relu": 0, "tanh": 1}
N_METRICS = 40

# metric enum of gs_engine.h (index -> reference metric key)
METRIC_KEYS = [
    "opt/loss/total", "opt/loss/policy", "opt/loss/entropy", "opt/policy/entropy", "opt/loss/entropy_scaled", "opt/loss/value",
    "opt/loss/value_scaled", "opt/ppo/clip_fraction", "opt/ppo/clip_fraction_vf", "opt/value/explained_var", "opt/ppo/kl",
    "opt/ppo/approx_kl", "roll/adv/norm/mean", "roll/adv/norm/std", "policy_targets_mean", "policy_targets_std",
    "opt/activations/backbone.0/mean", "opt/activations/backbone.0/std", "opt/activations/backbone.0/dead_pct",
    "opt/activations/backbone.0/dead_max", "opt/activations/backbone.2/mean", "opt/activations/backbone.2/std",
    "opt/activations/backbone.2/dead_pct", "opt/activations/backbone.2/dead_max", "opt/grads/norm/all", "opt/grads/norm/backbone",
    "opt/grads/norm/policy_head", "opt/grads/norm/value_head", "opt/grads/clip_coef", "roll/return/norm/mean", "roll/return/norm/std",
    "opt/batch_count",
]
M = {k: i for i, k in enumerate(METRIC_KEYS)}


class EngineError(RuntimeError):
    pass


class GsMlp(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("hidden1", C.c_int32), ("hidden2", C.c_int32), ("n_actions", C.c_int32),
                ("has_value", C.c_int32), ("activation", C.c_int32),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("wp", C.c_void_p), ("bp", C.c_void_p), ("wv", C.c_void_p), ("bv", C.c_void_p)]


class GsRollout(C.Structure):
    _fields_ = [("T", C.c_int32), ("obs_dim", C.c_int32), ("N", C.c_int64),
                ("obs", C.c_void_p), ("next_obs", C.c_void_p), ("actions", C.c_void_p), ("logprobs", C.c_void_p),
                ("values", C.c_void_p), ("rewards", C.c_void_p), ("dones", C.c_void_p), ("timeouts", C.c_void_p),
                ("last_obs", C.c_void_p), ("last_values", C.c_void_p), ("ep_return", C.c_void_p), ("ep_length", C.c_void_p)]


class GsBatch(C.Structure):
    _fields_ = [("n", C.c_int64), ("idx", C.c_void_p), ("perm_key", C.c_uint64), ("perm_offset", C.c_int64),
                ("perm_len", C.c_int64), ("idx_map", C.c_void_p), ("T", C.c_int32), ("obs_dim", C.c_int32), ("N", C.c_int64),
                ("obs", C.c_void_p), ("actions", C.c_void_p), ("logp_old", C.c_void_p), ("values_old", C.c_void_p),
                ("adv", C.c_void_p), ("ret", C.c_void_p), ("packed", C.c_void_p), ("prepared", C.c_int32), ("defer_reduce", C.c_int32),
                ("offsets", C.c_void_p)]


class GsPpoHparams(C.Structure):
    _fields_ = [("clip_range", C.c_float), ("clip_range_vf", C.c_float), ("vf_coef", C.c_float), ("ent_coef", C.c_float),
                ("normalize_adv", C.c_int32), ("track_activations", C.c_int32)]


class GsReinforceHparams(C.Structure):
    _fields_ = [("ent_coef", C.c_float), ("policy_targets", C.c_int32), ("normalize_returns", C.c_int32),
                ("normalize_adv", C.c_int32), ("track_activations", C.c_int32)]


class GsAdam(C.Structure):
    _fields_ = [("params_flat", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("step_count", C.c_void_p),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


class GsFinish(C.Structure):
    _fields_ = [("algo", C.c_int32), ("track_activations", C.c_int32), ("normalize_adv", C.c_int32), ("normalize_ret", C.c_int32),
                ("vf_coef", C.c_float), ("ent_coef", C.c_float), ("max_grad_norm", C.c_float), ("reserved_", C.c_int32)]


PEER_HANDLE_BYTES = 64
PEER_MAX_WORLD = 8

# every symbol include/gs_engine.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double
SIGNATURES = {
    "gs_version": (_i32, []),
    "gs_last_error": (C.c_char_p, []),
    "gs_device_sm_count": (_i32, [_i32]),
    "gs_env_create": (_i32, [_i32, _i64, _i64, _u64, _i32, _i32, C.POINTER(_vp)]),
    "gs_env_destroy": (_i32, [_vp]),
    "gs_env_set_obs_normalization": (_i32, [_vp, _vp, _vp, _i32]),
    "gs_env_obs_dim": (_i32, [_i32]),
    "gs_env_state_dim": (_i32, [_i32]),
    "gs_env_n_actions": (_i32, [_i32]),
    "gs_env_num_envs": (_i64, [_vp]),
    "gs_env_set_state": (_i32, [_vp, _vp, _vp, _vp]),
    "gs_env_get_state": (_i32, [_vp, _vp, _vp, _vp]),
    "gs_env_snapshot_bytes": (_i64, [_vp]),
    "gs_env_save": (_i32, [_vp, _vp, _vp]),
    "gs_env_load": (_i32, [_vp, _vp, _vp]),
    "gs_env_reset": (_i32, [_vp, _vp, _vp]),
    "gs_env_step": (_i32, [_vp] * 9),
    "gs_wrapper_attach": (_i32, [_vp, _i32, C.POINTER(_f64), _i32]),
    "gs_policy_act": (_i32, [C.POINTER(GsMlp), _vp, _i64, _u64, _u64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gs_policy_values": (_i32, [C.POINTER(GsMlp), _vp, _i64, _vp, _vp]),
    "gs_rollout_collect": (_i32, [_vp, C.POINTER(GsMlp), C.POINTER(GsRollout), _vp, _u64, _u64, _i32, _vp]),
    "gs_gae": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _f64, _f64, _vp, _vp, _vp]),
    "gs_gae_zero_boot": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _f64, _f64, _vp, _vp, _vp]),
    "gs_mc_returns": (_i32, [_vp, _vp, _vp, _i32, _i64, _f64, _i32, _vp, _vp, _vp]),
    "gs_returns_to_full_episode": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp]),
    "gs_valid_index_map": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "gs_valid_index_map_workspace_bytes": (_i64, [_i64]),
    "gs_moments": (_i32, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "gs_moments_valid": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "gs_normalize": (_i32, [_vp, _i64, _vp, _f32, _vp, _vp]),
    "gs_shift_by_mean": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "gs_batch_moments": (_i32, [C.POINTER(GsBatch), _vp, _vp, _vp]),
    "gs_update_workspace_bytes": (_i64, [C.POINTER(GsMlp), _i32, _i64]),
    "gs_rollout_pack": (_i32, [C.POINTER(GsBatch), _vp, _vp]),
    "gs_batch_prepare": (_i32, [C.POINTER(GsMlp), C.POINTER(GsBatch), _i32, _i32, _vp, _vp, _i64, _vp]),
    "gs_set_update_impl": (_i32, [_i32]),
    "gs_mlp_param_count": (_i64, [C.POINTER(GsMlp)]),
    "gs_ppo_step": (_i32, [C.POINTER(GsMlp), C.POINTER(GsBatch), C.POINTER(GsPpoHparams), _vp, _vp, _vp, _vp, _i64, _vp]),
    "gs_reinforce_step": (_i32, [C.POINTER(GsMlp), C.POINTER(GsBatch), C.POINTER(GsReinforceHparams), _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "gs_clip_grad_norm": (_i32, [C.POINTER(GsMlp), _vp, _f32, _vp, _vp]),
    "gs_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _vp]),
    "gs_peer_create": (_i32, [_i32, _i32, _i64, _i32, C.POINTER(_vp), _vp]),
    "gs_peer_connect": (_i32, [_vp, _vp]),
    "gs_peer_destroy": (_i32, [_vp]),
    "gs_peer_allreduce_f64": (_i32, [_vp, _vp, _i64, _vp]),
    "gs_update_finish": (_i32, [C.POINTER(GsMlp), C.POINTER(GsBatch), C.POINTER(GsFinish), _vp, C.POINTER(GsAdam), _vp, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


def lib():
    """Load csrc/libgs_engine.so (raises EngineError if it has not been built — there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} not found: build it with `python -m gymnasium_solver_b200.build` "
                              "(the engine has no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.gs_version() != 100:
            raise EngineError(f"libgs_engine.so version {L.gs_version()} != header version 100")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(lib().gs_last_error().decode() or f"engine call failed ({rc})")


def ptr(t) -> int | None:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EngineError("engine calls take CUDA tensors (the engine has no CPU path)")
    if not t.is_contiguous():
        raise EngineError("engine calls take contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def mlp_struct(model) -> GsMlp:
    """gs_mlp_t over the live nn.Parameter storage of an MLPActorCritic / MLPPolicy (utils/models.py).  Built once per parameter storage:
    the struct only holds device pointers and shapes, and it is asked for twice per minibatch (walking the modules and validating eight
    tensors cost ~30 us of host time per minibatch -- a fifth of a sharded 128 x 128 step on 8 GPUs)."""
    key = tuple(int(t.data_ptr()) for t in model.parameters())
    cached = getattr(model, "_gs_mlp_struct_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    s = _build_mlp_struct(model)
    try:
        object.__setattr__(model, "_gs_mlp_struct_cache", (key, s))
    except Exception:
        pass
    return s


def _build_mlp_struct(model) -> GsMlp:
    lin = [m for m in model.backbone if isinstance(m, torch.nn.Linear)]
    if len(lin) not in (1, 2):
        raise EngineError(f"engine supports 1 or 2 hidden layers, got {len(lin)}")
    acts = [m for m in model.backbone if not isinstance(m, torch.nn.Linear)]
    act_name = type(acts[0]).__name__.lower() if acts else "relu"
    if act_name not in ACTIVATIONS:
        raise EngineError(f"activation {act_name!r} unsupported by the engine (relu, tanh)")
    vh = getattr(model, "value_head", None)
    s = GsMlp()
    s.obs_dim, s.hidden1 = lin[0].in_features, lin[0].out_features
    s.hidden2 = lin[1].out_features if len(lin) == 2 else 0
    s.n_actions, s.has_value, s.activation = model.policy_head.out_features, int(vh is not None), ACTIVATIONS[act_name]
    for t in model.parameters():
        if t.dtype != torch.float32 or not t.is_cuda:
            raise EngineError("engine needs float32 CUDA parameters")
    s.w1, s.b1 = ptr(lin[0].weight), ptr(lin[0].bias)
    if len(lin) == 2:
        s.w2, s.b2 = ptr(lin[1].weight), ptr(lin[1].bias)
    s.wp, s.bp = ptr(model.policy_head.weight), ptr(model.policy_head.bias)
    if vh is not None:
        s.wv, s.bv = ptr(vh.weight), ptr(vh.bias)
    return s


def mlp_struct_from_params(p: dict, activation: str = "relu") -> GsMlp:
    """gs_mlp_t from a dict of CUDA tensors keyed w1,b1,[w2,b2],wp,bp,[wv,bv] (tests)."""
    s = GsMlp()
    s.obs_dim, s.hidden1 = p["w1"].shape[1], p["w1"].shape[0]
    s.hidden2 = p["w2"].shape[0] if "w2" in p else 0
    s.n_actions, s.has_value, s.activation = p["wp"].shape[0], int("wv" in p), ACTIVATIONS[activation]
    for k in ("w1", "b1", "w2", "b2", "wp", "bp", "wv", "bv"):
        if k in p:
            setattr(s, k, ptr(p[k]))
    return s
