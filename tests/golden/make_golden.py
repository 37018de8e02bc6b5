"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures (*.npz) are committed; tests never import /root/reference.

What runs unmodified from /root/reference:
  utils/returns_advantages.py  (GAE, MC returns, episode conversion, valid mask / index map, normalisers)
  utils/rollout_buffer.py      (RolloutBuffer.add + flatten_slice_env_major env-major order)
  utils/models.py + utils/policy_ops.py + utils/distributions.py (MLPActorCritic / MLPPolicy forward, policy_act)
  agents/ppo/ppo_agent.py, agents/reinforce/reinforce_agent.py (losses_for_batch + backward + grad norms)
  gym_wrappers/MountainCarV0/state_count_bonus.py, gym_wrappers/MountainCarV0/reward_shaper.py,
  gym_wrappers/CartPoleV1/reward_shaper.py (the per-env reward wrappers, stepped over a scripted sub-env that replays a
  physics trajectory; see golden_wrappers)
  utils/dataloaders.py (build_index_collate_loader_from_collector over a recording collector) and utils/train_launcher.py
  (_parse_config_overrides / _apply_config_overrides)  -> dataloader.json
  trainer_callbacks/hyperparameter_scheduler.py, utils/schedule_resolver.py, utils/rollout_stats.py, utils/config.py::load_config,
  utils/torch.py helpers, utils/samplers.py  -> host_logic.json;  utils/models.py construction + init  -> model_init.npz
Shims: the reference's own Lightning stub (tests/conftest.py:15-81) and a stub `gymnasium` module (the real
package is absent here); neither touches the arithmetic above.  REINFORCE needs config.normalize_advantages
supplied because the reference reads a field REINFORCEConfig lacks.
"""
from __future__ import annotations

import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_shims():
    sys.path.insert(0, REF)
    # Lightning stub shipped with the reference's tests
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_conftest", os.path.join(REF, "tests", "conftest.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # gymnasium stub: only names touched at import time
    gym = types.ModuleType("gymnasium")

    class _Base:
        def __init__(self, *a, **k):
            pass

    for n in ("Wrapper", "ObservationWrapper", "ActionWrapper", "RewardWrapper", "Env"):
        setattr(gym, n, type(n, (_Base,), {}))

    class _Wrapper:                     # gymnasium.Wrapper as far as the reward wrappers use it: keeps the inner env
        def __init__(self, env):
            self.env = env

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)

    gym.Wrapper = _Wrapper
    spaces = types.ModuleType("gymnasium.spaces")
    for n in ("Box", "Discrete", "MultiBinary", "MultiDiscrete", "Dict", "Tuple", "Space"):
        setattr(spaces, n, type(n, (_Base,), {}))
    vector = types.ModuleType("gymnasium.vector")
    for n in ("VectorWrapper", "VectorEnv", "SyncVectorEnv", "AsyncVectorEnv"):
        setattr(vector, n, type(n, (_Base,), {}))
    gym.spaces, gym.vector = spaces, vector
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.vector": vector})
    os.environ.setdefault("WANDB_MODE", "disabled")


def golden_returns():
    from utils import returns_advantages as ra

    cases = {}
    rng = np.random.default_rng(1234)
    for name, (T, N, pdone) in {"small": (4, 2, 0.3), "ragged": (17, 5, 0.15), "mid": (64, 33, 0.05), "nodone": (9, 3, 0.0),
                                "t1": (1, 7, 0.4)}.items():
        values = rng.standard_normal((T, N)).astype(np.float32)
        rewards = rng.standard_normal((T, N)).astype(np.float32) if name != "mid" else np.ones((T, N), np.float32)
        dones = rng.random((T, N)) < pdone
        timeouts = dones & (rng.random((T, N)) < 0.4)
        last_values = rng.standard_normal(N).astype(np.float32)
        boot = np.where(timeouts, rng.standard_normal((T, N)), 0.0).astype(np.float32)
        gamma, lam = (0.99, 0.95) if name != "ragged" else (0.98, 0.8)
        adv, ret = ra.compute_batched_gae_advantages_and_returns(values, rewards, dones, timeouts, last_values, boot, gamma, lam)
        adv0, ret0 = ra.compute_batched_gae_advantages_and_returns(values, rewards, dones, timeouts, last_values, np.zeros_like(boot), gamma, lam)
        mc_term = ra.compute_batched_mc_returns(rewards, dones, np.zeros_like(timeouts), gamma)
        mc_keep = ra.compute_batched_mc_returns(rewards, dones, timeouts, gamma)
        ep_term = ra.convert_returns_to_full_episode(mc_term.copy(), dones, np.zeros_like(timeouts))
        ep_keep = ra.convert_returns_to_full_episode(mc_keep.copy(), dones, timeouts)
        vm_t, im_t = ra._build_valid_mask_and_index_map(dones, np.zeros_like(timeouts))
        vm_k, im_k = ra._build_valid_mask_and_index_map(dones, timeouts)
        d = dict(values=values, rewards=rewards, dones=dones, timeouts=timeouts, last_values=last_values, boot=boot,
                 gamma=np.float64(gamma), lam=np.float64(lam), adv=adv, ret=ret, adv0=adv0, ret0=ret0,
                 mc_term=mc_term, mc_keep=mc_keep, ep_term=ep_term, ep_keep=ep_keep,
                 norm_adv=ra._normalize_advantages(adv), norm_ret=ra._normalize_returns(ret))
        for k, v in dict(vm_t=vm_t, im_t=im_t, vm_k=vm_k, im_k=im_k).items():
            d[k] = np.zeros(0) if v is None else v
            d[k + "_none"] = np.bool_(v is None)
        cases[name] = d
    # the survey's hand-checked case, re-run through the reference
    values = np.array([[.5, -.2], [.1, .3], [.7, 0], [-.4, .9]], np.float32)
    rewards = np.array([[1, 1], [1, 1], [1, 0], [1, 1]], np.float32)
    dones = np.array([[0, 0], [1, 0], [0, 1], [0, 0]], bool)
    timeouts = np.array([[0, 0], [0, 0], [0, 1], [0, 0]], bool)
    boot = np.where(timeouts, 2.0, 0.0).astype(np.float32)
    last_values = np.array([.25, -.5], np.float32)
    adv, ret = ra.compute_batched_gae_advantages_and_returns(values, rewards, dones, timeouts, last_values, boot, 0.99, 0.95)
    cases["survey"] = dict(values=values, rewards=rewards, dones=dones, timeouts=timeouts, last_values=last_values, boot=boot,
                           gamma=np.float64(0.99), lam=np.float64(0.95), adv=adv, ret=ret)
    for name, d in cases.items():
        np.savez(os.path.join(OUT, f"returns_{name}.npz"), **d)
    print("returns:", list(cases))


def golden_buffer():
    from utils.rollout_buffer import RolloutBuffer

    T, N, D = 5, 3, 4
    rng = np.random.default_rng(7)
    buf = RolloutBuffer(N, (D,), np.float32, torch.device("cpu"), maxsize=T)
    start = buf.begin_rollout(T)
    rec = {k: [] for k in ("obs", "next_obs", "actions", "logps", "values", "rewards", "dones", "timeouts")}
    for t in range(T):
        step = dict(obs=rng.standard_normal((N, D)).astype(np.float32), next_obs=rng.standard_normal((N, D)).astype(np.float32),
                    actions=rng.integers(0, 2, N), logps=rng.standard_normal(N).astype(np.float32),
                    values=rng.standard_normal(N).astype(np.float32), rewards=rng.standard_normal(N).astype(np.float32),
                    dones=rng.random(N) < 0.3, timeouts=rng.random(N) < 0.1)
        buf.add(start + t, step["obs"], step["next_obs"], step["actions"], step["logps"], step["values"], step["rewards"],
                step["dones"], step["timeouts"])
        for k, v in step.items():
            rec[k].append(v)
    adv = rng.standard_normal((T, N)).astype(np.float32)
    ret = rng.standard_normal((T, N)).astype(np.float32)
    traj = buf.flatten_slice_env_major(start, start + T, adv, ret)
    out = {"in_" + k: np.stack(v) for k, v in rec.items()}
    out.update(in_adv=adv, in_ret=ret)
    out.update({"out_" + k: getattr(traj, k).numpy() for k in traj._fields})
    np.savez(os.path.join(OUT, "buffer_env_major.npz"), **out)
    print("buffer: ok")


def _state_to_params(sd):
    m = {"backbone.0.weight": "w1", "backbone.0.bias": "b1", "backbone.2.weight": "w2", "backbone.2.bias": "b2",
         "policy_head.weight": "wp", "policy_head.bias": "bp", "value_head.weight": "wv", "value_head.bias": "bv"}
    return {m[k]: v.detach().numpy().copy() for k, v in sd.items()}


def golden_policy_and_losses():
    from utils.models import MLPActorCritic, MLPPolicy
    from utils.policy_ops import policy_act
    from agents.ppo.ppo_agent import PPOAgent
    from agents.reinforce.reinforce_agent import REINFORCEAgent
    import torch.nn as nn

    torch.manual_seed(0)
    for tag, (D, hidden, A, B) in {"cartpole64": (4, (64, 64), 2, 96), "acrobot128": (6, (128, 128), 3, 64),
                                   "mcar256": (2, (256, 256), 3, 48), "tiny64": (4, (64,), 2, 40)}.items():
        model = MLPActorCritic(input_shape=(D,), hidden_dims=hidden, output_shape=(A,), activation="relu")
        # perturb away from the orthogonal/zero-bias init so every gradient term is exercised
        with torch.no_grad():
            for p_ in model.parameters():
                p_.add_(0.05 * torch.randn_like(p_))
        obs = torch.randn(B, D)
        dist, v = model(obs)
        a_det, lp_det, v_det = policy_act(model, obs, deterministic=True)
        out = dict(obs=obs.numpy(), logits=dist.logits.detach().numpy(), probs=dist.probs.detach().numpy(),
                   value=v.detach().numpy(), act_det=a_det.numpy(), logp_det=lp_det.numpy(), value_det=v_det.numpy(),
                   entropy=dist.entropy().detach().numpy())
        out.update({"p_" + k: val for k, val in _state_to_params(model.state_dict()).items()})

        # PPO loss + backward through the reference agent (constructed like tests/test_ppo.py:73-87)
        actions = torch.randint(0, A, (B,))
        old_logp = dist.log_prob(actions).detach() + 0.15 * torch.randn(B)
        values_old = v.detach() + 0.3 * torch.randn(B)
        adv = torch.randn(B) * 2.0 + 0.5
        ret = values_old + adv
        for norm in ("batch", "off"):
            agent = object.__new__(PPOAgent)
            nn.Module.__init__(agent)
            agent.config = SimpleNamespace(normalize_advantages=norm, target_kl=None)
            agent.clip_range, agent.clip_range_vf, agent.vf_coef, agent.ent_coef = 0.2, 0.2, 0.5, 0.01
            agent.policy_model = model
            rec = {}
            agent.metrics_recorder = SimpleNamespace(record=lambda stage, m, rec=rec: rec.update(m))
            model.zero_grad()
            model._track_activations = True
            batch = SimpleNamespace(observations=obs, actions=actions, logprobs=old_logp, values=values_old, advantages=adv, returns=ret)
            res = agent.losses_for_batch(batch, 0)
            acts = model.compute_activation_stats()
            model._track_activations = False
            res["loss"].backward()
            gn = model.compute_grad_norms()
            grads = {k: val for k, val in _state_to_params({n: p_.grad for n, p_ in model.named_parameters()}).items()}
            out.update({f"ppo_{norm}_loss": res["loss"].detach().numpy()})
            out.update({f"ppo_{norm}_g_{k}": val for k, val in grads.items()})
            out.update({f"ppo_{norm}_m_{k}": np.float64(float(val)) for k, val in rec.items()})
            out.update({f"ppo_{norm}_m_{k}": np.float64(val) for k, val in acts.items()})
            out.update({f"ppo_{norm}_m_{k}": np.float64(val) for k, val in gn.items()})
            # clip_grad_norm_ as Lightning's clip_gradients(..., "norm") does
            total = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
            out[f"ppo_{norm}_clip_total"] = np.float64(float(total))
            out.update({f"ppo_{norm}_gc_{k}": val for k, val in _state_to_params({n: p_.grad for n, p_ in model.named_parameters()}).items()})
        out.update(actions=actions.numpy(), old_logp=old_logp.numpy(), values_old=values_old.numpy(), adv=adv.numpy(), ret=ret.numpy())

        # REINFORCE on a policy-only model
        pol = MLPPolicy(input_shape=(D,), hidden_dims=hidden, output_shape=(A,), activation="relu")
        with torch.no_grad():
            for p_ in pol.parameters():
                p_.add_(0.05 * torch.randn_like(p_))
        dist_p, _ = pol(obs)
        old_logp_p = dist_p.log_prob(actions).detach() + 0.05 * torch.randn(B)
        out.update({"rp_" + k: val for k, val in _state_to_params(pol.state_dict()).items()})
        out["r_old_logp"] = old_logp_p.numpy()
        for targets, nr, na in (("returns", "off", "off"), ("advantages", "off", "batch"), ("returns", "batch", "off")):
            agent = object.__new__(REINFORCEAgent)
            nn.Module.__init__(agent)
            agent.config = SimpleNamespace(normalize_returns=nr, normalize_advantages=na, policy_targets=targets)
            agent.ent_coef = 0.01
            agent.policy_model = pol
            rec = {}
            agent.metrics_recorder = SimpleNamespace(record=lambda stage, m, rec=rec: rec.update(m))
            pol.zero_grad()
            batch = SimpleNamespace(observations=obs, actions=actions, logprobs=old_logp_p, advantages=adv, returns=ret)
            res = agent.losses_for_batch(batch, 0)
            res["loss"].backward()
            key = f"rf_{targets}_{nr}_{na}"
            out[f"{key}_loss"] = res["loss"].detach().numpy()
            out.update({f"{key}_g_{k}": val for k, val in _state_to_params({n: p_.grad for n, p_ in pol.named_parameters()}).items()})
            out.update({f"{key}_m_{k}": np.float64(float(val)) for k, val in rec.items()})
        np.savez(os.path.join(OUT, f"policy_{tag}.npz"), **out)
        print("policy/loss:", tag)


def golden_masked_categorical():
    from utils.distributions import MaskedCategorical

    logits = torch.tensor([[0.3, float("-inf"), -1.2, 0.8], [1.0, 2.0, float("-inf"), float("-inf")]])
    d = MaskedCategorical(logits=logits)
    a = torch.tensor([2, 1])
    np.savez(os.path.join(OUT, "masked_categorical.npz"), logits=logits.numpy(), entropy=d.entropy().numpy(),
             logp=d.log_prob(a).numpy(), actions=a.numpy(), probs=d.probs.numpy())
    print("masked categorical: ok")


def golden_wrappers():
    """The reference's reward wrappers, executed as shipped, over a scripted sub-env that replays one physics trajectory (produced by
    oracle/envs.c WITHOUT a wrapper: the physics is not what is being pinned here) through the NEXT_STEP autoreset protocol of
    SyncVectorEnv: after a done the vector env calls ``reset()`` on the wrapped sub-env instead of ``step()``.

    numpy promotion: the reference pins numpy 1.26.4 (uv.lock), where a float32 scalar combined with a Python float promotes to
    float64; this container has numpy 2.3 (NEP 50 keeps float32).  The wrappers only ever combine observation entries with Python
    scalars, so feeding the float32-ROUNDED observations as float64 arrays reproduces the pinned behaviour exactly."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle.envs import OracleVecEnv
    from gym_wrappers.CartPoleV1.reward_shaper import CartPoleV1_RewardShaper
    from gym_wrappers.MountainCarV0.reward_shaper import MountainCarV0_RewardShaper
    from gym_wrappers.MountainCarV0.state_count_bonus import MountainCarV0_StateCountBonus

    class Scripted:
        def __init__(self, obs0, script):
            self.script, self.i, self.obs0, self.unwrapped = script, 0, obs0, self
        def reset(self, **_):
            if self.i == 0 and self.obs0 is not None:
                o, self.obs0 = self.obs0, None
                return o, {}
            o = self.script[self.i][0]
            self.i += 1
            return o, {}
        def step(self, action):
            o, r, te, tr = self.script[self.i]
            self.i += 1
            return o, r, te, tr, {}

    cases = [("mcar_count", "MountainCar-v0", MountainCarV0_StateCountBonus, dict(position_bins=50, velocity_bins=50, bonus_scale=0.1, bonus_type="count"), 30),
             ("mcar_count_log", "MountainCar-v0", MountainCarV0_StateCountBonus, dict(position_bins=7, velocity_bins=5, bonus_scale=1.0, bonus_type="log", min_count=2), 25),
             ("mcar_count_inverse", "MountainCar-v0", MountainCarV0_StateCountBonus, dict(position_bins=12, velocity_bins=9, bonus_scale=0.5, bonus_type="inverse"), 40),
             ("mcar_shaper", "MountainCar-v0", MountainCarV0_RewardShaper, dict(position_reward_scale=100.0, velocity_reward_scale=10.0, height_reward_scale=50.0), 35),
             ("cartpole_shaper", "CartPole-v1", CartPoleV1_RewardShaper, dict(angle_reward_scale=1.0, position_reward_scale=0.25, clip_potential=True), 500),
             ("cartpole_shaper_noclip", "CartPole-v1", CartPoleV1_RewardShaper, dict(angle_reward_scale=0.7, position_reward_scale=1.5, clip_potential=False), 12)]
    out = {}
    T = 240
    for tag, env_id, cls, kwargs, max_steps in cases:
        seed = 100 + len(out)
        phys = OracleVecEnv(env_id, 1, seed=seed, max_episode_steps=max_steps)
        obs0, _ = phys.reset()
        state0, elapsed0 = phys.get_state()
        rng = np.random.default_rng(seed)
        n_act = phys.single_action_space.n
        actions = rng.integers(0, n_act, size=T).astype(np.int32)
        script, base_r, dones = [], [], []
        for t in range(T):
            o, r, te, tr, _ = phys.step(actions[t:t + 1])
            script.append((o[0].astype(np.float64), float(r[0]), bool(te[0]), bool(tr[0])))   # float32-rounded values as float64
            base_r.append(float(r[0])); dones.append(bool(te[0] or tr[0]))
        wrapped = cls(Scripted(obs0[0].astype(np.float64), script), **kwargs)
        wrapped.reset()
        shaped, prev_done = [], False
        for t in range(T):
            if prev_done:                       # SyncVectorEnv NEXT_STEP autoreset: reset instead of step, reward 0
                wrapped.reset()
                rew, te, tr = 0.0, False, False
            else:
                _, rew, te, tr, _ = wrapped.step(int(actions[t]))
            shaped.append(float(rew))
            prev_done = bool(te or tr)
        assert sum(dones) >= 2, (tag, sum(dones))
        params = {k: (v if not isinstance(v, str) else {"count": 0, "inverse": 1, "log": 2}[v]) for k, v in kwargs.items()}
        out.update({f"{tag}_seed": np.int64(seed), f"{tag}_max_steps": np.int64(max_steps), f"{tag}_actions": actions,
                    f"{tag}_state0": state0, f"{tag}_elapsed0": elapsed0, f"{tag}_base_reward": np.array(base_r),
                    f"{tag}_reward": np.array(shaped), f"{tag}_done": np.array(dones),
                    f"{tag}_kwargs": np.array(json.dumps({"id": cls.__name__, **kwargs}))})
        print("wrapper:", tag, "episodes", sum(dones), "shaping abs mean", float(np.mean(np.abs(np.array(shaped) - np.array(base_r)))))
    np.savez(os.path.join(OUT, "wrappers.npz"), **out)


def golden_model_init():
    """MLPActorCritic / MLPPolicy as the reference constructs and initialises them (utils/models.py:233-346 + utils/torch.py:204-258
    orthogonal init) from torch.manual_seed(seed): every parameter tensor, plus the forward outputs on a fixed observation batch."""
    from utils.models import MLPActorCritic, MLPPolicy

    out = {}
    cases = [("ac_relu_64x64", MLPActorCritic, (4,), (64, 64), (2,), "relu", 0), ("ac_tanh_128x128", MLPActorCritic, (6,), (128, 128), (3,), "tanh", 1),
             ("ac_relu_64", MLPActorCritic, (2,), (64,), (3,), "relu", 2), ("ac_relu_256x256", MLPActorCritic, (4,), (256, 256), (2,), "relu", 3),
             ("pol_relu_64x64", MLPPolicy, (4,), (64, 64), (2,), "relu", 4)]
    for tag, cls, ishape, hidden, oshape, act, seed in cases:
        torch.manual_seed(seed)
        m = cls(input_shape=ishape, hidden_dims=hidden, output_shape=oshape, activation=act)
        for k, v in m.state_dict().items():
            out[f"{tag}/{k}"] = v.numpy().copy()
        x = torch.linspace(-1, 1, 5 * ishape[0]).reshape(5, ishape[0])
        res = m(x)
        dist, value = res if isinstance(res, tuple) else (res, None)
        out[f"{tag}/_logits"] = dist.logits.detach().numpy()
        if value is not None:
            out[f"{tag}/_value"] = value.detach().numpy()
        out[f"{tag}/_x"] = x.numpy()
        out[f"{tag}/_meta"] = np.array(json.dumps(dict(cls=cls.__name__, input_shape=ishape, hidden_dims=hidden, output_shape=oshape, activation=act, seed=seed)))
    np.savez(os.path.join(OUT, "model_init.npz"), **out)
    print("model init:", [c[0] for c in cases])


def golden_host_logic():
    """Host-side helpers of the path, executed from the reference: schedule interpolation + warm-up + position resolution
    (trainer_callbacks/hyperparameter_scheduler.py, utils/schedule_resolver.py), RunningStats / RollingWindow
    (utils/rollout_stats.py) and the resolved hyper-parameters of the three target YAML configs (utils/config.py::load_config)."""
    # trainer_callbacks/__init__.py imports every callback (watchdog, wandb video, ...): load the one module by path instead
    import importlib.util
    pkg = types.ModuleType("trainer_callbacks")
    pkg.__path__ = [os.path.join(REF, "trainer_callbacks")]
    sys.modules["trainer_callbacks"] = pkg
    spec = importlib.util.spec_from_file_location("trainer_callbacks.hyperparameter_scheduler",
                                                  os.path.join(REF, "trainer_callbacks", "hyperparameter_scheduler.py"))
    hs = importlib.util.module_from_spec(spec)
    sys.modules["trainer_callbacks.hyperparameter_scheduler"] = hs
    spec.loader.exec_module(hs)
    HyperparameterSchedulerCallback = hs.HyperparameterSchedulerCallback
    from utils.rollout_stats import RollingWindow, RunningStats
    from utils.schedule_resolver import schedule_pos_to_vec_steps

    out = {"schedules": [], "positions": [], "running_stats": [], "rolling_window": [], "configs": {}}
    n_envs, max_env_steps = 8, 100000
    for kind in ("linear", "cosine", "exponential"):
        for warm in (0.0, 0.25):
            for (v0, v1, p0, p1) in ((3e-4, 0.0, 0.0, 1.0), (0.2, 0.05, 0.1, 0.6), (0.0, 0.01, 20000, 60000)):
                s0 = schedule_pos_to_vec_steps(p0, param="x", default_to_max=False, max_env_steps=max_env_steps, n_envs=n_envs)
                s1 = schedule_pos_to_vec_steps(p1, param="x", default_to_max=True, max_env_steps=max_env_steps, n_envs=n_envs)
                got = []
                cb = HyperparameterSchedulerCallback(schedule=kind, parameter="x", start_value=v0, end_value=v1, start_step=s0, end_step=s1,
                                                     warmup_fraction=warm, set_value_fn=lambda m, v: got.append(v))
                steps = [0, 1000, 9999, 10000, 20000, 25000, 33333, 50000, 60000, 75000, 100000, 120000]
                for env_steps in steps:
                    mod = SimpleNamespace(get_rollout_collector=lambda stage, e=env_steps: SimpleNamespace(total_vec_steps=e / n_envs))
                    cb.on_train_epoch_end(None, mod)
                out["schedules"].append(dict(kind=kind, warmup=warm, start_value=v0, end_value=v1, start=p0, end=p1, env_steps=steps, values=got))
    for raw in (None, 0.0, 0.3, 1.0, 5000.0):
        for dflt in (False, True):
            out["positions"].append(dict(raw=raw, default_to_max=dflt,
                                         vec_steps=schedule_pos_to_vec_steps(raw, param="x", default_to_max=dflt, max_env_steps=max_env_steps, n_envs=n_envs)))
    rng = np.random.default_rng(0)
    rs, seq = RunningStats(), []
    for shape in ((7,), (3, 4), (0,), (128, 2)):
        v = (rng.standard_normal(shape) * 3 + 1).astype(np.float64 if len(shape) == 2 else np.float32)
        rs.update(v)
        seq.append(dict(values=v.ravel().tolist(), shape=list(shape), dtype=str(v.dtype), count=rs.count, mean=rs.mean(), std=rs.std()))
    out["running_stats"] = seq
    rw, means = RollingWindow(5), []
    vals = rng.standard_normal(12).tolist()
    for x in vals:
        rw.append(x)
        means.append(rw.mean())
    out["rolling_window"] = dict(maxlen=5, values=vals, means=means, final_len=len(rw))
    from utils.config import load_config
    keys = ["env_id", "algo_id", "n_envs", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range", "clip_range_vf", "vf_coef",
            "ent_coef", "max_grad_norm", "policy_lr", "max_env_steps", "normalize_advantages", "returns_type", "advantages_type", "hidden_dims",
            "activation", "optimizer", "seed", "eval_episodes", "eval_freq_epochs", "env_wrappers", "target_kl", "model_id", "max_episode_steps"]
    for env_id, variant in (("CartPole-v1", "ppo"), ("Acrobot-v1", "ppo"), ("MountainCar-v0", "ppo")):
        cfg = load_config(env_id, variant)
        row = {}
        for k in keys:
            v = getattr(cfg, k, "<absent>")
            row[k] = list(v) if isinstance(v, tuple) else v
        if env_id == "MountainCar-v0":
            row["n_envs"] = "<cpu_count>"        # n_envs "auto" resolves to os.cpu_count() (machine dependent)
        out["configs"][f"{env_id}:{variant}"] = row
    # utils/torch.py:97-119, 177-190: batch_normalize (unbiased std + 1e-8), KL diagnostics (log-ratio clamped to +-20), group grad norm
    from utils import torch as ref_torch
    tg = torch.Generator().manual_seed(9)
    xs = torch.randn(257, generator=tg) * 3 + 0.5
    old_lp = -torch.rand(64, generator=tg) * 2
    new_lp = old_lp + torch.randn(64, generator=tg) * 0.3
    new_lp[0], new_lp[1] = old_lp[0] + 25.0, old_lp[1] - 30.0           # beyond the clamp
    kl, akl = ref_torch.compute_kl_diagnostics(old_lp, new_lp)
    ps = [torch.nn.Parameter(torch.randn(5, 3, generator=tg)), torch.nn.Parameter(torch.randn(7, generator=tg)), torch.nn.Parameter(torch.zeros(2))]
    ps[0].grad, ps[1].grad = torch.randn(5, 3, generator=tg), torch.randn(7, generator=tg)
    out["torch_utils"] = dict(x=xs.tolist(), x_normalized=ref_torch.batch_normalize(xs).tolist(), old_logp=old_lp.tolist(), new_logp=new_lp.tolist(),
                              kl=float(kl), approx_kl=float(akl), grads=[ps[0].grad.reshape(-1).tolist(), ps[1].grad.tolist()],
                              grad_norm=ref_torch.compute_param_group_grad_norm(ps), grad_norm_none=ref_torch.compute_param_group_grad_norm([ps[2]]))
    # MultiPassRandomSampler (utils/samplers.py:7-37): index stream of 3 passes over 37 items from a seeded generator, then after set_epoch
    from utils.samplers import MultiPassRandomSampler
    torch.manual_seed(1234)
    g = torch.Generator().manual_seed(42)
    smp = MultiPassRandomSampler(37, 3, generator=g)
    first = list(smp)
    smp.set_epoch(5)
    out["sampler"] = dict(data_len=37, num_passes=3, torch_seed=1234, generator_seed=42, first=first, after_set_epoch_5=list(smp), length=len(smp))
    # the `spec:` blocks of the three target YAML files: the reference's own statement of each env's spaces, rewards and returns
    import yaml
    out["env_specs"] = {}
    for env_id in ("CartPole-v1", "Acrobot-v1", "MountainCar-v0"):
        doc = yaml.safe_load(open(os.path.join(REF, "config", "environments", f"{env_id}.yaml")))
        sp = doc["spec"]
        out["env_specs"][env_id] = dict(action_space=sp["action_space"], observation_state=sp["observation_space"]["variants"]["state"],
                                        rewards=sp["rewards"], returns=sp["returns"])
    with open(os.path.join(OUT, "host_logic.json"), "w") as f:
        json.dump(out, f, indent=1, default=lambda o: o if isinstance(o, (int, float, str, type(None))) else str(o))
    print("host logic: schedules", len(out["schedules"]), "configs", list(out["configs"]))


def golden_dataloader():
    """utils/dataloaders.py:20-77 executed from the reference over a recording collector: which index batches reach
    ``collector.slice_trajectories`` (n_epochs permutations from a seeded generator, cut into consecutive minibatches), the loader
    length, and the two ValueErrors (no trajectories; batch size that does not divide the rollout)."""
    from utils.dataloaders import build_index_collate_loader_from_collector

    class Collector:
        def slice_trajectories(self, traj, idxs):
            return [int(i) for i in idxs]

    traj = SimpleNamespace(observations=torch.zeros(48, 4))
    torch.manual_seed(77)
    g = torch.Generator().manual_seed(9)
    loader = build_index_collate_loader_from_collector(collector=Collector(), trajectories=traj, batch_size=16, num_passes=3, generator=g)
    first = [b for b in loader]
    second = [b for b in loader]                       # the generator moves on: the next epoch draws new permutations
    errors = {}
    for tag, kw in (("indivisible", dict(trajectories=traj, batch_size=10)), ("missing", dict(batch_size=16))):
        try:
            build_index_collate_loader_from_collector(collector=Collector(), num_passes=1, **kw)
        except ValueError as e:
            errors[tag] = str(e)
    out = dict(data_len=48, batch_size=16, num_passes=3, torch_seed=77, generator_seed=9, length=len(loader), first=first, second=second,
               errors=errors)
    # `--override KEY=VALUE` handling of the launcher (utils/train_launcher.py:22-53, 81-98): type inference and the two ValueErrors
    from utils.config import load_config
    from utils.train_launcher import _apply_config_overrides, _parse_config_overrides
    items = ["policy_lr=0.001", "n_envs=16", "eval_async=true", "normalize_advantages = rollout", "ent_coef=-0.5", "eval_deterministic=FALSE",
             "n_steps= 64", "gamma=.9"]
    cli = dict(items=items, parsed=_parse_config_overrides(items), errors={})
    cfg = _apply_config_overrides(load_config("CartPole-v1", "ppo"), cli["parsed"])
    cli["applied"] = {k: getattr(cfg, k) for k in cli["parsed"]}
    for tag, fn in (("format", lambda: _parse_config_overrides(["policy_lr"])),
                    ("field", lambda: _apply_config_overrides(load_config("CartPole-v1", "ppo"), {"no_such_field": 1}))):
        try:
            fn()
        except ValueError as e:
            cli["errors"][tag] = str(e)
    out["cli_overrides"] = cli
    with open(os.path.join(OUT, "dataloader.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("dataloader:", len(first), "batches per epoch;", errors)


if __name__ == "__main__":
    _install_shims()
    golden_wrappers()
    golden_model_init()
    golden_host_logic()
    golden_dataloader()
    golden_returns()
    golden_buffer()
    golden_masked_categorical()
    golden_policy_and_losses()
