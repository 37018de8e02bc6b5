"""GPU tests of the host-side mirror of the reference surface on top of the engine: RolloutCollector (fused collect,
targets, lazy env-major trajectory, metrics, evaluate_episodes, MC path), PPOAgent / REINFORCEAgent (losses_for_batch on
reference-style batches, full iterations) and learning on the reference's own CartPole-v1:ppo configuration."""
import numpy as np
import pytest
import torch

from oracle import policy as OP
from oracle import returns as OR

pytestmark = pytest.mark.gpu


def _cfg(env="CartPole-v1", variant="ppo", **over):
    from gymnasium_solver_b200.utils.config import load_config

    cfg = load_config(env, variant)
    for k, v in over.items():
        setattr(cfg, k, v)
    if "model_id" in over:
        from gymnasium_solver_b200.utils.model_registry import resolve_model_spec
        cfg._hidden_dims = resolve_model_spec(over["model_id"]).hidden_dims
    cfg.validate()
    return cfg


def _oracle_params(model):
    sd = model.state_dict()
    return {k: sd[n].detach().cpu().clone() for k, n in zip(OP.PARAM_ORDER, sd.keys())}


def test_collector_trajectory_layout_targets_and_metrics():
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.rollout_buffer import RolloutTrajectory

    cfg = _cfg(n_envs=24, n_steps=40, batch_size=960, model_id="mlp_64x64", store_next_obs=True, max_episode_steps=25)
    agent = build_agent(cfg, rank=0, world_size=1)
    col = agent.get_rollout_collector("train")
    traj = col.collect()
    T, n = 40, 24
    assert len(traj) == T * n and len(traj.observations) == T * n
    assert traj._fields == RolloutTrajectory._fields
    assert traj.observations.shape == (T * n, 4) and traj.actions.dtype == torch.int64 and traj.dones.dtype == torch.bool
    tm = {k: v.cpu().numpy() for k, v in traj.tm.items() if isinstance(v, torch.Tensor)}
    # env-major order: sample i = env * T + t  (reference tests/test_rollout_buffer.py:82-126)
    for name, src in (("observations", "obs"), ("rewards", "rewards"), ("logprobs", "logprobs"), ("advantages", "adv"), ("returns", "ret"),
                      ("next_observations", "next_obs")):
        got = getattr(traj, name).cpu().numpy()
        np.testing.assert_array_equal(got, tm[src].swapaxes(0, 1).reshape(T * n, *tm[src].shape[2:]), err_msg=name)
    # next_obs[t] == obs[t+1] along the vector step axis
    np.testing.assert_array_equal(tm["next_obs"][:-1], tm["obs"][1:])
    # targets: GAE with the reference's zero bootstrapped array, bit-exact
    adv, ret = OR.gae(tm["values"], tm["rewards"], tm["dones"], tm["timeouts"], col._last_values.cpu().numpy(), np.zeros_like(tm["values"]),
                      cfg.gamma, cfg.gae_lambda)
    np.testing.assert_array_equal(tm["adv"], adv)
    np.testing.assert_array_equal(tm["ret"], ret)
    # NEXT_STEP autoreset: the step after a done carries reward 0 and no flags
    d = tm["dones"].astype(bool)
    assert d.sum() > 0 and (tm["rewards"][1:][d[:-1]] == 0).all() and not d[1:][d[:-1]].any()
    # policy outputs match the oracle forward of the same weights
    p = _oracle_params(agent.policy_model)
    logits, value = OP.forward(p, torch.from_numpy(tm["obs"].reshape(-1, 4)))
    np.testing.assert_allclose(value.numpy(), tm["values"].reshape(-1), rtol=1e-5, atol=2e-6)
    # metrics: reference key set and values recomputed from the buffers
    m = col.get_metrics()
    for k in ("cnt/total_env_steps", "cnt/total_vec_steps", "cnt/total_episodes", "cnt/total_rollouts", "roll/env_steps", "roll/vec_steps",
              "roll/episodes", "roll/fps", "roll/obs/mean", "roll/obs/std", "roll/reward/mean", "roll/reward/std", "roll/return/mean",
              "roll/return/std", "roll/adv/mean", "roll/adv/std", "roll/actions/mean", "roll/actions/std", "action_dist", "roll/baseline/mean",
              "roll/baseline/std", "roll/ep_rew/mean", "roll/ep_len/mean", "roll/ep_rew/best", "roll/ep_rew/last", "roll/ep_len/last"):
        assert k in m, k
    assert m["cnt/total_env_steps"] == T * n and m["cnt/total_vec_steps"] == T and m["roll/episodes"] == int(d.sum())
    np.testing.assert_allclose(m["roll/obs/mean"], tm["obs"].mean(dtype=np.float64), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(m["roll/reward/mean"], tm["rewards"].mean(dtype=np.float64), rtol=1e-6)
    np.testing.assert_allclose(m["roll/adv/std"], tm["adv"].astype(np.float64).std(), rtol=1e-5)
    np.testing.assert_array_equal(m["action_dist"], np.bincount(tm["actions"].ravel(), minlength=2))
    ep_r = agent.get_rollout_collector("train")._buffer.ep_return_buf.cpu().numpy()[d]  # (step, env) order == row-major
    np.testing.assert_allclose(m["roll/ep_rew/mean"], ep_r[-100:].mean(), rtol=1e-9)
    assert m["roll/ep_rew/best"] == ep_r.max() and m["roll/ep_rew/last"] == ep_r[-1]
    # slice_trajectories == fancy indexing of the env-major tensors (reference rollout_collector.py:657-682)
    idx = [5, 0, 17, 959]
    sl = col.slice_trajectories(traj, idx)
    np.testing.assert_array_equal(sl.observations.cpu().numpy(), traj.observations.cpu().numpy()[idx])
    np.testing.assert_array_equal(sl.advantages.cpu().numpy(), traj.advantages.cpu().numpy()[idx])


def test_collector_mc_path_baseline_index_map_and_episode_mode():
    from gymnasium_solver_b200.agents import build_agent

    for variant_rt in ("mc:rtg", "mc:episode"):
        cfg = _cfg("CartPole-v1", "reinforce", n_envs=16, n_steps=64, batch_size=1024, returns_type=variant_rt, max_episode_steps=30)
        agent = build_agent(cfg, rank=0, world_size=1)
        col = agent.get_rollout_collector("train")
        base = OR.RunningStats()
        for it in range(2):
            traj = col.collect()
            tm = {k: v.cpu().numpy() for k, v in traj.tm.items() if isinstance(v, torch.Tensor)}
            zeros = np.zeros_like(tm["timeouts"])
            exp_ret = OR.mc_returns(tm["rewards"], tm["dones"], zeros, cfg.gamma)
            if variant_rt == "mc:episode":
                exp_ret = OR.to_full_episode(exp_ret, tm["dones"], zeros)
            np.testing.assert_array_equal(tm["ret"], exp_ret)
            vm, im = OR.valid_mask_and_index_map(tm["dones"], zeros)
            np.testing.assert_array_equal(tm["idx_map"], im)
            base.update(exp_ret.swapaxes(0, 1).reshape(-1)[vm])
            np.testing.assert_allclose(tm["adv"], exp_ret - np.float32(base.mean()), rtol=1e-5, atol=1e-5)
        m = col.get_metrics()
        np.testing.assert_allclose(m["roll/baseline/mean"], base.mean(), rtol=1e-5)
        np.testing.assert_allclose(m["roll/baseline/std"], base.std(), rtol=1e-4)
        # remapped slicing keeps batch shapes stable and only returns valid samples
        sl = col.slice_trajectories(traj, list(range(1024)))
        np.testing.assert_array_equal(sl.returns.cpu().numpy(), traj.returns.cpu().numpy()[im])


def test_evaluate_episodes_balanced_quotas():
    from gymnasium_solver_b200.agents import build_agent

    cfg = _cfg(n_envs=8, n_steps=32, batch_size=256, model_id="mlp_64x64", max_episode_steps=20)
    agent = build_agent(cfg, rank=0, world_size=1)
    col = agent.get_rollout_collector("val")
    out = col.evaluate_episodes(n_episodes=21, deterministic=True)
    assert out["cnt/total_episodes"] == 21
    assert 1.0 <= out["roll/ep_len/mean"] <= 20.0 and out["roll/ep_rew/mean"] == out["roll/ep_len/mean"]
    assert "action_dist" not in out and out["cnt/total_env_steps"] % (8 * 32) == 0


@pytest.mark.parametrize("algo", ["ppo", "reinforce"])
def test_losses_for_batch_on_reference_style_batches(golden_dir, algo):
    """The reference's own calling convention: a gathered RolloutTrajectory of flat tensors -> {"loss", "early_stop_epoch"}."""
    import os
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.rollout_buffer import RolloutTrajectory

    d = np.load(os.path.join(golden_dir, "policy_cartpole64.npz"))
    variant = "ppo" if algo == "ppo" else "reinforce"
    cfg = _cfg("CartPole-v1", variant, n_envs=8, n_steps=12, batch_size=96, model_id="mlp_64x64", clip_range=0.2, ent_coef=0.01,
               **({"normalize_advantages": "batch"} if algo == "ppo" else {}))
    agent = build_agent(cfg, rank=0, world_size=1)
    prefix = "p_" if algo == "ppo" else "rp_"
    names = list(agent.policy_model.state_dict().keys())
    sd = {n: torch.from_numpy(d[prefix + k]) for k, n in zip(OP.PARAM_ORDER, names) if prefix + k in d.files}
    if algo == "reinforce":       # the engine's reinforce preset is actor-critic like the reference registry: zero the unused value head
        sd.update({n: torch.zeros_like(agent.policy_model.state_dict()[n]) for n in names if n not in sd})
    agent.policy_model.load_state_dict(sd)
    t = lambda k, dt=torch.float32: torch.from_numpy(d[k]).to(dt)
    batch = RolloutTrajectory(observations=t("obs"), actions=t("actions", torch.int64), rewards=torch.zeros(96), dones=torch.zeros(96, dtype=torch.bool),
                              logprobs=t("old_logp" if algo == "ppo" else "r_old_logp"), values=t("values_old"), advantages=t("adv"),
                              returns=t("ret"), next_observations=t("obs"))
    res = agent.losses_for_batch(batch, 0)
    assert set(res) == {"loss", "early_stop_epoch"} and res["early_stop_epoch"] is False
    key = "ppo_batch" if algo == "ppo" else "rf_returns_off_off"
    np.testing.assert_allclose(res["loss"].item(), float(d[f"{key}_loss"]), rtol=1e-4, atol=1e-6)
    ref = np.concatenate([d[f"{key}_g_{k}"].ravel() for k in OP.PARAM_ORDER if f"{key}_g_{k}" in d.files])
    g = agent.policy_model.flat_grads.cpu().numpy()[: ref.size]
    np.testing.assert_allclose(g, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    # parameters' .grad are views of the flat buffer: torch optimizers see the kernel's gradients
    np.testing.assert_array_equal(agent.policy_model.backbone[0].weight.grad.cpu().numpy().ravel(), g[:256])
    agent._backpropagate_and_step(res["loss"])
    m = agent.pop_epoch_metrics()
    np.testing.assert_allclose(m["opt/loss/total"], float(d[f"{key}_loss"]), rtol=1e-4, atol=1e-6)
    if algo == "ppo":
        np.testing.assert_allclose(m["opt/grads/norm/all"], float(d["ppo_batch_m_opt/grads/norm/all"]), rtol=1e-4)


def test_target_kl_early_stop_flag():
    from gymnasium_solver_b200.agents import build_agent

    cfg = _cfg(n_envs=8, n_steps=32, batch_size=256, model_id="mlp_64x64", target_kl=1e-9, policy_lr=0.01)
    agent = build_agent(cfg, rank=0, world_size=1)
    agent.train_one_rollout()   # after the first Adam step approx_kl > 1e-9 -> the remaining passes are skipped
    assert agent._early_stop_epoch is True
    # reference agents/ppo/ppo_agent.py:140: every evaluated minibatch records the flag; the first (ratio == 1) does not trigger, the second does
    assert agent.pop_epoch_metrics()["opt/ppo/kl_stop_triggered"] == 0.5
    free = build_agent(_cfg(n_envs=8, n_steps=32, batch_size=256, model_id="mlp_64x64"), rank=0, world_size=1)
    free.train_one_rollout()
    assert free._early_stop_epoch is False and free.pop_epoch_metrics()["opt/ppo/kl_stop_triggered"] == 0.0


def test_cartpole_ppo_learns_on_the_reference_configuration():
    """C1: CartPole-v1:ppo exactly as shipped (N=8, T=32, 256x256 MLP, 20 passes, 1e5 env steps).  The reference README
    states this configuration solves CartPole (eval mean >= 475) inside the budget."""
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.random import set_random_seed

    # The run stops as soon as the deterministic evaluation reaches the environment's reward threshold (475), which happens after
    # 25-50k env steps; the 100-episode TRAINING mean lags behind it (190-290 at that point).  The 256x256 update kernel accumulates
    # with fp32 atomics, so trajectories are not bit-reproducible and, as with any PPO run, an occasional seed does not solve the
    # task inside 1e5 steps (measured: ~1 run in 4 ends at eval 250-450): up to three seeds, one must meet the solve criterion.
    tried = []
    for seed in (42, 43, 44):
        cfg = _cfg("CartPole-v1", "ppo", seed=seed, seed_train=seed, seed_val=1000 + seed)
        set_random_seed(cfg.seed)
        agent = build_agent(cfg, rank=0, world_size=1)
        out = agent.learn()
        hist = out["history"]
        assert out["total_env_steps"] <= 1e5
        assert all(np.isfinite(r["train/opt/loss/total"]) for r in hist)
        train_curve = [r["train/roll/ep_rew/mean"] for r in hist if "train/roll/ep_rew/mean" in r]
        tried.append((seed, out["best_eval_reward"], max(train_curve)))
        if out["best_eval_reward"] >= 475.0 and max(train_curve) >= 100.0:
            break
    else:
        raise AssertionError(f"CartPole-v1:ppo did not reach eval >= 475 within 1e5 steps for any seed: {tried}")


def test_cartpole_reinforce_learns_on_the_reference_configuration():
    """CartPole-v1:reinforce exactly as shipped (8 envs x 512 steps, one 4,096-sample batch, 64-unit MLP, MC reward-to-go returns as policy
    targets, 2e5 env steps).  profiles/r1f_learning_curves.md: engine and CPU port follow the same curve over 10 seeds each and every run
    passes a 100-episode training mean of 195 after ~100 k steps; the bar here is lower (150) and up to two seeds may be tried."""
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.random import set_random_seed

    tried = []
    for seed in (42, 43):
        cfg = _cfg("CartPole-v1", "reinforce", seed=seed, seed_train=seed, seed_val=1000 + seed, eval_freq_epochs=None,
                   early_stop_on_eval_threshold=False, early_stop_on_train_threshold=False)
        set_random_seed(cfg.seed)
        out = build_agent(cfg, rank=0, world_size=1).learn()
        assert out["total_env_steps"] == 48 * 4096 and all(np.isfinite(r["train/opt/loss/total"]) for r in out["history"])
        curve = [r["train/roll/ep_rew/mean"] for r in out["history"] if "train/roll/ep_rew/mean" in r]
        tried.append((seed, max(curve)))
        if max(curve) >= 150.0 and curve[2] < 60.0:          # it starts from a random policy (~22) and learns
            return
    raise AssertionError(f"CartPole-v1:reinforce did not pass a training mean of 150 within 2e5 steps: {tried}")


def test_checkpoint_roundtrip(tmp_path):
    from gymnasium_solver_b200.agents import build_agent

    cfg = _cfg(n_envs=8, n_steps=32, batch_size=256, model_id="mlp_64x64")
    a = build_agent(cfg, rank=0, world_size=1)
    a.train_one_rollout()
    a.save_checkpoint(tmp_path / "ck")
    b = build_agent(cfg, rank=0, world_size=1)
    b.load_checkpoint(tmp_path / "ck")
    assert torch.equal(a.policy_model.flat_params, b.policy_model.flat_params)
    assert b.get_rollout_collector("train").total_steps == 256
    assert set(torch.load(tmp_path / "ck" / "model.pt").keys()) == {"backbone.0.weight", "backbone.0.bias", "backbone.2.weight", "backbone.2.bias",
                                                                   "policy_head.weight", "policy_head.bias", "value_head.weight", "value_head.bias"}
    # the reference's on-disk format (agents/base_agent.py:658-732): optimizer.pt is a LIST of state-dicts, state.json carries its keys
    import json
    import random

    opt_states = torch.load(tmp_path / "ck" / "optimizer.pt", weights_only=False)
    assert isinstance(opt_states, list) and len(opt_states) == 1 and set(opt_states[0]) == {"state", "param_groups"}
    state = json.loads((tmp_path / "ck" / "state.json").read_text())
    for k in ("epoch", "total_env_steps", "total_vec_steps", "run_id", "config", "best_train_reward", "best_val_reward", "rng_states"):
        assert k in state, k
    assert state["config"]["algo_id"] == "ppo" and state["config"]["n_envs"] == 8 and state["total_vec_steps"] == 32
    assert set(state["rng_states"]) == {"torch", "torch_cuda", "numpy", "random"}
    # host RNG streams continue from the checkpoint after a load
    torch.manual_seed(123); np.random.seed(123); random.seed(123)
    b.load_checkpoint(tmp_path / "ck")
    draws = (torch.rand(3).tolist(), np.random.rand(3).tolist(), random.random())
    b.load_checkpoint(tmp_path / "ck")
    assert draws == (torch.rand(3).tolist(), np.random.rand(3).tolist(), random.random())

    # a checkpoint WRITTEN THE REFERENCE'S WAY (plain torch state-dicts, its state.json keys only) loads into the engine
    ref = tmp_path / "ref_ck"
    ref.mkdir()
    sd = {k: torch.randn_like(v).cpu() for k, v in a.policy_model.state_dict().items()}
    torch.save(sd, ref / "model.pt")
    params = [torch.nn.Parameter(v.clone()) for v in sd.values()]
    topt = torch.optim.Adam(params, lr=7e-4)
    for p_ in params:
        p_.grad = torch.randn_like(p_)
    topt.step()
    torch.save([topt.state_dict()], ref / "optimizer.pt")
    (ref / "state.json").write_text(json.dumps({"epoch": 7, "total_env_steps": 1792, "total_vec_steps": 224, "run_id": "abc123", "best_train_reward": 88.0,
                                                "best_val_reward": 99.0}))
    c = build_agent(cfg, rank=0, world_size=1)
    c.load_checkpoint(ref)
    for k, v in c.policy_model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    got = c.optimizers().state_dict()
    assert got["param_groups"][0]["lr"] == 7e-4
    for i in range(len(params)):
        assert torch.equal(got["state"][i]["exp_avg"].cpu().reshape(-1), topt.state_dict()["state"][i]["exp_avg"].reshape(-1)), i
    col = c.get_rollout_collector("train")
    assert (c.current_epoch, col.total_steps, col.total_vec_steps, col._best_episode_reward) == (7, 1792, 224, 88.0)
    assert c.get_rollout_collector("val")._best_episode_reward == 99.0
    c.train_one_rollout()                                    # and training goes on from it
    # load_optimizer_only: optimizer warm start, fresh counters; strict=False: only tensors whose shape matches; no model file: FileNotFoundError
    e = build_agent(cfg, rank=0, world_size=1)
    e.load_checkpoint(ref, load_optimizer_only=True)
    assert e.current_epoch == 0 and e.get_rollout_collector("train").total_steps == 0 and e.optimizers().state_dict()["param_groups"][0]["lr"] == 7e-4
    odd = dict(sd)
    odd["policy_head.weight"] = torch.zeros(5, 64)
    torch.save(odd, ref / "model.pt")
    with pytest.raises(RuntimeError):
        e.load_checkpoint(ref, resume_training=False)
    before = e.policy_model.policy_head.weight.clone()
    e.load_checkpoint(ref, resume_training=False, strict=False)
    assert torch.equal(e.policy_model.policy_head.weight, before) and torch.equal(e.policy_model.value_head.weight.cpu(), sd["value_head.weight"])
    with pytest.raises(FileNotFoundError):
        e.load_checkpoint(tmp_path / "nothing_here")


@pytest.mark.parametrize("algo,model_id", [("ppo", "mlp_64x64"), ("ppo", "mlp_small"), ("reinforce", "mlp_64x64")])
def test_fused_step_tail_matches_the_generic_path(algo, model_id):
    """gs_update_finish (ordered reduction + metrics + clip + Adam in one launch) against the unfused sequence
    gs_*_step -> gs_clip_grad_norm -> torch.optim.Adam.step() on the same rollout, same minibatches."""
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.optimizer_factory import EngineAdam

    def make(fused):
        over = dict(n_envs=64, n_steps=32, model_id=model_id, fused_update=fused, max_grad_norm=0.05)   # small max norm: the clip is active
        over.update(dict(batch_size=512, n_epochs=3) if algo == "ppo" else dict(batch_size=2048))
        cfg = _cfg("CartPole-v1", "ppo" if algo == "ppo" else "reinforce", **over)
        agent = build_agent(cfg, rank=0, world_size=1)
        if not fused:        # reference optimizer: torch.optim.Adam on the per-tensor parameter views
            agent._optimizer = torch.optim.Adam(agent.policy_model.parameters(), lr=cfg.policy_lr)
        return agent

    a, b = make(True), make(False)
    assert isinstance(a.optimizers(), EngineAdam) and not isinstance(b.optimizers(), EngineAdam)
    np.testing.assert_array_equal(a.policy_model.flat_params.cpu().numpy(), b.policy_model.flat_params.cpu().numpy())
    for _ in range(2):
        ta, tb = a.train_one_rollout(), b.train_one_rollout()
        np.testing.assert_array_equal(ta.tm["obs"].cpu().numpy(), tb.tm["obs"].cpu().numpy())   # same weights -> same rollout
        wa, wb = a.policy_model.flat_params.cpu().numpy(), b.policy_model.flat_params.cpu().numpy()
        np.testing.assert_allclose(wa, wb, rtol=2e-5, atol=2e-6)
        ma, mb = a.pop_epoch_metrics(), b.pop_epoch_metrics()
        for k in ("opt/loss/total", "opt/grads/norm/all", "opt/grads/norm/backbone", "opt/grads/clip_coef", "opt/policy/entropy",
                  "opt/activations/backbone.0/mean"):
            np.testing.assert_allclose(ma[k], mb[k], rtol=1e-4, atol=1e-7, err_msg=k)
        assert ma["opt/grads/clip_coef"] < 1.0
    # optimizer state: torch.optim.Adam's layout, same moments
    sa, sb = a.optimizers().state_dict(), b.optimizers().state_dict()
    assert set(sa["state"].keys()) == set(sb["state"].keys())
    for i in sa["state"]:
        assert float(sa["state"][i]["step"]) == float(sb["state"][i]["step"])
        np.testing.assert_allclose(sa["state"][i]["exp_avg"].cpu().numpy(), sb["state"][i]["exp_avg"].cpu().numpy(), rtol=1e-4, atol=1e-9)
    # the engine optimizer loads a torch.optim.Adam checkpoint and keeps stepping
    a.optimizers().load_state_dict(sb)
    assert int(a.optimizers().step_dev.item()) == int(float(sb["state"][0]["step"]))
    a.train_one_rollout()
    assert np.isfinite(a.policy_model.flat_params.cpu().numpy()).all()


def test_fused_step_is_skipped_for_overridden_losses():
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.agents.ppo.ppo_agent import PPOAgent

    calls = []

    class MyAgent(PPOAgent):
        def losses_for_batch(self, batch, batch_idx):
            calls.append(batch_idx)
            return super().losses_for_batch(batch, batch_idx)

    cfg = _cfg(n_envs=8, n_steps=32, batch_size=128, n_epochs=2, model_id="mlp_64x64")
    agent = MyAgent(cfg, rank=0, world_size=1)
    agent.train_one_rollout()
    assert len(calls) == 4


def test_two_gpu_peer_gradient_exchange():
    """NVLink peer exchange of gs_update_finish vs the NCCL path on 2 ranks (skipped on a single-GPU box)."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", os.path.join(root, "tests", "dist_peer_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_CHECK_OK" in r.stdout


@pytest.mark.parametrize("env,variant,over", [
    ("CartPole-v1", "ppo", dict(n_envs=64, n_steps=32, batch_size=512, n_epochs=2, model_id="mlp_64x64", max_episode_steps=40)),
    ("MountainCar-v0", "ppo", dict(n_envs=32, n_steps=48, batch_size=512, n_epochs=2, model_id="mlp_64x64", max_episode_steps=30)),
    ("CartPole-v1", "reinforce", dict(n_envs=32, n_steps=64, batch_size=2048, model_id="mlp_64x64", max_episode_steps=25)),
])
def test_checkpoint_resume_continues_bit_for_bit(tmp_path, env, variant, over):
    """weights + Adam state + device env snapshot (physics state, episode accumulators, autoreset flags, reset-stream
    counters, count-bonus tables) + collector state -> a resumed run is bit-identical to the uninterrupted one."""
    from gymnasium_solver_b200.agents import build_agent

    cfg = _cfg(env, variant, **over)
    a = build_agent(cfg, rank=0, world_size=1)
    for _ in range(3):
        a.train_one_rollout()
        a.current_epoch += 1
    a.save_checkpoint(tmp_path / "ck")
    for _ in range(2):
        ta = a.train_one_rollout()
        a.current_epoch += 1
    b = build_agent(cfg, rank=0, world_size=1)
    b.load_checkpoint(tmp_path / "ck")
    assert b.current_epoch == 3
    for _ in range(2):
        tb = b.train_one_rollout()
        b.current_epoch += 1
    for k in ("obs", "actions", "rewards", "dones", "adv", "ret"):
        assert torch.equal(ta.tm[k], tb.tm[k]), k
    assert torch.equal(a.policy_model.flat_params, b.policy_model.flat_params)
    assert torch.equal(a.optimizers().exp_avg_sq, b.optimizers().exp_avg_sq)
    sa, sb = a.get_env("train").snapshot(), b.get_env("train").snapshot()
    assert torch.equal(sa, sb)
    assert (tmp_path / "ck" / "env_state.rank0.pt").exists()
    # a snapshot only loads into an env of the same kind / size / wrapper
    other = build_agent(_cfg(env, variant, **{**over, "n_envs": over["n_envs"] * 2}), rank=0, world_size=1)
    with pytest.raises(ValueError, match="does not fit"):
        other.get_env("train").restore(sa)


def test_fit_through_the_trainer_shell_equals_learn(tmp_path):
    """the reference's callback protocol (gymnasium_solver_b200/trainer.py: DispatchMetrics -> scheduler -> EarlyStopping
    -> ModelCheckpoint in callback_builder.py's order) drives the same run as the built-in learn() loop: same epochs, bit-identical
    weights; metrics arrive under the reference's keys, checkpoints under epoch=NN with last / best links."""
    import os

    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.trainer import CsvMetricsLogger

    def make():
        cfg = _cfg(n_envs=32, n_steps=32, batch_size=256, n_epochs=3, model_id="mlp_64x64", max_env_steps=32 * 32 * 9, eval_freq_epochs=3,
                   eval_episodes=16, eval_warmup_epochs=0, early_stop_on_eval_threshold=False, policy_lr=1e-3, policy_lr_schedule="cosine",
                   policy_lr_schedule_start_value=1e-3, policy_lr_schedule_end_value=1e-4, policy_lr_schedule_start=0.0, policy_lr_schedule_end=1.0)
        return build_agent(cfg, rank=0, world_size=1)

    a, b = make(), make()
    assert a.config.policy_lr_schedule == "cosine"
    out_a = a.learn()
    csv_path = tmp_path / "metrics.csv"
    out_b = b.fit(checkpoint_dir=tmp_path / "ck", loggers=[CsvMetricsLogger(csv_path)])
    assert out_a["epochs"] == out_b["epochs"] == 9 and out_a["total_env_steps"] == out_b["total_env_steps"] == 32 * 32 * 9
    assert "would exceed" in out_b["stop_reason"]
    assert torch.equal(a.policy_model.flat_params, b.policy_model.flat_params)
    assert a.optimizers().param_groups[0]["lr"] == b.optimizers().param_groups[0]["lr"] < 1e-3
    lm = b.trainer.logged_metrics
    for k in ("train/opt/loss/total", "train/opt/ppo/approx_kl", "train/roll/ep_rew/mean", "train/cnt/total_env_steps", "train/sys/timing/fps",
              "train/progress", "train/cnt/epoch", "val/roll/ep_rew/mean", "val/cnt/epoch"):
        assert k in lm, k
    assert lm["train/progress"] == 1.0 and lm["train/cnt/total_env_steps"] == 32 * 32 * 9
    # three validation epochs (epochs 2, 5, 8): the first always saves and is best; later ones only when better
    ck = tmp_path / "ck"
    assert (ck / "epoch=02" / "model.pt").exists() and (ck / "epoch=02" / "env_state.rank0.pt").exists()
    assert os.readlink(ck / "last").startswith("epoch=") and os.readlink(ck / "best").startswith("epoch=")
    header = csv_path.read_text().splitlines()[0].split(",")
    assert "train/opt/loss/total" in header and "val/roll/ep_rew/mean" in header


def test_async_evaluation_beside_training_matches_synchronous_evaluation(tmp_path):
    """SURVEY 8(f) n2 / reference agents/base_agent.py:204-213, 387-463: with ``eval_async`` the val collector owns a copy of the network
    and a background thread evaluates a snapshot of the weights on its own CUDA stream while the training stream goes on.  Training is
    unaffected bit for bit; the first background evaluation equals the synchronous evaluation of the same epoch's weights (same val env
    seed, deterministic actions); requests made while one is running collapse into one pending evaluation of the newest weights."""
    from gymnasium_solver_b200.agents import build_agent

    def make(**over):
        cfg = _cfg(n_envs=32, n_steps=32, batch_size=256, n_epochs=3, model_id="mlp_64x64", eval_episodes=64, eval_warmup_epochs=0,
                   eval_deterministic=True, early_stop_on_eval_threshold=False, early_stop_on_train_threshold=False, **over)
        return build_agent(cfg, rank=0, world_size=1)

    sync = make(eval_freq_epochs=6, eval_async=False)
    out_s = sync.learn(max_epochs=6)
    ev_s = {k: v for k, v in out_s["history"][-1].items() if k.startswith("val/")}
    assert ev_s["val/cnt/total_episodes"] == 64

    # (1) one request at the same epoch: same evaluation, same training
    a = make(eval_freq_epochs=6, eval_async=True)
    assert a.get_rollout_collector("val").policy_model is not a.policy_model
    assert a.get_rollout_collector("test").policy_model is a.policy_model
    a.learn(max_epochs=6)                      # on_fit_end joins the background evaluation
    assert a._async_eval_thread is None
    assert torch.equal(a.policy_model.flat_params, sync.policy_model.flat_params)
    ev_a = a.wait_async_eval()
    assert ev_a["eval/model_epoch"] == 5 and ev_a["cnt/epoch"] == 5 and ev_a["epoch_fps"] > 0
    for k in ("roll/ep_rew/mean", "roll/ep_len/mean", "cnt/total_episodes", "cnt/total_env_steps"):
        assert ev_a[k] == ev_s[f"val/{k}"], k
    assert a.get_async_eval_metric("roll/ep_rew/mean") == ev_a["roll/ep_rew/mean"]
    assert torch.equal(a._eval_models["val"].flat_params, a.policy_model.flat_params)      # the snapshot of the last epoch

    # (2) a request every epoch, through the trainer shell: training still bit-identical, evaluations of increasing epochs arrive
    b = make(eval_freq_epochs=1, eval_async=True)
    out_b = b.fit(checkpoint_dir=tmp_path / "ck", max_epochs=6)
    assert out_b["epochs"] == 6 and b._async_eval_thread is None
    assert torch.equal(b.policy_model.flat_params, sync.policy_model.flat_params)
    ev_b = b.wait_async_eval()
    assert 0 <= ev_b["eval/model_epoch"] <= 5 and ev_b["cnt/total_episodes"] == 64
    epochs_seen = [r["val/eval/model_epoch"] for r in b.metrics_recorder.history if "val/eval/model_epoch" in r]
    assert epochs_seen == sorted(epochs_seen)
    # after shutdown no new evaluation starts
    b._launch_async_eval()
    assert b._async_eval_thread is None
    # ... until the next fit: one more epoch, evaluated in the background again
    out_b2 = b.learn(max_epochs=7)
    assert out_b2["epochs"] == 7 and b.wait_async_eval()["eval/model_epoch"] == 6

    # (3) eval_async switched on after construction has no network copy to evaluate: loud error, no silent shared-weights race
    c = make(eval_freq_epochs=1, eval_async=False)
    c.config.eval_async = True
    with pytest.raises(Exception, match="eval_async"):
        c.validation_epoch()


def test_train_dataloader_drives_training_step_like_lightning():
    """SURVEY 8a row a12 / reference agents/base_agent.py:253-283, 330-366: the index-collate loader built once over ``_trajectories`` feeds
    reference-style gathered minibatches (RolloutTrajectory of flat tensors) into ``training_step``; a later rollout is served by the same
    loader.  Each minibatch's loss equals the oracle's PPO loss of the same gathered batch under the weights of that step."""
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.rollout_buffer import RolloutTrajectory

    cfg = _cfg(n_envs=16, n_steps=32, batch_size=128, n_epochs=2, model_id="mlp_64x64")
    agent = build_agent(cfg, rank=0, world_size=1)
    loader = agent.train_dataloader()
    assert len(loader) == 2 * (16 * 32 // 128) and loader.batch_size == 128
    with pytest.raises(AssertionError, match="only be called once"):
        agent.current_epoch = 1
        agent.train_dataloader()
    agent.current_epoch = 0
    w0 = agent.policy_model.flat_params.clone()
    for epoch in range(2):
        if epoch == 1:
            agent._trajectories = agent.get_rollout_collector("train").collect()
        first_obs = agent._trajectories.observations
        seen = torch.zeros(16 * 32, dtype=torch.int64)
        for i, batch in enumerate(loader):
            assert isinstance(batch, RolloutTrajectory) and batch.observations.shape == (128, 4) and batch.actions.dtype == torch.int64
            if i == 0:
                p = _oracle_params(agent.policy_model)
                want, _ = OP.ppo_loss(p, *(getattr(batch, f).cpu() for f in ("observations", "actions", "logprobs", "values", "advantages", "returns")),
                                      clip_range=cfg.clip_range, clip_range_vf=cfg.clip_range_vf, vf_coef=cfg.vf_coef, ent_coef=cfg.ent_coef,
                                      normalize_adv=cfg.normalize_advantages == "batch")
                got = agent.losses_for_batch(batch, 0)["loss"]
                np.testing.assert_allclose(float(got.item()), float(want), rtol=1e-4, atol=1e-5)
            # which rows of the rollout this minibatch holds (observations are unique per sample with probability 1)
            match = (batch.observations[:, None, :] == first_obs[None, :, :]).all(dim=2)
            seen += match.any(dim=0).cpu().to(torch.int64)
            agent.training_step(batch, i)
        assert i == len(loader) - 1
        assert int(seen.min()) >= 2                    # two passes: every sample of the rollout was gathered at least twice
    assert not torch.equal(w0, agent.policy_model.flat_params)
    m = agent.pop_epoch_metrics()
    assert np.isfinite(m["opt/loss/total"]) and m["opt/grads/norm/all"] > 0
