"""The reference's OWN callback classes (trainer_callbacks/early_stopping.py, hyperparameter_scheduler.py, loaded from /root/reference
by path with a `pytorch_lightning` module that maps Callback / Trainer onto the engine's trainer shell) driven by
gymnasium_solver_b200.trainer.Trainer: evidence that the shell speaks the protocol those callbacks were written against.

Runs only where /root/reference exists (the build container); skipped elsewhere — nothing on the GPU box reads the reference."""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture()
def ref_callbacks(monkeypatch):
    from gymnasium_solver_b200 import trainer as shell

    pl = types.ModuleType("pytorch_lightning")
    pl.Callback, pl.Trainer, pl.LightningModule = shell.Callback, shell.Trainer, object
    monkeypatch.setitem(sys.modules, "pytorch_lightning", pl)
    fmt = types.ModuleType("utils.formatting")            # early_stopping.py only formats its message with it
    fmt.format_metric_value = lambda key, value: f"{value:.2f}"
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = []
    monkeypatch.setitem(sys.modules, "utils", utils_pkg)
    monkeypatch.setitem(sys.modules, "utils.formatting", fmt)
    mods = {}
    for name in ("early_stopping", "hyperparameter_scheduler"):
        spec = importlib.util.spec_from_file_location(f"_ref_cb_{name}", os.path.join(REF, "trainer_callbacks", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def _fake_agent(n_envs=8, n_steps=16, max_env_steps=8 * 16 * 20):
    col = SimpleNamespace(total_steps=0, total_vec_steps=0, resolve_episodes_async=lambda: None)
    agent = SimpleNamespace(config=SimpleNamespace(max_epochs=None, eval_freq_epochs=None, eval_warmup_epochs=0, max_env_steps=max_env_steps),
                            current_epoch=0, world_size=1, best_eval_reward=float("-inf"), _early_stop_reason="", _fit_t0=0.0, _trajectories=None,
                            trainer=None, policy_lr=None, hp={}, metrics_recorder=SimpleNamespace(record=lambda stage, m: agent.hp.setdefault("recorded", []).append((stage, m))))
    agent.on_fit_start = lambda: None
    agent.get_rollout_collector = lambda stage: col

    def epoch_start():
        if col.total_steps + n_envs * n_steps > max_env_steps:
            return False
        col.total_steps += n_envs * n_steps
        col.total_vec_steps += n_steps
        return True

    agent.on_train_epoch_start = epoch_start
    agent.train_on_rollout = lambda traj: None
    agent.set_hyperparameter = lambda name, value: agent.hp.__setitem__(name, value)
    agent.set_early_stop_reason = lambda r: setattr(agent, "_early_stop_reason", r)
    return agent, col


def test_reference_scheduler_and_early_stopping_run_on_the_engine_trainer(ref_callbacks):
    from gymnasium_solver_b200.trainer import Callback, Trainer
    from gymnasium_solver_b200.utils.schedules import cosine

    Sched = ref_callbacks["hyperparameter_scheduler"].HyperparameterSchedulerCallback
    Early = ref_callbacks["early_stopping"].EarlyStoppingCallback
    assert issubclass(Sched, Callback) and issubclass(Early, Callback)
    agent, col = _fake_agent()

    class Log(Callback):                                   # stands in for DispatchMetricsCallback: a rising training curve
        def on_train_epoch_end(self, trainer, pl_module):
            trainer.log_dict({"train/roll/ep_rew/mean": 50.0 * (pl_module.current_epoch + 1)})

    # the reference positions schedules in vec steps: end = max_env_steps / n_envs
    sched = Sched(schedule="cosine", parameter="policy_lr", start_value=1e-3, end_value=1e-4, start_step=0.0, end_step=16.0 * 20)
    tr = Trainer(callbacks=[Log(), sched, Early("train/roll/ep_rew/mean", 475.0)])
    out = tr.fit(agent)
    # 50 * (epoch + 1) >= 475 first at epoch 9: ten epochs ran, the reference's callback stopped the engine's trainer and reported why
    assert out["epochs"] == 10 and tr.should_stop
    assert "train/roll/ep_rew/mean" in out["stop_reason"] and ">=" in out["stop_reason"]
    assert agent.hp["recorded"] == [("train", {"solved": 1})]
    assert agent.hp["policy_lr"] == pytest.approx(cosine(1e-3, 1e-4, 10 / 20), rel=1e-12)       # the reference's own scheduler set it
    # without the early stop the budget check of the agent ends the run at 20 epochs and the schedule reaches its end value
    agent2, _ = _fake_agent()
    out2 = Trainer(callbacks=[Sched(schedule="linear", parameter="clip_range", start_value=0.2, end_value=0.05, start_step=0.0, end_step=16.0 * 20)]).fit(agent2)
    assert out2["epochs"] == 20 and agent2.hp["clip_range"] == pytest.approx(0.05)
