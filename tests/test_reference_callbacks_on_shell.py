"""The reference's OWN callback classes (trainer_callbacks/early_stopping.py, hyperparameter_scheduler.py, dispatch_metrics.py, loaded from /root/reference
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
    wb = types.ModuleType("wandb")                        # dispatch_metrics.py only asks whether a W&B run is active
    wb.run = None
    monkeypatch.setitem(sys.modules, "wandb", wb)
    mods = {}
    for name in ("early_stopping", "hyperparameter_scheduler", "dispatch_metrics"):
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


def test_reference_dispatch_metrics_callback_runs_on_the_engine_trainer(ref_callbacks):
    """trainer_callbacks/dispatch_metrics.py unchanged: it needs ``pl_module.timings`` (TimingsTracker markers set by the agent's hooks),
    ``metrics_recorder`` (reset_epoch / compute_epoch_means / update_history), ``get_rollout_collector(stage).get_metrics()``,
    ``calc_training_progress`` and ``log_dict`` -- here the engine's own MetricsRecorder / TimingsTracker classes on a host-only agent."""
    import time

    from gymnasium_solver_b200.agents.base_agent import MetricsRecorder
    from gymnasium_solver_b200.trainer import Callback, Trainer
    from gymnasium_solver_b200.utils.timings_tracker import TimingsTracker

    Dispatch = ref_callbacks["dispatch_metrics"].DispatchMetricsCallback
    Early = ref_callbacks["early_stopping"].EarlyStoppingCallback
    assert issubclass(Dispatch, Callback)
    agent, col = _fake_agent(n_envs=8, n_steps=16, max_env_steps=8 * 16 * 6)
    agent.config.eval_freq_epochs, agent.config.eval_episodes = 2, 4
    agent.metrics_recorder, agent.timings = MetricsRecorder(), TimingsTracker()
    col.get_metrics = lambda: {"cnt/total_env_steps": col.total_steps, "cnt/total_vec_steps": col.total_vec_steps, "roll/ep_rew/mean": 10.0 * col.total_vec_steps,
                               "action_dist": [3, 4]}
    agent.calc_training_progress = lambda: col.total_steps / agent.config.max_env_steps
    counters = lambda: {"cnt/total_env_steps": col.total_steps, "cnt/total_vec_steps": col.total_vec_steps, "cnt/epoch": agent.current_epoch}
    agent.on_fit_start = lambda: agent.timings.start("on_fit_start", values=counters())
    inner = agent.on_train_epoch_start

    def epoch_start():
        agent.timings.start("on_train_epoch_start", values=counters())
        return inner()

    agent.on_train_epoch_start = epoch_start

    def train(traj):
        time.sleep(0.002)
        agent.metrics_recorder.record("train", {"opt/loss/total": 1.0 / (agent.current_epoch + 1)})

    agent.train_on_rollout = train
    agent.validation_epoch = lambda: agent.metrics_recorder.record("val", {"roll/ep_rew/mean": 100.0 * (agent.current_epoch + 1), "cnt/total_episodes": 4})

    def log_dict(metrics):
        agent.trainer.log_dict(metrics)

    agent.log_dict = log_dict
    tr = Trainer(callbacks=[Dispatch(), Early("val/roll/ep_rew/mean", 400.0)])
    out = tr.fit(agent)
    lm = tr.logged_metrics
    # validation at epochs 1, 3 (every 2nd): 100 * (epoch + 1) >= 400 at epoch 3 -> the reference's early stop ends the run after 4 epochs
    assert out["epochs"] == 4 and "val/roll/ep_rew/mean" in out["stop_reason"]
    assert lm["train/cnt/total_env_steps"] == 8 * 16 * 4 and lm["train/cnt/epoch"] == 3 and lm["train/opt/loss/total"] == pytest.approx(0.25)
    assert lm["train/progress"] == pytest.approx(4 / 6) and "train/action_dist" not in lm
    assert lm["val/roll/ep_rew/mean"] == pytest.approx(400.0) and lm["val/cnt/total_episodes"] == 4
    # rates come from the agent's markers: env steps since on_fit_start / on_train_epoch_start per second
    assert 0 < lm["train/sys/timing/fps"] < 8 * 16 * 4 / 0.008 and 0 < lm["train/sys/timing/fps_instant"] <= 8 * 16 / 0.002
    assert lm["train/sys/timing/eps"] > 0 and lm["train/sys/timing/eta_s"] == pytest.approx(8 * 16 * 6 / lm["train/sys/timing/fps"])
    hist = agent.metrics_recorder.history                   # update_history: one snapshot per dispatched stage
    assert [("train/cnt/epoch" in h, "val/cnt/epoch" in h) for h in hist] == [(True, False), (True, False), (False, True), (True, False), (True, False), (False, True)]


def test_timings_tracker_equals_the_reference_class(monkeypatch):
    """utils/timings_tracker.py:22-74 executed beside the engine's TimingsTracker on one scripted clock: same seconds, same rates, same
    treatment of non-numeric / missing / decreasing counters, KeyError for an unknown marker."""
    import time

    from gymnasium_solver_b200.utils.timings_tracker import TimingsTracker

    spec = importlib.util.spec_from_file_location("_ref_timings_tracker", os.path.join(REF, "utils", "timings_tracker.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = ref_mod                       # dataclasses resolve annotations through sys.modules
    try:
        spec.loader.exec_module(ref_mod)
        clock = {"ns": 1_000_000_000}
        monkeypatch.setattr(time, "perf_counter_ns", lambda: clock["ns"])
        ours, ref = TimingsTracker(), ref_mod.TimingsTracker()
        start = {"cnt/total_env_steps": 1024, "cnt/epoch": 2, "roll/fps": 3.5, "action_dist": [1, 2], "name": "x", "none": None, "flag": True}
        for t in (ours, ref):
            t.start("fit", values=start)
            t.start("bare", values={})             # (the reference requires a mapping; the engine's class also accepts none)
        assert ours.seconds_since("fit") == ref.seconds_since("fit") == 1e-12          # zero elapsed time is clamped, never 0
        clock["ns"] += 2_500_000_000
        now = {"cnt/total_env_steps": 5120, "cnt/epoch": 1, "new/counter": 10, "roll/fps": 7.0, "action_dist": [3], "flag": True}
        assert ours.seconds_since("fit") == ref.seconds_since("fit") == 2.5
        assert ours.throughput_since("fit", values_now=now) == ref.throughput_since("fit", values_now=now)
        assert ours.throughput_since("fit", values_now=now)["cnt/total_env_steps"] == (5120 - 1024) / 2.5
        assert ours.throughput_since("fit", values_now=now)["cnt/epoch"] == 0.0        # a counter that went down
        assert ours.throughput_since("bare", values_now=now) == ref.throughput_since("bare", values_now=now)
        for t in (ours, ref):
            with pytest.raises(KeyError):
                t.seconds_since("never started")
    finally:
        sys.modules.pop(spec.name, None)
