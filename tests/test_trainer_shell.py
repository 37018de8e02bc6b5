"""Trainer shell (gymnasium_solver_b200/trainer.py): the reference's callback protocol without Lightning.  CPU tests drive the
callbacks with stand-in trainer / module objects exactly as the reference's fit loop would (hook order of SURVEY.md 3.2);
tests/test_gpu_agent.py::test_fit_through_the_trainer_shell_equals_learn runs it on the engine."""
import json
import os
from types import SimpleNamespace

import pytest

from gymnasium_solver_b200.trainer import (Callback, CsvMetricsLogger, EarlyStoppingCallback, HyperparameterSchedulerCallback,
                                           ModelCheckpointCallback, Trainer)


class _Recorder:
    def __init__(self):
        self.rows = []

    def record(self, stage, metrics):
        self.rows.append((stage, dict(metrics)))


def _module(**kw):
    m = SimpleNamespace(metrics_recorder=_Recorder(), reason=None, current_epoch=3, rank=0, saved=[])
    m.set_early_stop_reason = lambda r: setattr(m, "reason", r)
    m.save_checkpoint = lambda d: (os.makedirs(d, exist_ok=True), m.saved.append(str(d)))
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def test_early_stopping_thresholds_modes_and_disabled():
    """trainer_callbacks/early_stopping.py:34-77."""
    tr, mod = SimpleNamespace(should_stop=False, logged_metrics={}), _module()
    cb = EarlyStoppingCallback("val/roll/ep_rew/mean", 475.0)
    cb.on_validation_epoch_end(tr, mod)                      # metric not logged yet: nothing happens
    assert not tr.should_stop
    tr.logged_metrics["val/roll/ep_rew/mean"] = 474.9
    cb.on_train_epoch_end(tr, mod)
    assert not tr.should_stop
    tr.logged_metrics["val/roll/ep_rew/mean"] = 475.0        # inclusive
    cb.on_validation_epoch_end(tr, mod)
    assert tr.should_stop and mod.metrics_recorder.rows == [("val", {"solved": 1})] and ">=" in mod.reason and "val/roll/ep_rew/mean" in mod.reason
    tr2, mod2 = SimpleNamespace(should_stop=False, logged_metrics={"train/loss": 0.1}), _module()
    EarlyStoppingCallback("train/loss", 0.2, mode="min").on_train_epoch_end(tr2, mod2)
    assert tr2.should_stop and "<=" in mod2.reason
    tr3 = SimpleNamespace(should_stop=False, logged_metrics={"train/loss": 0.1})
    EarlyStoppingCallback("train/loss", None, mode="min").on_train_epoch_end(tr3, _module())      # no threshold: disabled
    assert not tr3.should_stop
    with pytest.raises(AssertionError):
        EarlyStoppingCallback("x", 1.0, mode="up")


def test_model_checkpoint_saves_first_best_and_when_stopping(tmp_path):
    """trainer_callbacks/model_checkpoint.py:33-69: first evaluation, a new best, or a stopping trainer save; only the first two
    are marked best."""
    cb = ModelCheckpointCallback(tmp_path / "ck")
    tr, mod = SimpleNamespace(should_stop=False, logged_metrics={}), _module()
    cb.on_validation_epoch_end(tr, mod)                      # metric absent (warm-up): skipped
    assert mod.saved == []
    for epoch, value, stop, saved, best in ((1, 100.0, False, True, "epoch=01"), (2, 90.0, False, False, "epoch=01"), (3, 150.0, False, True, "epoch=03"),
                                            (4, 120.0, True, True, "epoch=03")):
        mod.current_epoch, tr.should_stop = epoch, stop
        tr.logged_metrics["val/roll/ep_rew/mean"] = value
        n0 = len(mod.saved)
        cb.on_validation_epoch_end(tr, mod)
        assert (len(mod.saved) > n0) == saved, epoch
        assert os.readlink(tmp_path / "ck" / "best") == best
    assert os.readlink(tmp_path / "ck" / "last") == "epoch=04" and cb.best_value == 150.0
    assert json.loads((tmp_path / "ck" / "epoch=03" / "metrics.json").read_text())["val/roll/ep_rew/mean"] == 150.0


def test_scheduler_callback_follows_the_reference_fixture(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "host_logic.json")))
    from gymnasium_solver_b200.utils.schedules import position_to_env_steps

    for case in fx["schedules"]:
        got = []
        cb = HyperparameterSchedulerCallback(schedule=case["kind"], parameter="x", start_value=case["start_value"], end_value=case["end_value"],
                                             start_step=position_to_env_steps(case["start"], param="x", default_to_max=False, max_env_steps=100000),
                                             end_step=position_to_env_steps(case["end"], param="x", default_to_max=True, max_env_steps=100000),
                                             warmup_fraction=case["warmup"], set_value_fn=lambda m, v: got.append(v))
        for e in case["env_steps"]:      # 2 ranks x e/2 local env steps
            cb.on_train_epoch_end(None, SimpleNamespace(world_size=2, get_rollout_collector=lambda stage, e=e: SimpleNamespace(total_steps=e / 2)))
        assert got == pytest.approx(case["values"], rel=1e-12, abs=1e-18)
    with pytest.raises(ValueError, match="end_step"):
        HyperparameterSchedulerCallback(schedule="linear", parameter="x", start_value=1, end_value=0, start_step=5, end_step=1)


def test_trainer_drives_hooks_in_the_reference_order(tmp_path):
    """per epoch callbacks.on_train_epoch_start -> agent collect -> passes -> callbacks.on_train_epoch_end -> (every
    eval_freq_epochs) validation hooks; the budget check stops the loop before a rollout is collected."""
    log = []

    class Spy(Callback):
        def on_fit_start(self, t, m): log.append("cb.fit_start")
        def on_train_epoch_start(self, t, m): log.append(f"cb.epoch_start[{m.current_epoch}]")
        def on_train_epoch_end(self, t, m):
            log.append(f"cb.epoch_end[{m.current_epoch}]")
            m.log_dict({"train/roll/ep_rew/mean": 10.0 * (m.current_epoch + 1)})
        def on_validation_epoch_start(self, t, m): log.append("cb.val_start")
        def on_validation_epoch_end(self, t, m): log.append("cb.val_end")
        def on_fit_end(self, t, m): log.append("cb.fit_end")

    col = SimpleNamespace(total_steps=0, resolve_episodes_async=lambda: log.append("resolve"))
    agent = SimpleNamespace(config=SimpleNamespace(max_epochs=None, eval_freq_epochs=2, eval_warmup_epochs=0), current_epoch=0, world_size=1,
                            best_eval_reward=float("-inf"), _early_stop_reason="", _fit_t0=0.0, _trajectories=None, trainer=None)
    agent.on_fit_start = lambda: log.append("agent.fit_start")
    agent.get_rollout_collector = lambda stage: col

    def epoch_start():
        if col.total_steps >= 3 * 64:
            agent._early_stop_reason = "budget"
            return False
        col.total_steps += 64
        agent._trajectories = f"traj{agent.current_epoch}"
        log.append("agent.collect")
        return True

    agent.on_train_epoch_start = epoch_start
    agent.train_on_rollout = lambda traj: log.append(f"agent.train({traj})")
    agent.validation_epoch = lambda: log.append("agent.validate")
    agent.set_early_stop_reason = lambda r: setattr(agent, "_early_stop_reason", r)
    csv_path = tmp_path / "metrics.csv"
    tr = Trainer(callbacks=[Spy()], loggers=[CsvMetricsLogger(csv_path)])
    agent.log_dict = tr.log_dict
    out = tr.fit(agent)
    assert log == ["agent.fit_start", "cb.fit_start",
                   "cb.epoch_start[0]", "agent.collect", "agent.train(traj0)", "resolve", "cb.epoch_end[0]",
                   "cb.epoch_start[1]", "agent.collect", "agent.train(traj1)", "resolve", "cb.epoch_end[1]", "cb.val_start", "agent.validate", "cb.val_end",
                   "cb.epoch_start[2]", "agent.collect", "agent.train(traj2)", "resolve", "cb.epoch_end[2]",
                   "cb.epoch_start[3]", "cb.fit_end"]
    assert out["epochs"] == 3 and out["stop_reason"] == "budget" and tr.logged_metrics["train/roll/ep_rew/mean"] == 30.0
    rows = csv_path.read_text().strip().splitlines()
    assert rows[0] == "epoch,train/roll/ep_rew/mean" and rows[1:] == ["0,10.0", "1,20.0", "2,30.0"]
    # a callback that sets should_stop ends the loop after the current epoch
    class Stop(Callback):
        def on_train_epoch_end(self, t, m): t.should_stop = True
    col.total_steps, agent.current_epoch = 0, 0
    assert Trainer(callbacks=[Stop()], max_epochs=10).fit(agent)["epochs"] == 1


def test_trainer_reports_a_training_failure_through_on_exception():
    """Lightning's on_exception protocol (reference agents/base_agent.py:509-511 joins its background evaluation there): callbacks first,
    then the module, then the error propagates; on_fit_end hooks do not run; a callback without the hook is skipped."""
    log = []

    class Spy(Callback):
        def on_exception(self, t, m, exc): log.append(f"cb.exception({exc})")
        def on_fit_end(self, t, m): log.append("cb.fit_end")

    class Bare:                                             # a Lightning-style callback that defines only what it needs
        def on_train_epoch_start(self, t, m): log.append("bare.epoch_start")

    col = SimpleNamespace(total_steps=0, resolve_episodes_async=lambda: None)
    agent = SimpleNamespace(config=SimpleNamespace(max_epochs=None, eval_freq_epochs=None, eval_warmup_epochs=0), current_epoch=0, world_size=1,
                            best_eval_reward=float("-inf"), _early_stop_reason="", _fit_t0=0.0, _trajectories=None, trainer=None)
    agent.on_fit_start = lambda: None
    agent.get_rollout_collector = lambda stage: col
    agent.on_train_epoch_start = lambda: True
    agent.on_fit_end = lambda: log.append("agent.fit_end")
    agent.on_exception = lambda t, m, exc: log.append(f"agent.exception({exc})")

    def train(traj):
        if agent.current_epoch == 2:
            raise RuntimeError("kernel failed")

    agent.train_on_rollout = train
    with pytest.raises(RuntimeError, match="kernel failed"):
        Trainer(callbacks=[Spy(), Bare()]).fit(agent)
    assert log == ["bare.epoch_start"] * 3 + ["cb.exception(kernel failed)", "agent.exception(kernel failed)"]
