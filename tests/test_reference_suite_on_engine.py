"""The reference's OWN unit tests for the host-side pieces of the hot path, run UNCHANGED against the engine's modules.

The reference's test files are copied at run time from /root/reference/tests into a temporary directory (never into this repo) next
to a conftest that maps the module names they import (``utils.samplers``, ``utils.dataloaders``, ``utils.datasets``,
``utils.rollouts``, ``utils.distributions``, ``utils.models``, ``utils.policy_factory``, ``trainer_callbacks.hyperparameter_scheduler``,
``gym_wrappers.env_wrapper_registry``, ``agents.hyperparameter_mixin``) onto ``gymnasium_solver_b200``; pytest runs them in a subprocess.  These are the tests SURVEY.md §4
lists as passing on the reference and pinning behaviour the engine must reproduce.

Build container only (skipped where /root/reference is absent: nothing on the GPU box reads the reference).  One test is deselected and
documented: ``test_rollout_buffer.py::test_add_stores_observations_and_dtypes_correctly`` compares the dtype objects of the buffer's
internal arrays with numpy dtypes (``buf.actions_buf.dtype == np.int64``); the engine's buffers are device tensors (``torch.int32``
actions) — the values that test checks are covered by tests/test_rollout_buffer.py.
"""
import os
import shutil
import subprocess
import sys

import pytest

REF_TESTS = "/root/reference/tests"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not present")

FILES = ["test_multipass_random_sampler.py", "test_epoch_shuffling.py", "test_index_dataset.py", "test_rollout_buffer.py",
         "test_masked_categorical.py", "test_models.py", "test_policy_factory_initialization.py", "test_schedulers.py",
         "test_env_wrapper_registry.py", "test_hyperparameter_mixin.py"]
DESELECT = ["test_rollout_buffer.py::test_add_stores_observations_and_dtypes_correctly"]

CONFTEST = '''
import importlib, sys, types
sys.path.insert(0, {repo!r})
E = "gymnasium_solver_b200"

def pkg(name):
    m = types.ModuleType(name); m.__path__ = []; sys.modules[name] = m
    return m

def alias(name, target):
    sys.modules[name] = importlib.import_module(target)

pkg("utils"); pkg("trainer_callbacks"); pkg("gym_wrappers"); pkg("agents")
alias("agents.hyperparameter_mixin", E + ".agents.hyperparameter_mixin")
for name, src in (("samplers", "samplers"), ("dataloaders", "dataloaders"), ("datasets", "dataloaders"), ("rollout_buffer", "rollout_buffer"),
                  ("rollouts", "rollout_buffer"), ("distributions", "distributions"), ("models", "models"), ("policy_factory", "policy_factory"),
                  ("rollout_stats", "rollout_stats"), ("torch", "torch"), ("policy_ops", "policy_ops"), ("model_registry", "model_registry")):
    alias("utils." + name, E + ".utils." + src)
alias("gym_wrappers.env_wrapper_registry", E + ".gym_wrappers.env_wrapper_registry")
hs = types.ModuleType("trainer_callbacks.hyperparameter_scheduler")     # the schedule functions live in utils/schedules.py here
hs.__dict__.update({{k: v for k, v in vars(importlib.import_module(E + ".utils.schedules")).items() if not k.startswith("__")}})
sys.modules["trainer_callbacks.hyperparameter_scheduler"] = hs

def pytest_configure(config):
    for m in ("unit", "integration", "slow"):
        config.addinivalue_line("markers", m)
'''


def test_reference_unit_tests_pass_against_the_engine_modules(tmp_path):
    for f in FILES:
        shutil.copy(os.path.join(REF_TESTS, f), tmp_path / f)
    (tmp_path / "conftest.py").write_text(CONFTEST.format(repo=REPO))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path), "."]
    for d in DESELECT:
        cmd += ["--deselect", d]
    env = {k: v for k, v in os.environ.items() if k != "PYTEST_CURRENT_TEST"}
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, env=env, timeout=600)
    tail = r.stdout[-3000:] + r.stderr[-1000:]
    assert r.returncode == 0, tail
    summary = [l for l in r.stdout.splitlines() if " passed" in l][-1]
    n_passed = int(summary.split(" passed")[0].split()[-1])
    assert n_passed >= 65 and "failed" not in summary and "error" not in summary, summary
