"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol include/gs_engine.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gs_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from gymnasium_solver_b200 import build, _native

    path = build.build()
    assert os.path.exists(path)
    L = _native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"libgs_engine.so does not export {name}"
        assert name in _native.SIGNATURES, f"_native.py does not bind {name}"
    assert L.gs_version() == 100


def test_static_queries_and_error_reporting_without_gpu():
    from gymnasium_solver_b200 import _native as N
    import ctypes as C

    L = N.lib()
    assert [L.gs_env_obs_dim(k) for k in range(3)] == [4, 6, 2]
    assert [L.gs_env_state_dim(k) for k in range(3)] == [4, 4, 2]
    assert [L.gs_env_n_actions(k) for k in range(3)] == [2, 3, 3]
    assert L.gs_env_obs_dim(7) == -1
    m = N.GsMlp()
    m.obs_dim, m.hidden1, m.hidden2, m.n_actions, m.has_value, m.activation = 4, 96, 96, 2, 1, 0
    assert L.gs_mlp_param_count(C.byref(m)) == -1
    assert b"unsupported" in L.gs_last_error()
    h = C.c_void_p()
    assert L.gs_env_create(9, 4, 0, 0, 0, 0, C.byref(h)) != 0
    assert b"unknown env kind" in L.gs_last_error()


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "gymnasium_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} references the oracle"
    # the launcher and the measurement scripts are product-side too: only tests/, smoke() and bench.py's CPU-baseline leg may run the oracle
    extra = [os.path.join(ROOT, "train.py")] + [os.path.join(ROOT, "scripts", f) for f in os.listdir(os.path.join(ROOT, "scripts")) if f.endswith(".py")]
    for path in extra:
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, f"{path} references the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gymnasium_solver_b200 import _native as N

    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(N.EngineError):
        N.lib()
