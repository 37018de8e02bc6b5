"""GPU parity: gs_policy_act / gs_policy_values vs the torch fp32 oracle and the reference-generated fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import policy as P

pytestmark = pytest.mark.gpu

CONFIGS = [("cartpole64", 4, (64, 64), 2), ("acrobot128", 6, (128, 128), 3), ("mcar256", 2, (256, 256), 3), ("tiny64", 4, (64,), 2)]


def _params(d, prefix):
    return {k: torch.from_numpy(d[prefix + k]) for k in P.PARAM_ORDER if prefix + k in d.files}


@pytest.mark.parametrize("tag,D,hidden,A", CONFIGS)
def test_policy_act_matches_reference_fixture(golden_dir, tag, D, hidden, A):
    import engine_api as E

    d = np.load(os.path.join(golden_dir, f"policy_{tag}.npz"))
    p = _params(d, "p_")
    a, lp, v, lg = E.policy_act(E.dev_params(p), d["obs"], deterministic=True)
    np.testing.assert_array_equal(a, d["act_det"])
    np.testing.assert_allclose(lp, d["logp_det"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v, d["value_det"], rtol=1e-5, atol=1e-6)
    logp_all = lg - torch.from_numpy(lg).logsumexp(-1, keepdim=True).numpy()
    np.testing.assert_allclose(logp_all, d["logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(E.policy_values(E.dev_params(p), d["obs"]), d["value"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag,D,hidden,A", CONFIGS)
@pytest.mark.parametrize("n", [1, 63, 64, 65, 5000])
@pytest.mark.parametrize("activation", ["relu", "tanh"])
def test_policy_act_sampling_with_injected_uniforms(tag, D, hidden, A, n, activation):
    import engine_api as E

    p = P.random_params(D, hidden, A, seed=n)
    g = torch.Generator().manual_seed(n + 1)
    obs = torch.randn(n, D, generator=g)
    u = torch.rand(n, generator=g)
    oa, olp, ov, olg = P.act(p, obs, uniforms=u, activation=activation)
    a, lp, v, lg = E.policy_act(E.dev_params(p), obs.numpy(), uniforms=u.numpy(), activation=activation)
    np.testing.assert_allclose(lg, olg.numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(v, ov.numpy(), rtol=1e-5, atol=2e-6)
    # actions agree except where u sits within rounding distance of a CDF edge
    cdf = (olg - olg.logsumexp(-1, keepdim=True)).exp().cumsum(-1)
    safe = ((cdf - u[:, None]).abs().min(dim=-1).values > 1e-5).numpy()
    np.testing.assert_array_equal(a[safe], oa.numpy()[safe])
    assert safe.mean() > 0.99
    np.testing.assert_allclose(lp[safe], olp.numpy()[safe], rtol=1e-5, atol=2e-6)


def test_policy_only_model_returns_zero_values_and_philox_sampling_is_unbiased():
    import engine_api as E

    p = P.random_params(4, (64, 64), 2, has_value=False, seed=3)
    obs = torch.zeros(200000, 4) + 0.1
    a, lp, v, lg = E.policy_act(E.dev_params(p), obs.numpy(), seed=123, offset=5)
    assert (v == 0).all()
    probs = torch.softmax(torch.from_numpy(lg[0]), -1).numpy()
    freq = np.bincount(a, minlength=2) / a.size
    np.testing.assert_allclose(freq, probs, atol=4e-3)  # ~4 sigma at n = 2e5
    # counter-based: same (seed, offset, row) -> same draw; different offset -> different draws
    a2, *_ = E.policy_act(E.dev_params(p), obs.numpy(), seed=123, offset=5)
    a3, *_ = E.policy_act(E.dev_params(p), obs.numpy(), seed=123, offset=6)
    np.testing.assert_array_equal(a, a2)
    assert (a != a3).mean() > 0.2


def test_unsupported_shapes_fail_loudly():
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    p = E.dev_params(P.random_params(4, (96, 96), 2))
    with pytest.raises(N.EngineError, match="unsupported"):
        E.policy_act(p, np.zeros((4, 4), np.float32))
    p = E.dev_params(P.random_params(4, (64, 64), 5))
    with pytest.raises(N.EngineError, match="n_actions"):
        E.policy_act(p, np.zeros((4, 4), np.float32))
