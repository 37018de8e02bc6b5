"""GPU parity: gs_gae / gs_mc_returns / gs_valid_index_map / gs_moments / gs_normalize vs the numpy oracle and the
reference-generated golden fixtures.  fp32 scans are BIT-EXACT (same rounding sequence as numpy)."""
import glob
import os

import numpy as np
import pytest

from oracle import returns as R

pytestmark = pytest.mark.gpu


def _rand_case(rng, T, N, pdone=0.05, ptimeout=0.3):
    values = rng.standard_normal((T, N)).astype(np.float32)
    rewards = rng.standard_normal((T, N)).astype(np.float32)
    dones = rng.random((T, N)) < pdone
    timeouts = dones & (rng.random((T, N)) < ptimeout)
    last_values = rng.standard_normal(N).astype(np.float32)
    boot = np.where(timeouts, rng.standard_normal((T, N)), 0).astype(np.float32)
    return values, rewards, dones, timeouts, last_values, boot


def test_gae_golden_fixtures_bit_exact(golden_dir):
    import engine_api as E

    files = sorted(glob.glob(os.path.join(golden_dir, "returns_*.npz")))
    assert files
    for f in files:
        d = np.load(f)
        adv, ret = E.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], d["boot"], float(d["gamma"]), float(d["lam"]))
        np.testing.assert_array_equal(adv, d["adv"], err_msg=f)
        np.testing.assert_array_equal(ret, d["ret"], err_msg=f)
        if "adv0" in d.files:
            a0, r0 = E.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], np.zeros_like(d["boot"]), float(d["gamma"]), float(d["lam"]))
            np.testing.assert_array_equal(a0, d["adv0"], err_msg=f)
            np.testing.assert_array_equal(r0, d["ret0"], err_msg=f)


@pytest.mark.parametrize("T,N", [(1, 1), (1, 77), (3, 1), (7, 5), (8, 64), (9, 65), (128, 1000), (33, 4097), (128, 65536)])
def test_gae_vs_oracle_bit_exact(T, N):
    import engine_api as E

    rng = np.random.default_rng(T * 100003 + N)
    v, r, d, to, lv, boot = _rand_case(rng, T, N)
    for b in (boot, None):
        adv, ret = E.gae(v, r, d, to, lv, b, 0.99, 0.95)
        oadv, oret = R.gae(v, r, d, to, lv, b, 0.99, 0.95)
        np.testing.assert_array_equal(adv, oadv)
        np.testing.assert_array_equal(ret, oret)


@pytest.mark.parametrize("T,N", [(5, 3), (128, 1000), (33, 4097)])
def test_gae_zero_bootstrap_variant_is_bit_identical_to_a_zero_array(T, N):
    """gs_gae_zero_boot (what the collector calls: the bootstrap array of the reference is identically zero under NEXT_STEP
    autoreset and is not read) == gs_gae given a dense zero array == the numpy oracle, bit for bit."""
    import engine_api as E

    rng = np.random.default_rng(T * 1000 + N)
    v, r = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    d = rng.random((T, N)) < 0.1
    to = d & (rng.random((T, N)) < 0.4)
    lv = rng.standard_normal(N).astype(np.float32)
    zero = np.zeros((T, N), np.float32)
    a0, r0 = E.gae(v, r, d, to, lv, zero, 0.99, 0.95)
    a1, r1 = E.gae(v, r, d, to, lv, "zero", 0.99, 0.95)
    oa, orr = R.gae(v, r, d, to, lv, zero, 0.99, 0.95)
    np.testing.assert_array_equal(a1, a0)
    np.testing.assert_array_equal(r1, r0)
    np.testing.assert_array_equal(a1, oa)
    np.testing.assert_array_equal(r1, orr)


def test_gae_reference_known_answers():
    import engine_api as E

    # reference tests/test_rollout_collector.py:93-106
    g, lam = 0.99, 0.95
    dones = np.array([[False], [True], [False]])
    boot = np.where(dones, 420.0, 0.0).astype(np.float32)
    adv, _ = E.gae(np.zeros((3, 1), np.float32), np.zeros((3, 1), np.float32), dones, dones, np.zeros(1, np.float32), boot, g, lam)
    np.testing.assert_allclose(adv.ravel(), [g * g * lam * 420.0, g * 420.0, 0.0], rtol=1e-6)


def test_gae_linearity_property_full_size():
    """Size-independent property at the BASELINE size: with no terminals GAE is linear in (rewards, values)."""
    import engine_api as E

    T, N = 128, 65536
    rng = np.random.default_rng(5)
    v1, r1, _, _, lv1, _ = _rand_case(rng, T, N, pdone=0.0)
    z = np.zeros((T, N), bool)
    a1, _ = E.gae(v1, r1, z, z, lv1, None, 0.99, 0.95)
    a2, _ = E.gae(2 * v1, 2 * r1, z, z, 2 * lv1, None, 0.99, 0.95)
    np.testing.assert_allclose(a2, 2 * a1, rtol=1e-5, atol=1e-5)  # scaling by 2 is exact in fp32
    np.testing.assert_array_equal(a2, 2 * a1)


@pytest.mark.parametrize("T,N", [(1, 3), (4, 1), (17, 5), (64, 33), (128, 4096)])
@pytest.mark.parametrize("treat_timeouts_as_terminals", [True, False])
def test_mc_returns_episode_and_valid_map(T, N, treat_timeouts_as_terminals):
    import engine_api as E

    rng = np.random.default_rng(T * 7 + N)
    _, r, d, to, _, _ = _rand_case(rng, T, N, pdone=0.12)
    to_arg = None if treat_timeouts_as_terminals else to
    to_np = np.zeros_like(to) if treat_timeouts_as_terminals else to
    ret, lt = E.mc_returns(r, d, to_arg, 0.99, episode_mode=False)
    np.testing.assert_array_equal(ret, R.mc_returns(r, d, to_np, 0.99))
    np.testing.assert_array_equal(lt, R.last_terminal(d, to_np))
    ret_ep, _ = E.mc_returns(r, d, to_arg, 0.99, episode_mode=True)
    np.testing.assert_array_equal(ret_ep, R.to_full_episode(R.mc_returns(r, d, to_np, 0.99), d, to_np))
    mask, imap, nv = E.valid_index_map(lt, T)
    omask, omap = R.valid_mask_and_index_map(d, to_np)
    if omask is None:
        assert nv == 0
    else:
        assert nv == int(omask.sum())
        np.testing.assert_array_equal(mask, omask)
        np.testing.assert_array_equal(imap, omap)


def test_valid_map_reference_known_answers_and_sparse_terminals():
    import engine_api as E

    # reference tests/test_rollouts_extra.py:28-49
    lt = np.array([1, -1], np.int32)
    mask, imap, nv = E.valid_index_map(lt, 5)
    np.testing.assert_array_equal(mask, [1, 1, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(imap, [0, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    assert nv == 2
    # no terminal at all -> n_valid == 0; the reference returns (None, None) and trains on every sample WITHOUT a remap
    # (returns_advantages.py:49-50): the map must be the identity, not "everything -> sample 0"
    mask, imap, nv = E.valid_index_map(np.full(7, -1, np.int32), 4)
    assert nv == 0 and not mask.any()
    np.testing.assert_array_equal(imap, np.arange(28))
    # ... and the rollout's running statistics then cover every element (rollout_collector.py:435-455), the baseline none
    x = np.arange(28, dtype=np.float32).reshape(4, 7)
    np.testing.assert_allclose(E.moments_valid(x, np.full(7, -1, np.int32)), [x.sum(), (x.astype(np.float64) ** 2).sum(), 28])
    np.testing.assert_allclose(E.moments(x, np.full(7, -1, np.int32)), [0, 0, 0])
    lt = np.array([2, -1, 0, -1, -1, -1, -1], np.int32)
    np.testing.assert_allclose(E.moments_valid(x, lt), E.moments(x, lt))
    # long runs of envs without terminals, crossing the 1024-env scan blocks
    N, T = 5000, 6
    rng = np.random.default_rng(0)
    d = np.zeros((T, N), bool)
    for e in (1500, 1501, 4100):
        d[rng.integers(0, T), e] = True
    lt = R.last_terminal(d, np.zeros_like(d))
    mask, imap, nv = E.valid_index_map(lt, T)
    omask, omap = R.valid_mask_and_index_map(d, np.zeros_like(d))
    np.testing.assert_array_equal(mask, omask)
    np.testing.assert_array_equal(imap, omap)


def test_mc_reference_known_answers():
    import engine_api as E

    rewards = np.array([[1.0], [2.0], [3.0], [4.0]], np.float32)
    dones = np.array([[False], [True], [False], [False]])
    np.testing.assert_array_equal(E.mc_returns(rewards, dones, None, 1.0)[0].ravel(), [3, 2, 7, 4])
    np.testing.assert_array_equal(E.mc_returns(rewards, dones, dones, 1.0)[0].ravel(), [10, 9, 7, 4])


def test_moments_normalize_and_baseline_shift():
    import engine_api as E

    rng = np.random.default_rng(3)
    x = (rng.standard_normal((128, 1000)) * 3 + 1.5).astype(np.float32)
    m = E.moments(x)
    np.testing.assert_allclose(m, [x.astype(np.float64).sum(), (x.astype(np.float64) ** 2).sum(), x.size], rtol=1e-12)
    np.testing.assert_allclose(E.normalize(x), R.normalize(x), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(E.normalize(x, shift_only=True), x - x.mean(dtype=np.float64), rtol=1e-5, atol=1e-6)
    # masked by last_terminal: RunningStats over valid returns (rollout_collector.py:415-418)
    lt = rng.integers(-1, 128, 1000).astype(np.int32)
    valid = np.arange(128)[:, None] <= lt[None, :]
    m = E.moments(x, lt)
    rs = R.RunningStats()
    rs.update(x[valid])
    assert m[2] == rs.count
    np.testing.assert_allclose(m[0] / m[2], rs.mean(), rtol=1e-5)
    np.testing.assert_allclose(np.sqrt(m[1] / m[2] - (m[0] / m[2]) ** 2), rs.std(), rtol=1e-4)
    # reference tests/test_rollouts_extra.py:69-75
    adv = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], np.float32)
    flat = E.normalize(adv).ravel()
    assert abs(float(flat.mean())) < 1e-6 and abs(float(flat.std()) - 1.0) < 1e-5
