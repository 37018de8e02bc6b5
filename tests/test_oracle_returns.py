"""Pins oracle/returns.py against (a) the reference's own known answers and (b) fixtures produced by running
the reference's utils/returns_advantages.py (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from oracle import returns as R


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "returns_*.npz")))


def test_reference_kat_idx_map_basic():
    # reference tests/test_rollouts_extra.py:17-24 (flat mask == single env column)
    mask = np.array([False, False, True, False, True, False])
    pos = np.arange(mask.size)
    filled = np.maximum.accumulate(np.where(mask, pos, -1))
    filled[filled < 0] = int(np.argmax(mask))
    np.testing.assert_array_equal(filled, [2, 2, 2, 2, 4, 4])


def test_reference_kat_valid_mask_multi_env():
    # reference tests/test_rollouts_extra.py:28-49
    T, N = 5, 2
    dones = np.zeros((T, N), bool)
    dones[1, 0] = True
    vm, im = R.valid_mask_and_index_map(dones, np.zeros((T, N), bool))
    np.testing.assert_array_equal(vm, [1, 1, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(im, [0, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    assert R.valid_mask_and_index_map(np.zeros((T, N), bool), np.zeros((T, N), bool)) == (None, None)


def test_reference_kat_mc_returns_timeout_toggle():
    # reference tests/test_rollouts_extra.py:53-65
    rewards = np.array([[1.0], [2.0], [3.0], [4.0]], np.float32)
    dones = np.array([[False], [True], [False], [False]])
    timeouts = np.array([[False], [True], [False], [False]])
    np.testing.assert_array_equal(R.mc_returns(rewards, dones, np.zeros_like(timeouts), 1.0).ravel(), [3, 2, 7, 4])
    np.testing.assert_array_equal(R.mc_returns(rewards, dones, timeouts, 1.0).ravel(), [10, 9, 7, 4])


def test_reference_kat_gae_timeout_bootstrap():
    # reference tests/test_rollout_collector.py:93-106: values 0, rewards 0, timeout at t=1 bootstrapped with 420
    g, lam = 0.99, 0.95
    T = 3
    values = np.zeros((T, 1), np.float32)
    rewards = np.zeros((T, 1), np.float32)
    dones = np.array([[False], [True], [False]])
    timeouts = dones.copy()
    boot = np.where(timeouts, 420.0, 0.0).astype(np.float32)
    adv, _ = R.gae(values, rewards, dones, timeouts, np.zeros(1, np.float32), boot, g, lam)
    np.testing.assert_allclose(adv.ravel(), [g * g * lam * 420.0, g * 420.0, 0.0], rtol=1e-6)


def test_reference_kat_normalize():
    # reference tests/test_rollouts_extra.py:69-75
    adv = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], np.float32)
    flat = R.normalize(adv).ravel()
    assert abs(float(flat.mean())) < 1e-6 and abs(float(flat.std()) - 1.0) < 1e-5


def test_survey_known_answers(golden_dir):
    d = np.load(os.path.join(golden_dir, "returns_survey.npz"))
    np.testing.assert_allclose(d["adv"], [[1.44545, 3.5781355], [0.9, 2.2127967], [1.4534738, 1.6085025], [1.6474999, -0.395]], rtol=1e-6)
    adv, ret = R.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], d["boot"], float(d["gamma"]), float(d["lam"]))
    np.testing.assert_array_equal(adv, d["adv"])
    np.testing.assert_array_equal(ret, d["ret"])


def test_golden_bit_exact(golden_dir):
    files = [f for f in _cases(golden_dir) if not f.endswith("survey.npz")]
    assert len(files) >= 5
    for f in files:
        d = np.load(f)
        g, lam = float(d["gamma"]), float(d["lam"])
        zeros = np.zeros_like(d["timeouts"])
        adv, ret = R.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], d["boot"], g, lam)
        np.testing.assert_array_equal(adv, d["adv"], err_msg=f)
        np.testing.assert_array_equal(ret, d["ret"], err_msg=f)
        adv0, ret0 = R.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], None, g, lam)
        # boot given as zeros overrides next_values at timeouts; None leaves them — different on purpose
        a0, r0 = R.gae(d["values"], d["rewards"], d["dones"], d["timeouts"], d["last_values"], np.zeros_like(d["boot"]), g, lam)
        np.testing.assert_array_equal(a0, d["adv0"], err_msg=f)
        np.testing.assert_array_equal(r0, d["ret0"], err_msg=f)
        assert adv0.shape == a0.shape and ret0.shape == r0.shape
        mt = R.mc_returns(d["rewards"], d["dones"], zeros, g)
        mk = R.mc_returns(d["rewards"], d["dones"], d["timeouts"], g)
        np.testing.assert_array_equal(mt, d["mc_term"], err_msg=f)
        np.testing.assert_array_equal(mk, d["mc_keep"], err_msg=f)
        np.testing.assert_array_equal(R.to_full_episode(mt, d["dones"], zeros), d["ep_term"], err_msg=f)
        np.testing.assert_array_equal(R.to_full_episode(mk, d["dones"], d["timeouts"]), d["ep_keep"], err_msg=f)
        for tag, to in (("t", zeros), ("k", d["timeouts"])):
            vm, im = R.valid_mask_and_index_map(d["dones"], to)
            if bool(d[f"vm_{tag}_none"]):
                assert vm is None and im is None
            else:
                np.testing.assert_array_equal(vm, d[f"vm_{tag}"], err_msg=f)
                np.testing.assert_array_equal(im, d[f"im_{tag}"], err_msg=f)
        np.testing.assert_allclose(R.normalize(d["adv"]), d["norm_adv"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(R.normalize(d["ret"]), d["norm_ret"], rtol=1e-6, atol=1e-7)


def test_running_stats_and_rolling_window():
    # reference tests/test_rollouts_extra.py:79-89
    rs = R.RunningStats()
    assert rs.mean() == 0.0 and rs.std() == 0.0
    rs.update(np.array([], np.float32))
    rs.update(np.array([1.0, 2.0, 3.0], np.float32))
    rs.update(np.array([4.0, -1.0, 0.0], np.float32))
    data = np.array([1.0, 2.0, 3.0, 4.0, -1.0, 0.0], np.float32)
    assert rs.count == 6
    np.testing.assert_allclose(rs.mean(), float(data.mean()), rtol=1e-6)
    np.testing.assert_allclose(rs.std(), float(data.std()), rtol=1e-6)
