"""GPU parity: gs_ppo_step / gs_reinforce_step / gs_clip_grad_norm / gs_adam_step vs the torch autograd oracle and the
fixtures produced by the reference's PPOAgent / REINFORCEAgent.  Bar: loss and gradients within 1e-4 (relative to the
gradient scale), metrics within 1e-4."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import policy as P

pytestmark = pytest.mark.gpu

TAGS = ["cartpole64", "acrobot128", "mcar256", "tiny64"]
PPO_METRICS = ["opt/loss/total", "opt/loss/policy", "opt/loss/entropy", "opt/policy/entropy", "opt/loss/entropy_scaled",
               "opt/loss/value", "opt/loss/value_scaled", "opt/ppo/clip_fraction", "opt/ppo/clip_fraction_vf",
               "opt/value/explained_var", "opt/ppo/kl", "opt/ppo/approx_kl"]


def _params(d, prefix):
    return {k: torch.from_numpy(d[prefix + k]) for k in P.PARAM_ORDER if prefix + k in d.files}


def _assert_grads_close(g, ref, tol=1e-4):
    scale = np.abs(ref).max()
    np.testing.assert_allclose(g, ref, rtol=tol, atol=tol * scale)


def _away_from_kinks(p, obs, activation, gen, margin=2e-5, scale=1.0):
    """Resample the observations whose hidden pre-activations land within `margin` of a ReLU kink.

    A gradient of these synthetic batches is a sum of ~n cancelling terms of size 1/n, so ONE sample whose ReLU unit falls
    on the other side of zero (|z| below the arithmetic noise of a summation order: ~1e-7 for fp32 FMA chains, ~3e-7 for the
    3xTF32 tensor-core path) moves an entry by ~5e-4 of the gradient scale — a discontinuity of the loss, not an arithmetic
    error.  Parity of the arithmetic is measured on batches without such samples (fp64 check, chunked)."""
    if activation != "relu":
        return obs
    flat = obs.reshape(-1, obs.shape[-1])
    w = {k: v.double() for k, v in p.items()}
    for _ in range(50):
        bad = torch.zeros(flat.shape[0], dtype=torch.bool)
        for lo in range(0, flat.shape[0], 1 << 18):
            x = flat[lo:lo + (1 << 18)].double()
            z1 = x @ w["w1"].T + w["b1"]
            b = z1.abs().lt(margin).any(-1)
            if "w2" in w:
                z2 = torch.relu(z1) @ w["w2"].T + w["b2"]
                b |= z2.abs().lt(margin).any(-1)
            bad[lo:lo + (1 << 18)] = b
        if not bad.any():
            return obs
        flat[bad] = torch.randn(int(bad.sum()), flat.shape[1], generator=gen) * scale
    raise AssertionError("could not move the batch away from the ReLU kinks")


def _noise_away_from(noise, edges, gen, std, margin=1e-4):
    """Resample noise entries within `margin` of a clipping edge (ratio = exp(-noise) vs 1 +- clip; |v - v_old| vs clip_vf)."""
    for _ in range(50):
        bad = torch.zeros_like(noise, dtype=torch.bool)
        for e in edges:
            bad |= (noise - e).abs() < margin
        if not bad.any():
            return noise
        noise[bad] = std * torch.randn(int(bad.sum()), generator=gen)
    raise AssertionError("could not move the noise away from the clipping edges")


def _ppo_hp(N, clip=0.2, clip_vf=0.2, vf=0.5, ent=0.01, norm=True, track=True):
    hp = N.GsPpoHparams()
    hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef = clip, clip_vf, vf, ent
    hp.normalize_adv, hp.track_activations = int(norm), int(track)
    return hp


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("norm", ["batch", "off"])
def test_ppo_step_matches_reference_fixture(golden_dir, tag, norm):
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    d = np.load(os.path.join(golden_dir, f"policy_{tag}.npz"))
    p = _params(d, "p_")
    B = d["obs"].shape[0]
    batch, keep = E.make_batch(1, B, E.cu(d["obs"][None]), E.cu(d["actions"][None].astype(np.int32)), E.cu(d["old_logp"][None]),
                               E.cu(d["values_old"][None]), E.cu(d["adv"][None]), E.cu(d["ret"][None]))
    g_raw, g_clip, m = E.update_step("ppo", E.dev_params(p), batch, _ppo_hp(N, norm=(norm == "batch")), max_norm=0.5)
    ref = np.concatenate([d[f"ppo_{norm}_g_{k}"].ravel() for k in P.PARAM_ORDER if f"ppo_{norm}_g_{k}" in d.files])
    _assert_grads_close(g_raw, ref)
    np.testing.assert_allclose(m["opt/loss/total"], float(d[f"ppo_{norm}_loss"]), rtol=1e-4, atol=1e-6)
    for k in PPO_METRICS:
        np.testing.assert_allclose(m[k], float(d[f"ppo_{norm}_m_{k}"]), rtol=1e-4, atol=2e-6, err_msg=k)
    if norm == "batch":
        np.testing.assert_allclose(m["roll/adv/norm/mean"], float(d[f"ppo_{norm}_m_roll/adv/norm/mean"]), atol=1e-6)
        np.testing.assert_allclose(m["roll/adv/norm/std"], float(d[f"ppo_{norm}_m_roll/adv/norm/std"]), rtol=1e-4)
    for layer in ("backbone.0", "backbone.2"):
        if f"ppo_{norm}_m_opt/activations/{layer}/mean" not in d.files:
            continue
        for stat, tol in (("mean", 1e-4), ("std", 1e-4), ("dead_pct", 1e-6), ("dead_max", 1e-6)):
            np.testing.assert_allclose(m[f"opt/activations/{layer}/{stat}"], float(d[f"ppo_{norm}_m_opt/activations/{layer}/{stat}"]),
                                       rtol=tol, atol=1e-6, err_msg=f"{layer}/{stat}")
    for grp in ("all", "backbone", "policy_head", "value_head"):
        np.testing.assert_allclose(m[f"opt/grads/norm/{grp}"], float(d[f"ppo_{norm}_m_opt/grads/norm/{grp}"]), rtol=1e-4)
    ref_clip = np.concatenate([d[f"ppo_{norm}_gc_{k}"].ravel() for k in P.PARAM_ORDER if f"ppo_{norm}_gc_{k}" in d.files])
    _assert_grads_close(g_clip, ref_clip)


@pytest.mark.parametrize("D,hidden,A", [(4, (64, 64), 2), (6, (128, 128), 3), (2, (256, 256), 3), (4, (64,), 2)])
@pytest.mark.parametrize("mode", ["idx", "identity", "idx_map"])
@pytest.mark.parametrize("activation", ["relu", "tanh"])
def test_ppo_step_gather_modes_vs_oracle(D, hidden, A, mode, activation):
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    T, Nn = 9, 61
    g = torch.Generator().manual_seed(T * Nn + D)
    p = P.random_params(D, hidden, A, seed=D)
    obs = torch.randn(T, Nn, D, generator=g)
    actions = torch.randint(0, A, (T, Nn), generator=g)
    with torch.no_grad():
        logits, v = P.forward(p, obs.reshape(-1, D), activation)
        lp_all = logits - logits.logsumexp(-1, keepdim=True)
    old_logp = (lp_all.gather(-1, actions.reshape(-1, 1)).squeeze(-1) + 0.2 * torch.randn(T * Nn, generator=g)).reshape(T, Nn)
    values_old = (v + 0.3 * torch.randn(T * Nn, generator=g)).reshape(T, Nn)
    adv = torch.randn(T, Nn, generator=g) * 2 + 0.3
    ret = values_old + adv
    total = T * Nn
    idx_map = None
    if mode == "idx":
        ids = torch.randint(0, total, (300,), generator=g).numpy()
        src = ids
        batch, keep = E.make_batch(T, Nn, E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret), idx=ids)
    elif mode == "identity":
        src = np.arange(100, 400)
        batch, keep = E.make_batch(T, Nn, E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret),
                                   n=300, perm_offset=100)
    else:
        ids = torch.randint(0, total, (257,), generator=g).numpy()
        idx_map = torch.randint(0, total, (total,), generator=g).numpy()
        src = idx_map[ids]
        batch, keep = E.make_batch(T, Nn, E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret),
                                   idx=ids, idx_map=idx_map)
    e, t = src // T, src % T   # env-major sample id -> (env, step)
    sel = lambda x: x[t, e]
    hp = _ppo_hp(N, clip=0.15, clip_vf=0.25, vf=0.7, ent=0.02)
    g_raw, _, m = E.update_step("ppo", E.dev_params(p), batch, hp, activation=activation)
    loss, flat, om = P.loss_and_grads(P.ppo_loss, p, sel(obs), sel(actions), sel(old_logp), sel(values_old), sel(adv), sel(ret),
                                      clip_range=0.15, clip_range_vf=0.25, vf_coef=0.7, ent_coef=0.02, normalize_adv=True, activation=activation)
    _assert_grads_close(g_raw, flat.numpy())
    np.testing.assert_allclose(m["opt/loss/total"], float(loss), rtol=1e-4, atol=1e-6)
    for k in PPO_METRICS:
        np.testing.assert_allclose(m[k], float(om[k]), rtol=1e-4, atol=2e-6, err_msg=k)
    assert m["opt/batch_count"] == len(src)


def test_ppo_step_full_size_unfiltered_minibatch_is_closer_to_fp64_than_the_fp32_reference():
    """The SAME 1M-sample minibatch WITHOUT moving samples off the kinks of the loss (ReLU zeros, ratio clip edges, clipped-value-loss
    ties).  There the reference's own fp32 arithmetic is 6e-4 away from an fp64 evaluation of the same formulas (a handful of samples
    falls on the other side of a kink), so agreement with torch fp32 to 1e-4 is not defined; what is: the engine (fp16x3 tensor-core
    products, fp32 accumulation) must be at least as close to the fp64 evaluation as the fp32 reference is.  Reported, not just asserted."""
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    T, Nn, D, A = 128, 65536, 4, 2
    g = torch.Generator().manual_seed(0)
    p = P.random_params(D, (64, 64), A, seed=1)
    obs = torch.randn(T, Nn, D, generator=g) * 0.5
    actions = torch.randint(0, A, (T, Nn), generator=g)
    with torch.no_grad():
        logits, v = P.forward(p, obs.reshape(-1, D))
        lp_all = logits - logits.logsumexp(-1, keepdim=True)
    old_logp = (lp_all.gather(-1, actions.reshape(-1, 1)).squeeze(-1) + 0.1 * torch.randn(T * Nn, generator=g)).reshape(T, Nn)
    values_old = (v + 0.3 * torch.randn(T * Nn, generator=g)).reshape(T, Nn)
    adv = torch.randn(T, Nn, generator=g)
    ret = values_old + adv
    B = 1 << 20
    dev = [E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret)]
    src = np.arange(3 * B, 4 * B)
    batch, keep = E.make_batch(T, Nn, *dev, n=B, perm_offset=3 * B)
    g_raw, _, m = E.update_step("ppo", E.dev_params(p), batch, _ppo_hp(N, track=False))
    e, t = src // T, src % T
    sel = lambda x: x[t, e]
    kw = dict(clip_range=0.2, clip_range_vf=0.2, vf_coef=0.5, ent_coef=0.01, normalize_adv=True)
    _, flat64, _ = P.loss_and_grads(P.ppo_loss, {k: v_.double() for k, v_ in p.items()}, sel(obs).double(), sel(actions), sel(old_logp).double(),
                                    sel(values_old).double(), sel(adv).double(), sel(ret).double(), **kw)
    _, flat32, _ = P.loss_and_grads(P.ppo_loss, p, sel(obs), sel(actions), sel(old_logp), sel(values_old), sel(adv), sel(ret), **kw)
    ref64, ref32 = flat64.numpy(), flat32.numpy()
    rel = lambda a: float(np.linalg.norm(a - ref64) / np.linalg.norm(ref64))
    e_engine, e_torch = rel(g_raw), rel(ref32)
    print(f"unfiltered 1M-sample minibatch, L2 error vs fp64: engine {e_engine:.3e}, torch fp32 reference {e_torch:.3e}; "
          f"engine max |err| / max |g| {np.abs(g_raw - ref64).max() / np.abs(ref64).max():.3e}")
    assert e_engine <= max(e_torch, 1e-4) * 1.05
    assert np.abs(g_raw - ref64).max() <= 5e-4 * np.abs(ref64).max()


def test_ppo_step_full_size_minibatch_vs_oracle():
    """A C2-sized minibatch (1,048,576 samples gathered from a 128 x 65,536 rollout by the device permutation)."""
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    T, Nn, D, A = 128, 65536, 4, 2
    g = torch.Generator().manual_seed(0)
    p = P.random_params(D, (64, 64), A, seed=1)
    import math
    obs = _away_from_kinks(p, torch.randn(T, Nn, D, generator=g) * 0.5, "relu", g, scale=0.5)
    actions = torch.randint(0, A, (T, Nn), generator=g)
    with torch.no_grad():
        logits, v = P.forward(p, obs.reshape(-1, D))
        lp_all = logits - logits.logsumexp(-1, keepdim=True)
    n_lp = _noise_away_from(0.1 * torch.randn(T * Nn, generator=g), [-math.log(1.2), -math.log(0.8)], g, 0.1)
    n_v = _noise_away_from(0.3 * torch.randn(T * Nn, generator=g), [-0.2, 0.2], g, 0.3)
    old_logp = (lp_all.gather(-1, actions.reshape(-1, 1)).squeeze(-1) + n_lp).reshape(T, Nn)
    values_old = (v + n_v).reshape(T, Nn)
    adv = torch.randn(T, Nn, generator=g)
    ret = values_old + adv
    total, B = T * Nn, 1 << 20
    dev = [E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret)]
    # the device permutation is a bijection of [0, total): minibatch k of an epoch covers disjoint ids
    ids = []
    import gymnasium_solver_b200.utils.samplers as S
    perm = S.feistel_permutation(total, key=77, device="cuda")
    assert torch.equal(torch.sort(perm).values, torch.arange(total, device="cuda"))
    src = perm[3 * B:4 * B].cpu().numpy()
    batch, keep = E.make_batch(T, Nn, *dev, n=B, perm_key=77, perm_offset=3 * B, perm_len=total)
    hp = _ppo_hp(N, track=False)
    g_raw, _, m = E.update_step("ppo", E.dev_params(p), batch, hp)
    e, t = src // T, src % T
    sel = lambda x: x[t, e]
    # oracle = the reference's own arithmetic (torch fp32 autograd).  An fp64 evaluation differs from it by ~6e-4 on this data
    # (max(lu, lc) tie-breaking of the clipped value loss flips with the rounding of v_old + (v - v_old)), so fp64 is not the bar.
    loss, flat, om = P.loss_and_grads(P.ppo_loss, p, sel(obs), sel(actions), sel(old_logp), sel(values_old), sel(adv), sel(ret),
                                      clip_range=0.2, clip_range_vf=0.2, vf_coef=0.5, ent_coef=0.01, normalize_adv=True)
    ref = flat.numpy()
    # 1M-sample fp32 sums in a different order than torch's (the batch holds no sample on a ReLU / clipping kink): north_star's 1e-4,
    # elementwise (of the gradient scale) and in L2
    _assert_grads_close(g_raw, ref, tol=1e-4)
    assert np.linalg.norm(g_raw - ref) <= 1e-4 * np.linalg.norm(ref)
    np.testing.assert_allclose(m["opt/loss/total"], float(loss), rtol=1e-4)
    for k in PPO_METRICS:
        np.testing.assert_allclose(m[k], float(om[k]), rtol=2e-4, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("cfg", [("returns", "off", "off"), ("advantages", "off", "batch"), ("returns", "batch", "off")])
def test_reinforce_step_matches_reference_fixture(golden_dir, tag, cfg):
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    targets, nr, na = cfg
    d = np.load(os.path.join(golden_dir, f"policy_{tag}.npz"))
    p = _params(d, "rp_")
    B = d["obs"].shape[0]
    batch, keep = E.make_batch(1, B, E.cu(d["obs"][None]), E.cu(d["actions"][None].astype(np.int32)), E.cu(d["r_old_logp"][None]),
                               E.cu(d["values_old"][None]), E.cu(d["adv"][None]), E.cu(d["ret"][None]))
    hp = N.GsReinforceHparams()
    hp.ent_coef, hp.policy_targets = 0.01, 0 if targets == "returns" else 1
    hp.normalize_returns, hp.normalize_adv, hp.track_activations = int(nr == "batch"), int(na == "batch"), 0
    g_raw, _, m = E.update_step("reinforce", E.dev_params(p), batch, hp)
    key = f"rf_{targets}_{nr}_{na}"
    ref = np.concatenate([d[f"{key}_g_{k}"].ravel() for k in P.PARAM_ORDER if f"{key}_g_{k}" in d.files])
    _assert_grads_close(g_raw, ref)
    for k in ("opt/loss/total", "opt/loss/policy", "opt/policy/entropy", "opt/ppo/kl", "opt/ppo/approx_kl", "policy_targets_mean", "policy_targets_std"):
        np.testing.assert_allclose(m[k], float(d[f"{key}_m_{k}"]), rtol=1e-4, atol=2e-6, err_msg=k)


def test_ppo_clip_math_reference_kat():
    """reference tests/test_ppo.py:52-107 — one ratio above 1+clip with adv > 0, one below 1-clip with adv < 0: the
    clipped branch is selected for both, so the policy gradient vanishes and loss == -(1.5*1.2 - 2.0*0.8)/2."""
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    p = P.random_params(4, (64, 64), 2, seed=5)
    obs = torch.randn(2, 4, generator=torch.Generator().manual_seed(1))
    actions = torch.tensor([0, 1])
    with torch.no_grad():
        logits, v = P.forward(p, obs)
        lp = (logits - logits.logsumexp(-1, keepdim=True)).gather(-1, actions[:, None]).squeeze(-1)
    ratios = torch.tensor([1.5, 0.55])
    old_logp = lp - torch.log(ratios)
    adv = torch.tensor([1.5, -2.0])
    batch, keep = E.make_batch(1, 2, E.cu(obs[None]), E.cu(actions[None].int()), E.cu(old_logp[None]), E.cu(v[None]), E.cu(adv[None]), E.cu(v[None]))
    hp = _ppo_hp(N, clip=0.2, clip_vf=1e6, vf=0.0, ent=0.0, norm=False, track=False)
    g_raw, _, m = E.update_step("ppo", E.dev_params(p), batch, hp)
    np.testing.assert_allclose(m["opt/loss/total"], -(1.5 * 1.2 + -2.0 * 0.8) / 2, rtol=1e-5)
    assert m["opt/ppo/clip_fraction"] == 1.0
    np.testing.assert_allclose(g_raw, 0.0, atol=1e-7)


def test_adam_step_matches_torch():
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    n = 4675
    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = E.cu(p0.clone())
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for it in range(5):
        grad = torch.randn(n, generator=g) * (0.1 + it)
        ref.grad = grad.clone()
        opt.step()
        gd = E.cu(grad)
        N.check(N.lib().gs_adam_step(N.ptr(p), N.ptr(gd), N.ptr(m), N.ptr(v), n, N.ptr(step), 1e-3, 0.9, 0.999, 1e-8, N.stream()))
    E.sync()
    assert int(step.item()) == 5
    np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_update_errors_fail_loudly():
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    p = P.random_params(4, (64, 64), 2, has_value=False)
    z = torch.zeros(1, 4)
    batch, keep = E.make_batch(1, 4, E.cu(torch.zeros(1, 4, 4)), E.cu(z.int()), E.cu(z), E.cu(z), E.cu(z), E.cu(z))
    with pytest.raises(N.EngineError, match="value head"):
        E.update_step("ppo", E.dev_params(p), batch, _ppo_hp(N))


@pytest.mark.parametrize("algo", ["ppo", "reinforce"])
@pytest.mark.parametrize("n,D,A,activation", [(96, 4, 2, "relu"), (128, 4, 2, "relu"), (129, 6, 3, "relu"), (1000, 2, 3, "tanh"), (40000, 4, 2, "relu")])
def test_tensor_core_and_simt_update_kernels_agree(algo, n, D, A, activation):
    """64x64 network: the tcgen05 (fp16x3, TMEM accumulators) kernel and the fp32 FMA-pipe kernel implement the same
    contract; both must match the fp32 torch oracle, and each other far inside the 1e-4 bar."""
    _tensor_core_vs_simt(algo, n, D, A, activation, 64)


@pytest.mark.parametrize("hidden,n,D,A,algo", [(128, 40000, 6, 3, "ppo"), (256, 129, 2, 3, "ppo"), (256, 40000, 2, 3, "ppo"), (256, 70000, 4, 2, "reinforce")])
def test_wide_tensor_core_update_kernels_agree_with_simt_and_oracle(hidden, n, D, A, algo):
    """128x128 (one tile set of update_f16_kernel) and 256x256 (update_wide_kernel + wgrad_wide_kernel: streamed W2, stored operand
    tiles) on minibatches of several tiles per CTA: same bar against the fp32 torch oracle and the FMA-pipe kernel."""
    _tensor_core_vs_simt(algo, n, D, A, "relu", hidden)


def _tensor_core_vs_simt(algo, n, D, A, activation, hidden):
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    g = torch.Generator().manual_seed(n + D)
    p = P.random_params(D, (hidden, hidden), A, seed=n, has_value=True)
    import math
    obs = _away_from_kinks(p, torch.randn(1, n, D, generator=g), activation, g)
    actions = torch.randint(0, A, (1, n), generator=g)
    with torch.no_grad():
        logits, v = P.forward(p, obs.reshape(-1, D), activation)
        lp = (logits - logits.logsumexp(-1, keepdim=True)).gather(-1, actions.reshape(-1, 1)).squeeze(-1)
    old_logp = (lp + _noise_away_from(0.2 * torch.randn(n, generator=g), [-math.log(1.15), -math.log(0.85)], g, 0.2)).reshape(1, n)
    values_old = (v + _noise_away_from(0.3 * torch.randn(n, generator=g), [-0.25, 0.25], g, 0.3)).reshape(1, n)
    adv = torch.randn(1, n, generator=g) * 2 + 0.3
    ret = values_old + adv
    batch, keep = E.make_batch(1, n, E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret))
    if algo == "ppo":
        hp = _ppo_hp(N, clip=0.15, clip_vf=0.25, vf=0.7, ent=0.02)
        loss, flat, om = P.loss_and_grads(P.ppo_loss, p, obs[0], actions[0], old_logp[0], values_old[0], adv[0], ret[0], clip_range=0.15,
                                          clip_range_vf=0.25, vf_coef=0.7, ent_coef=0.02, normalize_adv=True, activation=activation)
    else:
        hp = N.GsReinforceHparams()
        hp.ent_coef, hp.policy_targets, hp.normalize_returns, hp.normalize_adv, hp.track_activations = 0.02, 1, 0, 1, 1
        loss, flat, om = P.loss_and_grads(P.reinforce_loss, p, obs[0], actions[0], old_logp[0], adv[0], ret[0], ent_coef=0.02,
                                          policy_targets="advantages", normalize_adv=True, activation=activation)
    out = {}
    try:
        for impl in (0, 1):
            N.check(N.lib().gs_set_update_impl(impl))
            out[impl] = E.update_step(algo, E.dev_params(p), batch, hp, activation=activation, max_norm=0.5)
    finally:
        N.lib().gs_set_update_impl(0)
    ref = flat.numpy()
    scale = np.abs(ref).max()
    for impl in (0, 1):
        g_raw, g_clip, m = out[impl]
        np.testing.assert_allclose(g_raw, ref, rtol=1e-4, atol=1e-4 * scale, err_msg=f"impl {impl}")
        assert np.linalg.norm(g_raw - ref) <= 1e-4 * np.linalg.norm(ref), f"impl {impl}"
        np.testing.assert_allclose(m["opt/loss/total"], float(loss), rtol=1e-4, atol=1e-6)
        acts = om["_activations"]
        for name in acts:
            for stat in ("mean", "std"):
                np.testing.assert_allclose(m[f"opt/activations/{name}/{stat}"], acts[name][stat], rtol=1e-4, atol=1e-6, err_msg=f"impl {impl} {name}/{stat}")
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=2e-5, atol=2e-5 * scale)
    assert np.linalg.norm(out[0][0] - out[1][0]) <= 1e-4 * np.linalg.norm(out[1][0])
    for k in ("opt/loss/total", "opt/ppo/kl", "opt/ppo/approx_kl", "opt/policy/entropy", "opt/grads/norm/all"):
        np.testing.assert_allclose(out[0][2][k], out[1][2][k], rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("impl", [0, 1])
def test_dead_units_are_counted_and_do_not_serialise_the_kernel(impl):
    """A hidden unit that is dead for every sample (z == 0: |z| < 1e-6, utils/models.py:121-146) must be counted exactly, and
    counting it must not cost one global atomic per sample on a single address (a 1M-sample minibatch took 42 ms that way)."""
    import time
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    n, D, A = 1 << 18, 4, 2
    g = torch.Generator().manual_seed(5)
    p = P.random_params(D, (64, 64), A, seed=9, has_value=True)
    p["w1"][3].zero_(); p["b1"][3] = 0.0                      # layer-1 unit 3: z = 0 for every sample
    p["w2"][10].zero_(); p["b2"][10] = 0.0                    # layer-2 unit 10 as well
    obs = torch.randn(1, n, D, generator=g)
    actions = torch.randint(0, A, (1, n), generator=g)
    z = torch.randn(1, n, generator=g)
    batch, keep = E.make_batch(1, n, E.cu(obs), E.cu(actions.int()), E.cu(z * 0.1 - 0.7), E.cu(z), E.cu(z + 0.3), E.cu(2 * z))
    hp = _ppo_hp(N)
    try:
        N.check(N.lib().gs_set_update_impl(impl))
        E.update_step("ppo", E.dev_params(p), batch, hp)      # warm-up (module load, first launch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, m = E.update_step("ppo", E.dev_params(p), batch, hp)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        N.lib().gs_set_update_impl(0)
    np.testing.assert_allclose(m["opt/activations/backbone.0/dead_max"], 1.0, atol=1e-9)
    np.testing.assert_allclose(m["opt/activations/backbone.2/dead_max"], 1.0, atol=1e-9)
    np.testing.assert_allclose(m["opt/activations/backbone.0/dead_pct"], 1.0 / 64, atol=2e-4)
    np.testing.assert_allclose(m["opt/activations/backbone.2/dead_pct"], 1.0 / 64, atol=2e-4)
    assert dt < 0.05, f"update step took {dt * 1e3:.1f} ms with two dead units"


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("algo", ["ppo", "reinforce"])
def test_step_takes_minibatch_moments_itself_when_none_are_given(algo, impl):
    """NULL moments + a "batch" normalisation flag = moments of this minibatch, taken inside the call: same gradients and
    metrics as gs_batch_moments followed by the step."""
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    n, D, A = 5000, 4, 2
    g = torch.Generator().manual_seed(11)
    p = P.random_params(D, (64, 64), A, seed=3, has_value=True)
    obs = torch.randn(1, n, D, generator=g)
    actions = torch.randint(0, A, (1, n), generator=g)
    z = torch.randn(1, n, generator=g)
    batch, keep = E.make_batch(1, n, E.cu(obs), E.cu(actions.int()), E.cu(z * 0.1 - 0.7), E.cu(z), E.cu(1.5 * z + 0.3), E.cu(2 * z - 1))
    if algo == "ppo":
        hp = _ppo_hp(N)
    else:
        hp = N.GsReinforceHparams()
        hp.ent_coef, hp.policy_targets, hp.normalize_returns, hp.normalize_adv, hp.track_activations = 0.01, 0, 1, 1, 1
    try:
        N.check(N.lib().gs_set_update_impl(impl))
        g_a, _, m_a = E.update_step(algo, E.dev_params(p), batch, hp)
        g_b, _, m_b = E.update_step(algo, E.dev_params(p), batch, hp, internal_moments=True)
    finally:
        N.lib().gs_set_update_impl(0)
    np.testing.assert_allclose(g_b, g_a, rtol=1e-6, atol=1e-9)
    for k in m_a:
        np.testing.assert_allclose(m_b[k], m_a[k], rtol=1e-6, atol=1e-9, err_msg=k)


@pytest.mark.parametrize("D,A", [(4, 2), (6, 3), (2, 3)])
def test_packed_sample_records_layout_and_identical_update(D, A):
    """gs_rollout_pack writes one 64-byte record per (t, n): the fp16x3 layer-1 operand row x16 = [x_hi (cols 0..D-1), x_lo (cols 7..7+D-1),
    1, 1 (cols 14, 15)] in fp16 followed by {action bits, logp_old, values_old, adv, ret, 0, 0, 0}; the update step that gathers rollout records
    returns bit-identical gradients and metrics to the one that builds the minibatch's records from the seven arrays."""
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    T, Nn = 9, 700
    g = torch.Generator().manual_seed(21 + D)
    p = P.random_params(D, (64, 64), A, seed=4, has_value=True)
    obs = torch.randn(T, Nn, D, generator=g)
    actions = torch.randint(0, A, (T, Nn), generator=g)
    z = torch.randn(T, Nn, generator=g)
    arrs = [obs, actions.int(), z * 0.1 - 0.7, z, 1.5 * z + 0.3, 2 * z - 1]
    total = T * Nn
    results = []
    for use_packed in (False, True):
        batch, keep = E.make_batch(T, Nn, *[E.cu(a) for a in arrs], n=4096, perm_key=5, perm_offset=1000, perm_len=total)
        if use_packed:
            rec = E.pack_rollout(batch, keep).cpu().numpy()
            x = obs.reshape(total, D)
            x_hi = x.to(torch.float16)
            x_lo = (x - x_hi.float()).to(torch.float16)
            half = rec[:, :8].copy().view(np.uint16).reshape(total, 16)
            np.testing.assert_array_equal(half[:, :D], x_hi.view(torch.int16).numpy().view(np.uint16))
            np.testing.assert_array_equal(half[:, 7:7 + D], x_lo.view(torch.int16).numpy().view(np.uint16))
            assert (half[:, D:7] == 0).all() and (half[:, 7 + D:14] == 0).all() and (half[:, 14:] == 0x3C00).all()
            assert (rec[:, 13:] == 0).all()
            np.testing.assert_array_equal(rec[:, 8].view(np.int32), actions.reshape(total).numpy().astype(np.int32))
            for col, a in zip((9, 10, 11, 12), arrs[2:]):
                np.testing.assert_array_equal(rec[:, col], a.reshape(total).numpy().astype(np.float32))
        results.append(E.update_step("ppo", E.dev_params(p), batch, _ppo_hp(N), internal_moments=True))
    (g0, _, m0), (g1, _, m1) = results
    np.testing.assert_array_equal(g1, g0)
    for k in m0:
        np.testing.assert_array_equal(m1[k], m0[k], err_msg=k)


@pytest.mark.parametrize("algo", ["ppo", "reinforce"])
def test_two_phase_step_for_sharded_minibatches_matches_the_one_call_step(algo):
    """gs_batch_prepare (gather pass + local moments) -> [all-reduce by the caller] -> step with batch.prepared = 1 returns the
    same bits as the single call that takes the moments itself."""
    import ctypes as C
    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    T, Nn, D, A = 16, 512, 4, 2
    g = torch.Generator().manual_seed(33)
    p = P.random_params(D, (64, 64), A, seed=6, has_value=True)
    obs = torch.randn(T, Nn, D, generator=g)
    actions = torch.randint(0, A, (T, Nn), generator=g)
    z = torch.randn(T, Nn, generator=g)
    arrs = [E.cu(a) for a in (obs, actions.int(), z * 0.1 - 0.7, z, 1.5 * z + 0.3, 2 * z - 1)]
    batch, keep = E.make_batch(T, Nn, *arrs, n=3000, perm_key=9, perm_offset=2048, perm_len=T * Nn)
    if algo == "ppo":
        hp = _ppo_hp(N)
    else:
        hp = N.GsReinforceHparams()
        hp.ent_coef, hp.policy_targets, hp.normalize_returns, hp.normalize_adv, hp.track_activations = 0.01, 0, 1, 1, 1
    g_ref, _, m_ref = E.update_step(algo, E.dev_params(p), batch, hp, internal_moments=True)
    # two-phase
    pd = E.dev_params(p)
    m = N.mlp_struct_from_params(pd, "relu")
    L = N.lib()
    Pn = L.gs_mlp_param_count(C.byref(m))
    wsb = L.gs_update_workspace_bytes(C.byref(m), 0, int(batch.n))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    grads = torch.empty(Pn, device="cuda")
    met = torch.zeros(N.N_METRICS, dtype=torch.float64, device="cuda")
    mom = torch.full((6,), float("nan"), dtype=torch.float64, device="cuda")
    N.check(L.gs_batch_prepare(C.byref(m), C.byref(batch), 1, int(algo == "reinforce"), N.ptr(mom), N.ptr(ws), wsb, N.stream()))
    batch.prepared = 1
    if algo == "ppo":
        N.check(L.gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(mom), N.ptr(grads), N.ptr(met), N.ptr(ws), wsb, N.stream()))
    else:
        N.check(L.gs_reinforce_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(mom[3:6]), N.ptr(mom[0:3]), N.ptr(grads), N.ptr(met),
                                    N.ptr(ws), wsb, N.stream()))
    torch.cuda.synchronize()
    assert float(mom[2]) == batch.n
    np.testing.assert_array_equal(grads.cpu().numpy(), g_ref)
    mv = met.cpu().numpy()
    for i, k in enumerate(N.METRIC_KEYS):
        if not k.startswith("opt/grads"):
            np.testing.assert_array_equal(mv[i], m_ref[k], err_msg=k)


@pytest.mark.parametrize("tag", TAGS)
def test_deferred_step_and_update_finish_match_reference_fixture(golden_dir, tag):
    """C-ABI: gs_ppo_step with defer_reduce + gs_update_finish (ordered reduction, metrics, norms, clip, Adam in one launch) against
    the reference-generated fixture (gradients before / after clip_grad_norm_, metrics) and torch.optim.Adam on the clipped gradient."""
    import ctypes as C

    import engine_api as E
    from gymnasium_solver_b200 import _native as N

    d = np.load(os.path.join(golden_dir, f"policy_{tag}.npz"))
    p = _params(d, "p_")
    B = d["obs"].shape[0]
    batch, keep = E.make_batch(1, B, E.cu(d["obs"][None]), E.cu(d["actions"][None].astype(np.int32)), E.cu(d["old_logp"][None]),
                               E.cu(d["values_old"][None]), E.cu(d["adv"][None]), E.cu(d["ret"][None]))
    # flat parameter vector in nn.Module.parameters() order; the gs_mlp_t points into it (what flatten_parameters_ does for a model)
    order = [k for k in P.PARAM_ORDER if k in p]
    flat = torch.cat([p[k].reshape(-1) for k in order]).to("cuda").contiguous()
    views, off = {}, 0
    for k in order:
        n = p[k].numel()
        views[k] = flat[off:off + n].view(p[k].shape)
        off += n
    m = N.mlp_struct_from_params(views)
    Pn = int(N.lib().gs_mlp_param_count(C.byref(m)))
    assert Pn == flat.numel()
    wsb = N.lib().gs_update_workspace_bytes(C.byref(m), 0, B)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    grads = torch.full((Pn,), float("nan"), device="cuda")
    metrics = torch.zeros(N.N_METRICS, dtype=torch.float64, device="cuda")
    msum = torch.zeros(N.N_METRICS, dtype=torch.float64, device="cuda")
    hp = _ppo_hp(N, norm=True)
    batch.defer_reduce = 1
    L = N.lib()
    N.check(L.gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), None, N.ptr(grads), N.ptr(metrics), N.ptr(ws), wsb, N.stream()))
    fin = N.GsFinish()
    fin.algo, fin.track_activations, fin.normalize_adv, fin.normalize_ret = 0, hp.track_activations, 1, 0
    fin.vf_coef, fin.ent_coef, fin.max_grad_norm = hp.vf_coef, hp.ent_coef, 0.5
    # 1) no optimizer: grads_flat = clipped gradient, metrics complete, metrics_sum accumulates
    w0 = flat.clone()
    N.check(L.gs_update_finish(C.byref(m), C.byref(batch), C.byref(fin), N.ptr(grads), None, None, N.ptr(metrics), N.ptr(msum), N.ptr(ws), wsb, N.stream()))
    E.sync()
    assert torch.equal(flat, w0)
    ref_clip = np.concatenate([d[f"ppo_batch_gc_{k}"].ravel() for k in P.PARAM_ORDER if f"ppo_batch_gc_{k}" in d.files])
    _assert_grads_close(grads.cpu().numpy(), ref_clip)
    mv = metrics.cpu().numpy()
    mm = {k: mv[i] for i, k in enumerate(N.METRIC_KEYS)}
    np.testing.assert_allclose(mm["opt/loss/total"], float(d["ppo_batch_loss"]), rtol=1e-4, atol=1e-6)
    for k in PPO_METRICS:
        np.testing.assert_allclose(mm[k], float(d[f"ppo_batch_m_{k}"]), rtol=1e-4, atol=2e-6, err_msg=k)
    for grp in ("all", "backbone", "policy_head", "value_head"):
        np.testing.assert_allclose(mm[f"opt/grads/norm/{grp}"], float(d[f"ppo_batch_m_opt/grads/norm/{grp}"]), rtol=1e-4)
    np.testing.assert_array_equal(msum.cpu().numpy()[:32], mv[:32])
    # 2) with Adam: the same call sequence again, parameters move like torch.optim.Adam on the clipped gradient
    ref_p = w0.detach().cpu().clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    ref_p.grad = grads.detach().cpu().clone()
    opt.step()
    exp_avg, exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    adam = N.GsAdam()
    adam.params_flat, adam.exp_avg, adam.exp_avg_sq, adam.step_count = N.ptr(flat), N.ptr(exp_avg), N.ptr(exp_avg_sq), N.ptr(step)
    adam.lr, adam.beta1, adam.beta2, adam.eps = 1e-3, 0.9, 0.999, 1e-8
    N.check(L.gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), None, N.ptr(grads), N.ptr(metrics), N.ptr(ws), wsb, N.stream()))
    N.check(L.gs_update_finish(C.byref(m), C.byref(batch), C.byref(fin), N.ptr(grads), C.byref(adam), None, N.ptr(metrics), N.ptr(msum), N.ptr(ws), wsb, N.stream()))
    E.sync()
    assert int(step.item()) == 1
    np.testing.assert_allclose(flat.cpu().numpy(), ref_p.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(msum.cpu().numpy()[0], 2 * mv[0], rtol=1e-12)
    # 3) errors fail loudly
    with pytest.raises(N.EngineError, match="NULL"):
        N.check(L.gs_update_finish(C.byref(m), C.byref(batch), C.byref(fin), None, None, None, N.ptr(metrics), None, N.ptr(ws), wsb, N.stream()))
