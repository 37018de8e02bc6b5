"""dev helper: TC vs SIMT vs fp64 on the 40000-sample REINFORCE case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import engine_api as E
from gymnasium_solver_b200 import _native as N
from oracle import policy as P
n, D, A = 40000, 4, 2
g = torch.Generator().manual_seed(n + D)
p = P.random_params(D, (64, 64), A, seed=n, has_value=True)
obs = torch.randn(1, n, D, generator=g)
actions = torch.randint(0, A, (1, n), generator=g)
with torch.no_grad():
    logits, v = P.forward(p, obs.reshape(-1, D))
    lp = (logits - logits.logsumexp(-1, keepdim=True)).gather(-1, actions.reshape(-1, 1)).squeeze(-1)
old_logp = (lp + 0.2 * torch.randn(n, generator=g)).reshape(1, n)
values_old = (v + 0.3 * torch.randn(n, generator=g)).reshape(1, n)
adv = torch.randn(1, n, generator=g) * 2 + 0.3
ret = values_old + adv
batch, keep = E.make_batch(1, n, E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret))
hp = N.GsReinforceHparams(); hp.ent_coef, hp.policy_targets, hp.normalize_returns, hp.normalize_adv, hp.track_activations = 0.02, 1, 0, 1, 1
kw = dict(ent_coef=0.02, policy_targets="advantages", normalize_adv=True)
_, f32, _ = P.loss_and_grads(P.reinforce_loss, p, obs[0], actions[0], old_logp[0], adv[0], ret[0], **kw)
_, f64, _ = P.loss_and_grads(P.reinforce_loss, {k: t.double() for k, t in p.items()}, obs[0].double(), actions[0], old_logp[0].double(), adv[0].double(), ret[0].double(), **kw)
names = []
for k in P.PARAM_ORDER:
    names += [k] * p[k].numel()
names = np.array(names)
r64 = f64.numpy(); scale = np.abs(r64).max()
print("scale", scale, " torch fp32 vs fp64 max abs", np.abs(f32.numpy() - r64).max())
for impl in (0, 1):
    N.lib().gs_set_update_impl(impl)
    g_raw, _, m = E.update_step("reinforce", E.dev_params(p), batch, hp)
    err = np.abs(g_raw - r64)
    print(f"impl {impl}: max abs err vs fp64 {err.max():.3e} ({err.max()/scale:.2e} of scale)")
    for k in P.PARAM_ORDER:
        s_ = names == k
        print(f"   {k}: max abs err {err[s_].max():.3e}  max|ref| {np.abs(r64[s_]).max():.3e}")
N.lib().gs_set_update_impl(0)
