"""dev helper: where do the tensor-core kernel's gradients deviate most (full-size minibatch)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import engine_api as E
from gymnasium_solver_b200 import _native as N
from oracle import policy as P

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
total, B = T * Nn, 1 << 20
dev = [E.cu(obs), E.cu(actions.int()), E.cu(old_logp), E.cu(values_old), E.cu(adv), E.cu(ret)]
src = np.arange(3 * B, 4 * B)
batch, keep = E.make_batch(T, Nn, *dev, n=B, perm_offset=3 * B)
hp = N.GsPpoHparams(); hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef, hp.normalize_adv, hp.track_activations = 0.2, 0.2, 0.5, 0.01, 1, 0
e, t = src // T, src % T
sel = lambda x: x[t, e]
loss, flat, om = P.loss_and_grads(P.ppo_loss, {k: v_.double() for k, v_ in p.items()}, sel(obs).double(), sel(actions), sel(old_logp).double(),
                                  sel(values_old).double(), sel(adv).double(), sel(ret).double(), clip_range=0.2, clip_range_vf=0.2,
                                  vf_coef=0.5, ent_coef=0.01, normalize_adv=True)
ref = flat.numpy()
loss32, flat32, _ = P.loss_and_grads(P.ppo_loss, p, sel(obs), sel(actions), sel(old_logp), sel(values_old), sel(adv), sel(ret), clip_range=0.2,
                                     clip_range_vf=0.2, vf_coef=0.5, ent_coef=0.01, normalize_adv=True)
ref32 = flat32.numpy()
print(f"torch fp32 oracle vs fp64: L2 rel {np.linalg.norm(ref32-ref)/np.linalg.norm(ref):.3e}  |ref| {np.linalg.norm(ref):.3e}")
names = []
for k in P.PARAM_ORDER:
    names += [k] * p[k].numel()
names = np.array(names)
for impl in (0, 1):
    N.lib().gs_set_update_impl(impl)
    g_raw, _, m = E.update_step("ppo", E.dev_params(p), batch, hp)
    if impl == 0 and os.environ.get("GS_DEV_SAVE"):
        np.save(os.environ["GS_DEV_SAVE"], g_raw)
    err = np.abs(g_raw - ref)
    print(f"impl {impl}: max|ref| {np.abs(ref).max():.3e}  L2 rel vs fp64 {np.linalg.norm(g_raw-ref)/np.linalg.norm(ref):.3e}  vs torch fp32 {np.linalg.norm(g_raw-ref32)/np.linalg.norm(ref32):.3e}  max abs err / max|ref| {np.abs(g_raw-ref).max()/np.abs(ref).max():.3e}")
    for k in P.PARAM_ORDER:
        sel_ = names == k
        print(f"   {k}: max abs err {err[sel_].max():.3e}  max|ref| {np.abs(ref[sel_]).max():.3e}  mean signed rel {np.mean((g_raw[sel_]-ref[sel_])/ (np.abs(ref[sel_])+1e-12)):.2e}")
