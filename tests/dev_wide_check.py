"""dev helper: gs_ppo_step on the 256x256 tensor path vs the fp32 SIMT kernel (GS_UPDATE_IMPL=simt) on the same minibatch.

usage: python tests/dev_wide_check.py [n_samples ...]"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] != "--child":
    import numpy as np
    for n in sys.argv[1:]:
        out = {}
        for impl in ("tc", "simt"):
            env = dict(os.environ, GS_UPDATE_IMPL=impl, GS_DEV_N=n, GS_DEV_OUT=f"/tmp/wide_{impl}.npz")
            subprocess.run([sys.executable, __file__, "--child"], env=env, check=True)
            out[impl] = np.load(f"/tmp/wide_{impl}.npz")
        g0, g1 = out["tc"]["g"].astype(np.float64), out["simt"]["g"].astype(np.float64)
        po = out["tc"]["po"]
        names = ["w1", "b1", "w2", "b2", "wp", "bp", "wv", "bv"]
        print(f"n={n}: |g_simt| {np.linalg.norm(g1):.6e} |g_tc| {np.linalg.norm(g0):.6e} rel L2 err {np.linalg.norm(g0 - g1) / np.linalg.norm(g1):.3e}")
        for k, nm in enumerate(names):
            a, b = int(po[k]), int(po[k + 1]) if k + 1 < len(po) else len(g0)
            if b > a:
                print(f"   {nm:3s} |ref| {np.linalg.norm(g1[a:b]):.4e} err {np.linalg.norm(g0[a:b] - g1[a:b]):.4e} max {np.abs(g0[a:b] - g1[a:b]).max():.3e}")
        m0, m1 = out["tc"]["m"], out["simt"]["m"]
        rel = np.abs(m0 - m1)[:32] / (np.abs(m1[:32]) + 1e-9)
        print("   metrics rel err > 1e-4:", {int(k): (float(m0[k]), float(m1[k])) for k in np.nonzero(rel > 1e-4)[0]})
    sys.exit(0)
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
from gymnasium_solver_b200 import _native as N
import ctypes as C, numpy as np, torch
import engine_api as E
from oracle import policy as P
n = int(os.environ["GS_DEV_N"])
T, D, A = 128, 2, 3
Nn = max(1024, (n + T - 1) // T)
g = torch.Generator().manual_seed(0)
p = P.random_params(D, (256, 256), A, seed=1)
obs = torch.randn(T, Nn, D, generator=g) * 0.5
actions = torch.randint(0, A, (T, Nn), generator=g)
z = torch.randn(T, Nn, generator=g)
dev = [E.cu(obs), E.cu(actions.int()), E.cu(z * 0.1 - 1.0), E.cu(z), E.cu(z + 0.3), E.cu(z * 2)]
batch, keep = E.make_batch(T, Nn, *dev, n=n, perm_key=77, perm_offset=0, perm_len=T * Nn)
E.pack_rollout(batch, keep)
hp = N.GsPpoHparams(); hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef, hp.normalize_adv, hp.track_activations = 0.2, 0.2, 0.5, 0.01, 1, 1
pd = E.dev_params(p)      # keep the device tensors alive: the struct holds raw pointers
m = N.mlp_struct_from_params(pd, "relu")
Pn = N.lib().gs_mlp_param_count(C.byref(m))
wsb = N.lib().gs_update_workspace_bytes(C.byref(m), 0, n)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); grads = torch.zeros(Pn, device="cuda"); met = torch.zeros(N.N_METRICS, dtype=torch.float64, device="cuda")
mom = torch.tensor([0.3 * n, 1.1 * n, float(n)], dtype=torch.float64, device="cuda")
for _ in range(2):
    N.check(N.lib().gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(mom), N.ptr(grads), N.ptr(met), N.ptr(ws), wsb, N.stream()))
torch.cuda.synchronize()
H = 256
po = np.cumsum([0, H * D, H, H * H, H, A * H, A, H, 1])
np.savez(os.environ["GS_DEV_OUT"], g=grads.cpu().numpy(), m=met.cpu().numpy(), po=po)
