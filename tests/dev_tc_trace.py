"""dev helper: per-warp timeline of one pipeline iteration of the tensor-core update kernel.

Build the library with `GS_NVCC_EXTRA=-DGS_TC_TRACE python -m gymnasium_solver_b200.build --force`, run this on the GPU, then
rebuild normally (`python -m gymnasium_solver_b200.build --force`).
"""
import ctypes, os, sys
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
z = torch.randn(T, Nn, generator=g)
dev = [E.cu(obs), E.cu(actions.int()), E.cu(z * 0.1 - 0.7), E.cu(z), E.cu(z + 0.3), E.cu(z * 2)]
batch, keep = E.make_batch(T, Nn, *dev, n=1 << 20, perm_key=77, perm_offset=0, perm_len=T * Nn)
E.pack_rollout(batch, keep)
hp = N.GsPpoHparams(); hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef, hp.normalize_adv, hp.track_activations = 0.2, 0.2, 0.5, 0.01, 1, 1
for _ in range(3):
    E.update_step("ppo", E.dev_params(p), batch, hp)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 432)()
rc = N.lib().gs_debug_tc_trace(buf)
tr = np.array(buf[:]).reshape(18, 24)[:16]
t0 = tr[:, 0].min()
order = [0, 16, 17, 18, 19, 1, 2, 3, 20, 21, 22, 23, 4, 7, 8, 9, 10, 11, 12, 13, 14, 15]
names = ["top", "B1 done", "sync1", "issued A", "F1 done", "", "", "sync2", "issued fwd", "BAR_FWD ok", "F2 done", "sync3", "F3 done",
         "BAR_D ok+dz1", "BAR_W ok", "P,S written", "dz2->tmem", "BAR_T ok", "dz2T->P", "L1 recomputed", "sx loaded", "pf stage0", "L1(u)", "stats(u)"]
print("cycles since the earliest warp entered the iteration (rows: trace point, cols: warp 0..15)")
for k in order:
    print(f"{names[k]:>13s} " + " ".join(f"{int(v - t0):6d}" for v in tr[:, k]))
