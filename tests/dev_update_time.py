"""dev helper: time gs_ppo_step alone on a C2-sized minibatch (1,048,576 samples gathered from a 128 x 65,536 rollout).

usage: python tests/dev_update_time.py [path/to/libgs_engine.so ...]   (each library is timed in a fresh subprocess)
"""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        env = dict(os.environ, GS_DEV_LIB=lib)
        subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shutil
from gymnasium_solver_b200 import _native as N
lib = os.environ.get("GS_DEV_LIB")
if lib:
    shutil.copy(lib, N.LIB_PATH)
import ctypes as C, numpy as np, torch
import engine_api as E
from oracle import policy as P
T, Nn, D, A = 128, 65536, 4, 2
g = torch.Generator().manual_seed(0)
HID = int(os.environ.get("GS_DEV_HIDDEN", "64"))
p = P.random_params(D, (HID, HID), A, seed=1)
obs = torch.randn(T, Nn, D, generator=g) * 0.5
actions = torch.randint(0, A, (T, Nn), generator=g)
z = torch.randn(T, Nn, generator=g)
dev = [E.cu(obs), E.cu(actions.int()), E.cu(z * 0.1 - 0.7), E.cu(z), E.cu(z + 0.3), E.cu(z * 2)]
n = 1 << 20
batch, keep = E.make_batch(T, Nn, *dev, n=n, perm_key=77, perm_offset=0, perm_len=T * Nn)
if not os.environ.get("GS_DEV_UNPACKED"):
    E.pack_rollout(batch, keep)      # as the agent does once per rollout: the kernel gathers 64-byte records
hp_track = int(os.environ.get("GS_DEV_TRACK", "1"))
hp = N.GsPpoHparams(); hp.clip_range, hp.clip_range_vf, hp.vf_coef, hp.ent_coef, hp.normalize_adv, hp.track_activations = 0.2, 0.2, 0.5, 0.01, 1, hp_track
pd = E.dev_params(p)      # keep the device tensors alive: the struct holds raw pointers
m = N.mlp_struct_from_params(pd, "relu")
Pn = N.lib().gs_mlp_param_count(C.byref(m))
wsb = N.lib().gs_update_workspace_bytes(C.byref(m), 0, n)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); grads = torch.empty(Pn, device="cuda"); met = torch.zeros(N.N_METRICS, dtype=torch.float64, device="cuda")
mom = torch.tensor([0.3 * n, 1.1 * n, float(n)], dtype=torch.float64, device="cuda")
def step():
    N.check(N.lib().gs_ppo_step(C.byref(m), C.byref(batch), C.byref(hp), N.ptr(mom), N.ptr(grads), N.ptr(met), N.ptr(ws), wsb, N.stream()))
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 40
e0.record()
for _ in range(K): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{os.path.basename(lib or 'default')}: gs_ppo_step {ms*1000:.1f} us / 1M-sample minibatch  ({6 * n * (D * HID + HID * HID + (A + 1) * HID) / ms / 1e9:.1f} TFLOP/s algorithmic)  |grad| {float(grads.norm()):.6e}")
if os.environ.get("GS_DEV_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
