"""dev helper: where a minibatch's time goes on the bench workload — device time of the step kernels alone, of the fused
tail alone (CUDA events around back-to-back calls), and host enqueue time per minibatch."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gymnasium_solver_b200 import _native as N
args = bench.parse_args([])
agent, cfg = bench.build_agent_for_bench(args, 0, 1)
for _ in range(2):
    agent.train_one_rollout()
torch.cuda.synchronize()
col = agent.get_rollout_collector("train")
traj = col.collect()
agent._pack_rollout(traj)
batches = [b for _, _, b in agent.minibatches(traj, 777)]
def ev(fn, reps):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, (t1 - t0) / reps * 1e6
i = [0]
def step_only():
    agent._launch_step(batches[i[0] % len(batches)], defer=True); i[0] += 1
def fused():
    agent._fused_training_step(batches[i[0] % len(batches)]); i[0] += 1
def generic():
    r = agent.losses_for_batch(batches[i[0] % len(batches)], 0); agent._backpropagate_and_step(r["loss"]); i[0] += 1
for name, fn in (("step kernels only (gather + update, deferred)", step_only), ("fused training step", fused), ("generic training step", generic)):
    d, h = ev(fn, 80)
    print(f"{name:50s} device {d:8.1f} us/minibatch   host enqueue {h:8.1f} us/minibatch", flush=True)
# finish kernel alone: call it repeatedly on the partials the last step left
b = batches[0]
fin = agent._launch_step(b, defer=True)
mlp, adam = N.mlp_struct(agent.policy_model), agent.optimizers().adam_struct()
adam.lr = 0.0
def finish_only():
    N.check(N.lib().gs_update_finish(C.byref(mlp), C.byref(b.struct), C.byref(fin), N.ptr(agent.policy_model.flat_grads), C.byref(adam), None,
                                     N.ptr(agent._metrics_dev), None, N.ptr(agent._workspace), agent._ws_bytes, N.stream()))
d, h = ev(finish_only, 200)
print(f"{'gs_update_finish alone':50s} device {d:8.1f} us/call        host enqueue {h:8.1f} us/call", flush=True)
for k in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); agent.train_one_rollout(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"iteration: device {e0.elapsed_time(e1):7.2f} ms   host enqueue {1e3*(t1-t0):7.2f} ms", flush=True)
