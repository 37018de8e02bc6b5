"""dev helper: host-side time of the pieces of one training iteration (where does the host stall the GPU?)."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = bench.parse_args([])
agent, cfg = bench.build_agent_for_bench(args, 0, 1)
col = agent.get_rollout_collector("train")
for _ in range(3):
    agent.train_one_rollout()
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "nogc":
    gc.collect(); gc.freeze(); gc.disable()
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    col._sync_device_and_prepare_buffers()
    col._resolve_pending_episodes()
    t1 = time.perf_counter()
    traj = col.collect()
    t2 = time.perf_counter()
    agent._pack_rollout(traj)
    key = 1234 + i
    batches = [b for _, _, b in agent.minibatches(traj, key)]
    t3 = time.perf_counter()
    for k, b in enumerate(batches):
        agent.training_step(b, k)
    e1.record(); t4 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"it {i}: device {e0.elapsed_time(e1):6.2f} ms | host: resolve {1e3*(t1-t0):6.2f}  collect {1e3*(t2-t1):6.2f}  batches {1e3*(t3-t2):6.2f}  loop {1e3*(t4-t3):6.2f} ms  gc {gc.get_count()}", flush=True)
