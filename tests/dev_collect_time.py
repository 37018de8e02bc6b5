"""dev helper: device time of RolloutCollector.collect() pieces on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = bench.parse_args([])
agent, cfg = bench.build_agent_for_bench(args, 0, 1)
col = agent.get_rollout_collector("train")
for _ in range(3):
    col.collect()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    col.collect()
b.record(); torch.cuda.synchronize()
print(f"collect() (collect kernel + GAE + moments): {a.elapsed_time(b) / 10:.3f} ms", flush=True)
