"""dev helper: device time of each PPOAgent.train_one_rollout() on the bench workload, plus the host time to enqueue it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = bench.parse_args([]) if hasattr(bench, "parse_args") else None
if args is None:
    import argparse
    raise SystemExit("bench.parse_args missing")
agent, cfg = bench.build_agent_for_bench(args, 0, 1)
torch.cuda.synchronize()
for i in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    agent.train_one_rollout()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: device {e0.elapsed_time(e1):7.2f} ms   host enqueue {1e3*(t1-t0):7.2f} ms   wall {1e3*(t2-t0):7.2f} ms", flush=True)
