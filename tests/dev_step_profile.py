"""dev helper: per-kernel device time of one training iteration (torch profiler / CUPTI), 1 or N ranks (torchrun).  Rank 0 prints."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
args = argparse.Namespace(n_envs=int(os.environ.get("GS_DEV_ENVS", "65536")), n_steps=128, batch_size=1048576, n_epochs=10, model_id=os.environ.get("GS_MODEL", "mlp_64x64"),
                          track_activations=1, config=os.environ.get("GS_DEV_CONFIG", "c2"))
agent, cfg = bench.build_agent_for_bench(args, rank, world)
for _ in range(int(os.environ.get("GS_DEV_WARM", "3"))):
    agent.train_one_rollout()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    agent.train_one_rollout()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
