"""Run under torchrun with >= 2 GPUs.  Checks of the NVLink peer gradient exchange fused into gs_update_finish:
  1. stress: 400 back-to-back exchanges of random per-rank gradients (no optimizer) — every rank must hold the bit-identical
     mean, equal to the rank-ordered fp32 sum computed from an all_gather of the inputs; then gs_peer_allreduce_f64 (the moments
     exchange) against the process group's all-reduce;
  2. training: a sharded CartPole PPO job through the peer path keeps bit-identical weights on every rank, and one iteration
     from identical state agrees with the NCCL generic path to rounding (state is re-synchronised before every iteration so
     sampling flips cannot amplify rounding differences).
Not collected by pytest (no test_ prefix); tests/test_gpu_agent.py launches it when the box has two GPUs."""
import ctypes as C
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stress(rank, world, dev):
    from gymnasium_solver_b200 import _native as N
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config
    from gymnasium_solver_b200.utils.model_registry import resolve_model_spec

    cfg = load_config("CartPole-v1", "ppo")
    cfg.n_envs, cfg.n_steps, cfg.batch_size, cfg.n_epochs = 128 * world, 32, 4096 * world, 1
    cfg.model_id, cfg._hidden_dims = "mlp_medium", resolve_model_spec("mlp_medium").hidden_dims
    cfg._activation = "tanh"          # 256x256 tanh runs the FMA-pipe kernel (atomic accumulation): gs_update_finish takes the gradients from grads_flat
    cfg.eval_freq_epochs, cfg.grad_allreduce = None, "peer"
    cfg.validate()
    agent = build_agent(cfg, rank=rank, world_size=world)
    model = agent.policy_model
    P = model.flat_params.numel()
    traj = agent.get_rollout_collector("train").collect()
    b = next(iter(agent.minibatches(traj, 1)))[2]
    fin = agent._launch_step(b, defer=True, moments=agent._global_moments(b))     # leaves valid metric partials in the workspace
    fin.max_grad_norm = 0.0
    mlp = N.mlp_struct(model)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    for it in range(400):
        local = torch.randn(P, generator=g, device=dev) * (1.0 + it % 7)
        model.flat_grads.copy_(local)
        N.check(N.lib().gs_update_finish(C.byref(mlp), C.byref(b.struct), C.byref(fin), N.ptr(model.flat_grads), None, agent._peer.handle,
                                         N.ptr(agent._metrics_dev), None, N.ptr(agent._workspace), agent._ws_bytes, N.stream()))
        if it % 20 == 0 or it > 380:
            parts = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(parts, local)
            ref = torch.zeros_like(local)
            for r in range(world):
                ref = ref + parts[r]
            ref = ref * (1.0 / world)
            assert torch.equal(model.flat_grads, ref), f"stress {it}: max diff {(model.flat_grads - ref).abs().max().item()}"
    torch.cuda.synchronize()
    moments(rank, world, dev, agent._peer)


def moments(rank, world, dev, peer):
    """gs_peer_allreduce_f64 against the process group's all-reduce: same values (to the last bits of a different summation order), identical
    on every rank, over back-to-back calls of different lengths."""
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    for it in range(60):
        n = (6, 48, 480, 768, 8192)[it % 5]
        x = torch.randn(n, generator=g, device=dev, dtype=torch.float64) * (1.0 + it)
        ref = x.clone()
        dist.all_reduce(ref)
        y = x.clone()
        assert peer.allreduce_f64(y), "peer all-reduce refused a qualifying tensor"
        torch.testing.assert_close(y, ref, rtol=1e-14, atol=1e-12)
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y)
        for r in range(1, world):
            assert torch.equal(ys[0], ys[r]), f"moments {it}: rank {r} differs from rank 0"
    assert not peer.allreduce_f64(torch.zeros(8193, dtype=torch.float64, device=dev))      # too long: the caller falls back
    torch.cuda.synchronize()


def training(rank, world, dev):
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config
    from gymnasium_solver_b200.utils.model_registry import resolve_model_spec

    def make(mode, model_id, activation):
        cfg = load_config("CartPole-v1", "ppo")
        cfg.n_envs, cfg.n_steps, cfg.batch_size, cfg.n_epochs = 256 * world, 32, 2048 * world, 3
        cfg.model_id, cfg._hidden_dims = model_id, resolve_model_spec(model_id).hidden_dims
        cfg._activation = activation
        cfg.eval_freq_epochs = None
        cfg.grad_allreduce = mode
        cfg.fused_update = mode == "peer"
        cfg.validate()
        return build_agent(cfg, rank=rank, world_size=world)

    # two-set / one-set tensor-core kernels, the 256-wide pair of kernels (update_wide.cu), the FMA-pipe kernel
    for model_id, activation in (("mlp_64x64", "relu"), ("mlp_small", "relu"), ("mlp_medium", "relu"), ("mlp_medium", "tanh")):
        a, b = make("peer", model_id, activation), make("nccl", model_id, activation)
        assert a._peer is not None and a.grad_allreduce_mode == "peer" and b._peer is None and b.grad_allreduce_mode == "nccl"
        for it in range(4):
            # identical state before the iteration: weights, Adam moments, step count (envs and RNG counters evolve identically)
            b.policy_model.flat_params.copy_(a.policy_model.flat_params)
            oa, ob = a.optimizers(), b.optimizers()
            ob.exp_avg.copy_(oa.exp_avg); ob.exp_avg_sq.copy_(oa.exp_avg_sq); ob.step_dev.copy_(oa.step_dev)
            ta, tb = a.train_one_rollout(), b.train_one_rollout()
            torch.cuda.synchronize()
            assert torch.equal(ta.tm["obs"], tb.tm["obs"]) and torch.equal(ta.tm["actions"], tb.tm["actions"]), f"{model_id}: rollouts differ at {it}"
            for ag, name in ((a, "peer"), (b, "nccl")):
                w = ag.policy_model.flat_params.clone()
                ws = [torch.empty_like(w) for _ in range(world)]
                dist.all_gather(ws, w)
                for r in range(1, world):
                    assert torch.equal(ws[0], ws[r]), f"{model_id} {name}: rank {r} weights differ from rank 0 after iteration {it}"
            wa, wb = a.policy_model.flat_params.cpu().numpy(), b.policy_model.flat_params.cpu().numpy()
            np.testing.assert_allclose(wa, wb, rtol=2e-5, atol=2e-6, err_msg=f"{model_id}: peer vs nccl after iteration {it}")
            ma, mb = a.pop_epoch_metrics(), b.pop_epoch_metrics()
            for k in ("opt/grads/norm/all", "roll/adv/norm/std", "opt/loss/total"):
                np.testing.assert_allclose(ma[k], mb[k], rtol=1e-4, atol=1e-6, err_msg=k)
            assert abs(ma["roll/adv/norm/std"] - 1.0) < 0.2     # statistics of the GLOBAL minibatch


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        stress(rank, world, dev)
        training(rank, world, dev)
    except Exception:
        print(f"PEER_CHECK_FAILED rank {rank}\n{traceback.format_exc()}", flush=True)
        raise
    dist.barrier()
    if rank == 0:
        print("PEER_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
