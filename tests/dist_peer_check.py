"""Run under torchrun with >= 2 GPUs: trains the same sharded CartPole PPO job twice — gradient mean through the NVLink peer
exchange fused into gs_update_finish, and through NCCL on the generic path — and checks that (1) every rank holds
bit-identical weights after each iteration, (2) both paths agree to rounding, (3) the global-minibatch statistics make the
2-rank update match a 1-rank update of the same global minibatch (rank 0 recomputes it from gathered shards).
Not collected by pytest (no test_ prefix); tests/test_gpu_agent.py launches it when the box has two GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gymnasium_solver_b200.agents import build_agent
    from gymnasium_solver_b200.utils.config import load_config
    from gymnasium_solver_b200.utils.model_registry import resolve_model_spec

    def make(mode, model_id="mlp_64x64"):
        cfg = load_config("CartPole-v1", "ppo")
        cfg.n_envs, cfg.n_steps, cfg.batch_size, cfg.n_epochs = 256 * world, 32, 2048 * world, 3
        cfg.model_id, cfg._hidden_dims = model_id, resolve_model_spec(model_id).hidden_dims
        cfg.eval_freq_epochs = None
        cfg.grad_allreduce = mode
        cfg.fused_update = mode == "peer"
        cfg.validate()
        return build_agent(cfg, rank=rank, world_size=world)

    for model_id in ("mlp_64x64", "mlp_small"):
        a, b = make("peer", model_id), make("nccl", model_id)
        assert a._peer is not None and b._peer is None
        for it in range(3):
            a.train_one_rollout()
            b.train_one_rollout()
            torch.cuda.synchronize()
            for ag, name in ((a, "peer"), (b, "nccl")):
                w = ag.policy_model.flat_params.clone()
                ws = [torch.empty_like(w) for _ in range(world)]
                dist.all_gather(ws, w)
                for r in range(1, world):
                    assert torch.equal(ws[0], ws[r]), f"{name}: rank {r} weights differ from rank 0 after iteration {it}"
            wa, wb = a.policy_model.flat_params.cpu().numpy(), b.policy_model.flat_params.cpu().numpy()
            np.testing.assert_allclose(wa, wb, rtol=5e-5, atol=5e-6, err_msg=f"{model_id}: peer vs nccl after iteration {it}")
        ma, mb = a.pop_epoch_metrics(), b.pop_epoch_metrics()
        for k in ("opt/grads/norm/all", "roll/adv/norm/std"):
            np.testing.assert_allclose(ma[k], mb[k], rtol=1e-4, err_msg=k)
        assert abs(ma["roll/adv/norm/std"] - 1.0) < 0.2
    dist.barrier()
    if rank == 0:
        print("PEER_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
