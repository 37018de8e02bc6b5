"""World-size-2 gloo tests (CPU) of the data-parallel host logic: env sharding arithmetic, gradient averaging and global
minibatch moments.  The kernels themselves are exercised on the GPU box; this covers the N > 1 plumbing."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gymnasium_solver_b200.utils.distributed import (agree_any, allreduce_moments, average_gradients, broadcast_value, exchange_handles,
                                                      max_over_ranks, shard_spec)


def test_shard_spec_partitions_envs_and_batches():
    shards = [shard_spec(65536 * 8, 1048576 * 8, r, 8) for r in range(8)]
    assert [s.env_id_offset for s in shards] == [r * 65536 for r in range(8)]
    assert all(s.n_envs == 65536 and s.batch_size == 1048576 for s in shards)
    assert sum(s.n_envs for s in shards) == 65536 * 8
    with pytest.raises(ValueError):
        shard_spec(10, 8, 0, 4)
    with pytest.raises(ValueError):
        shard_spec(8, 10, 0, 4)
    with pytest.raises(ValueError):
        shard_spec(8, 8, 4, 4)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        grads = torch.randn(4675, generator=g)
        local = grads.clone()
        average_gradients(grads, world)
        # per-rank (sum, sumsq, count) of that rank's minibatch shard -> global moments
        x = torch.randn(1000 + 10 * rank, generator=g, dtype=torch.float64)
        mom = torch.tensor([x.sum(), (x * x).sum(), float(x.numel())], dtype=torch.float64)
        allreduce_moments(mom, world)
        t = max_over_ranks(1.0 + rank, torch.device("cpu"), world)
        # the 64-byte IPC handles of the NVLink peer group travel through the same process group, concatenated in rank order
        handles = exchange_handles(bytes([rank + 1]) * 64, world, torch.device("cpu"))
        assert handles == b"".join(bytes([r + 1]) * 64 for r in range(world))
        # lock-step exits: a stop decision only one rank reached is taken by all; a checkpoint decision follows rank 0's metric
        stop_seq = [agree_any(rank == 1 and epoch == 3, torch.device("cpu"), world) for epoch in range(5)]
        assert stop_seq == [False, False, False, True, False]
        assert broadcast_value(10.0 + rank, torch.device("cpu"), world) == 10.0
        # live hyper-parameters: every rank applies RANK 0's reading of the run's config file (the ranks' weights must stay identical)
        from gymnasium_solver_b200.agents.hyperparameter_mixin import HyperparameterMixin

        class _Run:
            def load_config(self): return {"policy_lr": 1e-3 * (rank + 1), "clip_range": 0.2}

        class _Cfg:
            policy_lr, clip_range = 5e-3, 0.2

        class _Agent(HyperparameterMixin):
            def __init__(self):
                self.config, self.run, self.rank, self.world_size, self.policy_lr, self._optimizer = _Cfg(), _Run(), rank, world, 5e-3, None
            def optimizers(self): return []

        a = _Agent()
        a._read_hyperparameters_from_run()
        assert a.policy_lr == 1e-3 and a.config.policy_lr == 1e-3, a.policy_lr
        out.put((rank, local.numpy(), grads.numpy(), x.numpy(), mom.numpy(), t))
    finally:
        dist.destroy_process_group()


def test_gradient_average_and_global_moments_world2():
    world, port = 2, 29500 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (res[0][1] + res[1][1]) / 2
    for r in res:
        np.testing.assert_allclose(r[2], mean, rtol=1e-6, atol=1e-7)       # every rank holds the same averaged gradient
    np.testing.assert_array_equal(res[0][2], res[1][2])                      # bit-identical -> identical optimizer steps
    allx = np.concatenate([res[0][3], res[1][3]])
    for r in res:
        np.testing.assert_allclose(r[4], [allx.sum(), (allx ** 2).sum(), allx.size], rtol=1e-12)
        assert r[5] == 2.0


def test_keyed_bijection_mirror_is_a_permutation_and_mixes():
    """Host mirror of the in-kernel sample-id bijection (csrc/common.cuh::feistel_permute): a permutation for every length, a
    different one per key, and without the structure a weak mixer would leave (order / neighbour correlation)."""
    import torch
    from gymnasium_solver_b200.utils.samplers import feistel_permutation

    for length in (1, 2, 3, 5, 7, 96, 1000, 4096, 32769, 1 << 17):
        perm = feistel_permutation(length, key=0xDEADBEEF12345678 & ((1 << 63) - 1))
        assert perm.dtype == torch.int64 and perm.numel() == length
        assert torch.equal(torch.sort(perm).values, torch.arange(length))
    n = 1 << 16
    a, b = feistel_permutation(n, key=1), feistel_permutation(n, key=2)
    assert float((a == b).float().mean()) < 1e-3
    x = torch.arange(n, dtype=torch.float64)
    for p in (a, b):
        pf = p.double()
        corr = float(torch.corrcoef(torch.stack([x, pf]))[0, 1])
        assert abs(corr) < 0.02, corr
        d = (pf[1:] - pf[:-1]).abs()
        assert 0.30 * n < float(d.mean()) < 0.37 * n           # neighbours land n/3 apart on average, like a random permutation
        # every 1/8 slice of positions draws about 1/8 of its samples from every 1/8 slice of ids (minibatch = unbiased subset)
        counts = torch.zeros(8, 8)
        for k in range(8):
            counts[k] = torch.bincount((p[k * n // 8:(k + 1) * n // 8] * 8 // n), minlength=8).float()
        assert float((counts / (n / 64) - 1).abs().max()) < 0.15


def test_stop_agreement_is_a_no_op_for_one_rank():
    assert agree_any(True, torch.device("cpu"), 1) is True and agree_any(False, torch.device("cpu"), 1) is False
    assert broadcast_value(3.5, torch.device("cpu"), 1) == 3.5
