"""Data-parallel plumbing: one process per GPU, environments sharded by rank, gradient averaging per minibatch.

The reference has no distributed code; this is the engine's own layer:
  * rank r owns envs [r*n/W, (r+1)*n/W) — global env ids keep Philox streams and MountainCar count tables identical for
    every W — and draws minibatches of batch/W samples from its own rollout shard; no trajectory data crosses NVLink;
  * the only data-path collective is ``average_gradients`` on the flat gradient buffer (18.7 KB for the 64x64 net), once
    per minibatch, before the global-norm clip — every rank then takes the identical optimizer step, so weights stay
    bit-identical without broadcasts;
  * ``allreduce_moments`` makes "batch" advantage normalisation use the GLOBAL minibatch statistics.
Works with the NCCL backend on GPUs and with gloo on CPU tensors (tests).

``PeerGroup`` is the NVLink exchange the fused step tail (``gs_update_finish``) uses instead of NCCL for the gradient: one
IPC-shared device buffer of receive slots + flags per rank; the kernel stores its gradient into every rank's slot, signals,
waits, and sums the slots in rank order.  torch.distributed only carries the 64-byte IPC handles at start-up."""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    n_envs: int          # local envs
    env_id_offset: int   # global id of local env 0
    batch_size: int      # local minibatch


def shard_spec(n_envs_total: int, batch_size_total: int, rank: int, world_size: int) -> Shard:
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    if n_envs_total % world_size:
        raise ValueError(f"n_envs={n_envs_total} must be divisible by world_size={world_size}")
    if batch_size_total % world_size:
        raise ValueError(f"batch_size={batch_size_total} must be divisible by world_size={world_size}")
    local = n_envs_total // world_size
    return Shard(rank, world_size, local, rank * local, batch_size_total // world_size)


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


class PeerGroup:
    """gs_peer_t of this rank, connected to every other rank of the default process group (one node, <= 8 GPUs)."""

    def __init__(self, rank: int, world_size: int, max_floats: int, device: torch.device, group=None):
        import ctypes as C

        from .. import _native as N

        if not (2 <= world_size <= N.PEER_MAX_WORLD):
            raise N.EngineError(f"peer exchange supports 2..{N.PEER_MAX_WORLD} ranks, got {world_size}")
        self._N, self.handle = N, C.c_void_p()
        mine = (C.c_ubyte * N.PEER_HANDLE_BYTES)()
        err = None
        try:
            N.check(N.lib().gs_peer_create(rank, world_size, int(max_floats), int(device.index), C.byref(self.handle), mine))
        except N.EngineError as e:
            err = e
        # every rank goes through the same collectives whatever happened locally, then all agree on success or failure
        gathered = exchange_handles(bytes(mine), world_size, device, group)
        if err is None:
            try:
                N.check(N.lib().gs_peer_connect(self.handle, gathered))
            except N.EngineError as e:
                err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # also the barrier: nobody signals before every rank has mapped every buffer
        if int(ok.item()) == 0:
            self.close()
            raise N.EngineError(f"NVLink peer group could not be set up on every rank (rank {rank}: {err or 'ok'})")

    MAX_F64 = 8192      # kPeerMomDoubles of csrc/update_kernels.cu

    def allreduce_f64(self, t: torch.Tensor) -> bool:
        """In-place rank-ordered sum of a short contiguous float64 CUDA tensor through the peer buffers (gs_peer_allreduce_f64, on the
        current stream).  False when the tensor does not qualify (the caller falls back to the process group)."""
        if not (self.handle and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= self.MAX_F64):
            return False
        N = self._N
        N.check(N.lib().gs_peer_allreduce_f64(self.handle, N.ptr(t), int(t.numel()), N.stream()))
        return True

    def close(self) -> None:
        if self.handle:
            self._N.lib().gs_peer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_handles(mine: bytes, world_size: int, device, group=None) -> bytes:
    """all_gather of one fixed-size opaque handle per rank, concatenated in rank order (any backend)."""
    backend = dist.get_backend(group)
    dev = device if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(out, t, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def average_gradients(flat_grads: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """In-place mean over ranks of the flat gradient buffer (sum all-reduce, then 1/W)."""
    if world_size > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        flat_grads.mul_(1.0 / world_size)
    return flat_grads


def allreduce_moments(moments: torch.Tensor, world_size: int, group=None, peer: "PeerGroup | None" = None) -> torch.Tensor:
    """In-place sum over ranks of (sum, sumsq, count) triples (any leading shape): through the NVLink peer buffers when a peer group is given
    (a one-block kernel that runs beside the update kernel), else an all-reduce of the process group.  The choice depends only on the
    tensor's shape and on the run's configuration, so every rank takes the same path."""
    if world_size > 1:
        if peer is None or not peer.allreduce_f64(moments):
            dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def max_over_ranks(value: float, device, world_size: int) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _collective_device(device):
    """Tensors of control-plane collectives live where the default group's backend can reduce them."""
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return device
    return torch.device("cpu")


def agree_any(flag: bool, device, world_size: int) -> bool:
    """True on EVERY rank as soon as one rank raises the flag (MAX all-reduce).  Ranks evaluate their own env shard, so a threshold
    on a rank-local mean is crossed at different epochs; every exit of the lock-step training loop has to be taken by all ranks
    together, or the ranks that continue wait forever for the one that left (gs_update_finish spins on its NVLink flag)."""
    if world_size <= 1 or not (dist.is_available() and dist.is_initialized()):
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=_collective_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


def broadcast_value(value: float, device, world_size: int, src: int = 0) -> float:
    """Rank ``src``'s value on every rank: decisions that must be identical everywhere (is this evaluation the best so far?)
    are derived from one rank's metric."""
    if world_size <= 1 or not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_collective_device(device))
    dist.broadcast(t, src=src)
    return float(t.item())
