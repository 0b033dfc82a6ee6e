"""Host-side logic of the multi-GPU path: replicas only, one process per GPU.

* one broadcast of rank 0's weight arena fills every replica: on GPUs that is `hb_model_load_broadcast` (ncclBroadcast
  issued inside libhelixb200.so, engine.Engine.load_broadcast); `broadcast_buffer` below is the same step over a
  torch.distributed group and only serves the CPU (gloo) test of the host logic;
* sessions are routed to replicas the way the scheduler routes requests to warm slots
  (api/pkg/scheduler/scheduler.go:1958-2009: fewest active requests first, ties by lowest load, then order);
* throughput of the job = all units processed / max-over-ranks time.
"""
from typing import List, Sequence


def shard_range(n_units: int, world: int, rank: int):
    """Contiguous index range of `n_units` independent units (sessions, chunks) owned by `rank`."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def route_least_active(active: Sequence[int], n_new: int) -> List[int]:
    """Replica index for each of n_new sessions: least active requests first (pickBestWarmSlot), stable ties."""
    act = list(active)
    out = []
    for _ in range(n_new):
        i = min(range(len(act)), key=lambda k: (act[k], k))
        out.append(i)
        act[i] += 1
    return out


def broadcast_buffer(dist, tensor, src=0):
    """One collective moves the whole weight arena (a flat uint8 view) from `src` to every replica."""
    dist.broadcast(tensor, src=src)
    return tensor


def aggregate_throughput(dist, units_local: float, seconds_local: float, device="cpu"):
    """(total units over ranks, max seconds over ranks, units/s)."""
    import torch
    t = torch.tensor([seconds_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u), float(t), float(u) / float(t)


class ArenaView:
    """Exposes a raw device allocation (the engine's weight arena) to torch without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
