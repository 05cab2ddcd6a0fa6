"""Multi-GPU sharding of a batch of independent QP instances.

Every instance is independent, so the batch is cut into contiguous blocks, one per rank
(one process per GPU); nothing is exchanged during the solve and the results are collected with a
single all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; the first total % world ranks get one extra item."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank blocks [n_local, ...] of a batch that was cut with shard_range()
    into the full [total, ...] tensor (one collective; blocks are padded to the largest shard)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    cap = (total + world - 1) // world
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        s, e = shard_range(total, r, world)
        pieces.append(out[r * cap: r * cap + (e - s)])
    del rank
    return torch.cat(pieces, dim=0)
