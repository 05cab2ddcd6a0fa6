"""Multi-GPU sharding of a batch of independent QP instances.

Every instance is independent, so the batch is cut into contiguous blocks, one per rank (one process per GPU); nothing is
exchanged during the solve and the results are collected with a single all-gather per step (NCCL over NVLink / NVSwitch
on the GPU box, gloo in the CPU tests).

``BatchGatherer`` is the collective of the path: buffers allocated once, the per-rank block and its status words packed
into one message, the all-gather issued on a side stream so that the gather of step k overlaps the kernels of step k + 1
(the first version allocated a zero pad, gathered on the compute stream and concatenated the pieces every step: a
constant ~1.7 ms of launch / allocator latency per step at every world size, SCALE_r01 efficiency 0.95)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; the first total % world ranks get one extra item."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist_on(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class BatchGatherer:
    """All-gather of the blocks ``values [n_local, width]`` (float64) and ``status [n_local]`` (int32) of a batch of
    ``total`` rows cut with shard_range().

    One collective per call: every rank contributes one ``[cap, width + 1]`` float64 message (cap = largest shard; the
    last column carries the status words), gathered straight into the pre-allocated ``[world * cap, width + 1]`` result.
    With equal shards the result rows ARE the batch rows (no compaction); otherwise the ``world`` pieces are copied into
    a second pre-allocated buffer.  On CUDA the packing copy, the collective and the compaction run on a side stream:
    ``start()`` returns immediately, ``finish()`` makes the current stream wait for the result."""

    def __init__(self, total: int, width: int, device, group=None):
        self.total, self.width, self.group = int(total), int(width), group
        self.device = torch.device(device)
        self.on = _dist_on(group)
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.cap = (self.total + self.world - 1) // self.world
        self.equal = (self.total % self.world == 0)
        self.lo, self.hi = shard_range(self.total, self.rank, self.world)
        self.cuda = self.device.type == "cuda"
        if self.on:
            f64 = dict(dtype=torch.float64, device=self.device)
            self._msg = torch.zeros((self.cap, self.width + 1), **f64)
            self._all = torch.empty((self.world * self.cap, self.width + 1), **f64)
            self._out = self._all if self.equal else torch.empty((self.total, self.width + 1), **f64)
            self._side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._pending = False

    def start(self, values: torch.Tensor, status: Optional[torch.Tensor] = None) -> None:
        """Begin gathering this rank's block (rows lo..hi of the batch).  The tensors may be re-used by the caller after
        the call (they are copied into the message buffer on the side stream, which has them recorded)."""
        n = self.hi - self.lo
        if values.shape[0] != n or values.shape[1] != self.width:
            raise ValueError(f"BatchGatherer: expected a [{n}, {self.width}] block, got {tuple(values.shape)}")
        if not self.on:
            self._local = (values, status)
            self._pending = True
            return
        if self._pending:
            raise RuntimeError("BatchGatherer.start() called again before finish()")

        def body():
            self._msg[:n, : self.width].copy_(values)
            if status is not None:
                self._msg[:n, self.width].copy_(status)
            dist.all_gather_into_tensor(self._all, self._msg, group=self.group)
            if not self.equal:
                for r in range(self.world):
                    s, e = shard_range(self.total, r, self.world)
                    self._out[s:e].copy_(self._all[r * self.cap: r * self.cap + (e - s)])

        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            self._side.wait_stream(cur)                 # the block is complete on the compute stream
            with torch.cuda.stream(self._side):
                values.record_stream(self._side)
                if status is not None:
                    status.record_stream(self._side)
                body()
        else:
            body()
        self._pending = True

    def finish(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(values [total, width] float64 view, status [total] int32) of the whole batch; the current stream waits for
        the gather that start() began."""
        if not self._pending:
            raise RuntimeError("BatchGatherer.finish() without start()")
        self._pending = False
        if not self.on:
            values, status = self._local
            st = status if status is not None else torch.zeros((values.shape[0],), dtype=torch.int32, device=values.device)
            return values, st
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        return self._out[:, : self.width], self._out[:, self.width].to(torch.int32)


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Convenience form for one tensor of any trailing shape / dtype: all-gather of the per-rank blocks [n_local, ...] of
    a batch cut with shard_range() into the full [total, ...] tensor (one collective; synchronous with the current
    stream).  The bench and long-running callers use BatchGatherer (pre-allocated, overlapped)."""
    if not _dist_on(group):
        return local
    world = dist.get_world_size(group)
    cap = (total + world - 1) // world
    if total % world == 0:                       # equal shards: gather straight into the result
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.empty((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    pad[local.shape[0]:] = 0
    allb = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allb, pad, group=group)
    out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        s, e = shard_range(total, r, world)
        out[s:e] = allb[r * cap: r * cap + (e - s)]
    return out
