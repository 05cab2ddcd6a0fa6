"""world_size-2 gloo test (CPU) of the multi-GPU sharding logic: contiguous blocks + one all-gather."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from global_racetrajectory_optimization_b200.sharding import BatchGatherer, gather_batch, shard_range


def test_shard_range_partitions_exactly():
    for total in [1, 7, 8, 1024, 1025]:
        for world in [1, 2, 3, 8]:
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(total, rank, world)
    full = torch.arange(total * 3, dtype=torch.float64).reshape(total, 3)
    out = gather_batch(full[s:e].clone(), total)
    ok = torch.equal(out, full)
    st = gather_batch(torch.full((e - s,), rank, dtype=torch.int32), total)
    owner = [r for r in range(world) for _ in range(*shard_range(total, r, world))]
    ok = ok and st.tolist() == owner
    # the collective of the path: values + status in one pre-allocated message, twice (buffers are re-used)
    g = BatchGatherer(total, 3, "cpu")
    for rep in range(2):
        g.start(full[s:e] + rep, torch.full((e - s,), rank + 10 * rep, dtype=torch.int32))
        vals, stat = g.finish()
        ok = ok and torch.equal(vals, full + rep) and stat.tolist() == [o + 10 * rep for o in owner] and stat.dtype == torch.int32
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_batch_gatherer_without_a_process_group_is_the_identity():
    g = BatchGatherer(5, 3, "cpu")
    v, st = torch.rand((5, 3), dtype=torch.float64), torch.arange(5, dtype=torch.int32)
    g.start(v, st)
    a, b = g.finish()
    assert a is v and b is st
    with pytest.raises(RuntimeError):
        g.finish()
    with pytest.raises(ValueError):
        g.start(v[:4], st[:4])


@pytest.mark.parametrize("total", [10, 7])        # equal shards (gather straight into the result) and unequal ones (compaction)
def test_gather_batch_two_ranks_gloo(tmp_path, total):
    world = 2
    port = 29500 + (os.getpid() + total) % 2000
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1", "1"]
