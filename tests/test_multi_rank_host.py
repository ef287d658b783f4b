"""Host-side logic of the multi-rank path on CPU: the shard rule of the C ABI and the
torch.distributed plumbing (gloo, world_size 2) that ships the communicator id and reduces the
timings.  The data-path collectives themselves run only on GPUs (tests/test_multi_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import ocffm


@pytest.mark.parametrize("rows", [0, 1, 7, 30_000, 360_000, 8_000_001])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_shard_ranges_partition_the_rows(rows, nranks):
    spans = [ocffm.shard_range(rows, nranks, r) for r in range(nranks)]
    assert spans[0][0] == 0 and spans[-1][1] == rows
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_shard_range_rejects_bad_ranks():
    with pytest.raises(ocffm.OcffmError):
        ocffm.shard_range(10, 2, 2)
    with pytest.raises(ocffm.OcffmError):
        ocffm.shard_range(10, 0, 0)


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import dist_util
    import torch.distributed as dist
    dist_util.init("gloo")
    uid = dist_util.share_unique_id(lambda: bytes(range(128)))
    slowest = dist_util.max_over_ranks(10.0 + rank)
    total = dist_util.sum_over_ranks(100.0 * (rank + 1))      # per-rank counters -> job totals (bench.py)
    dist_util.barrier()
    lo, hi = ocffm.shard_range(1001, world, rank)
    out.put((rank, uid, slowest, lo, hi, total))
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] == bytes(range(128)) for r in res)      # every rank received rank 0's id
    assert all(r[2] == 11.0 for r in res)                   # max over ranks
    assert all(r[5] == 300.0 for r in res)                  # sum over ranks
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 500, 500, 1001)
