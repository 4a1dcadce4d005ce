"""The N>1 path on CPU: two gloo ranks each own a shard of the pair set (no data-path collective),
and the counters / timing reduce exactly as bench.py reports them."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from genarchbench_b200 import dist as bdist
from genarchbench_b200 import pairio


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_split_by_cells_balances_work():
    b = pairio.generate(4, 20000, seed=12)
    cuts = bdist.split_by_cells(b.pairs["len1"], b.pairs["len2"], 8)
    assert cuts[0] == 0 and cuts[-1] == len(b) and (np.diff(cuts) >= 0).all()
    w = b.pairs["len1"].astype(np.int64) * b.pairs["len2"]
    loads = np.array([w[cuts[i]:cuts[i + 1]].sum() for i in range(8)])
    assert loads.max() / loads.mean() < 1.05


def _worker(rank, world, port, q):
    import oracle
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = bdist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    # strong split of one common batch: every rank computes only its contiguous slice
    full = pairio.generate(1, 4001, seed=77)
    lo, hi = bdist.shard_range(len(full), rank, world)
    mine = full.slice(lo, hi)
    cells = oracle.oracle_batch(mine, nthreads=1)       # CPU stand-in for the device step
    checksum = int(mine.outputs().astype(np.int64).sum())
    bdist.barrier()
    sums, maxes = bdist.reduce_stats([len(mine), cells, checksum], [10.0 + rank])
    q.put((rank, sums, maxes, bdist.shard_seed(1003, rank)))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_rank_gloo_shards_and_reduces():
    import oracle
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = pairio.generate(1, 4001, seed=77)
    cells = oracle.oracle_batch(full, nthreads=1)
    checksum = int(full.outputs().astype(np.int64).sum())
    for rank, sums, maxes, seed in got:
        assert sums == [4001.0, float(cells), float(checksum)]   # every pair processed exactly once
        assert maxes == [11.0]                                    # max over ranks, not a sum
    assert len({g[3] for g in got}) == 2                          # weak-scaling shards use distinct seeds
