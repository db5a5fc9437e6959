"""N>1 host logic on CPU (gloo, world_size 2): contiguous shard-range partition, u64-sum all-reduce of the
count buffers (Count / per-row counts / GroupBy tensor) and disjoint Row merge.  The per-rank 'executor' here is
the CPU oracle standing in for a GPU context — the collective and partition code under test is the product's."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partition():
    from featurebase_b200 import cluster
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1024, 4096, 4099):
            seen = []
            for r in range(world):
                lo, hi = cluster.shard_range(r, world, n)
                seen.extend(range(lo, hi))
                for s in (lo, hi - 1):
                    if lo < hi:
                        assert cluster.owner_of(s, world, n) == r
            assert seen == list(range(n))
            sizes = [cluster.shard_range(r, world, n)[1] - cluster.shard_range(r, world, n)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from featurebase_b200 import cluster, datagen as D, roaring_io
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_shards = 6
    mine = cluster.local_shards(np.arange(n_shards), rank, world, n_shards)
    total, per_row = 0, np.zeros(3, dtype=np.uint64)
    inter = O.Bitmap()
    gb = np.zeros((4, 4), dtype=np.uint64)
    for s in mine:
        fr = O.Bitmap.from_bytes(D.fragment(1, int(s), [0, 1, 2], 0.02))
        a, b = fr.row(0, int(s)), fr.row(1, int(s))
        x = a.intersect(b)
        total += x.count()
        inter = inter.union(x)
        for r in range(3):
            per_row[r] += fr.row(r, int(s)).count()
        fa, fb = D.groupby_fragments(1, 2, int(s), 0.01, 4, 4)
        O.groupby_shard([O.Bitmap.from_bytes(fa), O.Bitmap.from_bytes(fb)], int(s), [list(range(4)), list(range(4))], None, gb.reshape(-1))
    red = cluster.all_reduce_counts(np.concatenate([[total], per_row, gb.reshape(-1)]))
    rows = [None] * world
    dist.all_gather_object(rows, inter.to_bytes())
    merged = cluster.merge_rows(rows)
    if rank == 0:
        q.put((red.tolist(), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_count_reduce_and_row_merge():
    sys.path.insert(0, ROOT)
    from featurebase_b200 import datagen as D
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process ground truth over all 6 shards
    total, per_row = 0, np.zeros(3, dtype=np.uint64)
    inter = O.Bitmap()
    gb = np.zeros(16, dtype=np.uint64)
    for s in range(6):
        fr = O.Bitmap.from_bytes(D.fragment(1, s, [0, 1, 2], 0.02))
        x = fr.row(0, s).intersect(fr.row(1, s))
        total += x.count()
        inter = inter.union(x)
        for r in range(3):
            per_row[r] += fr.row(r, s).count()
        fa, fb = D.groupby_fragments(1, 2, s, 0.01, 4, 4)
        O.groupby_shard([O.Bitmap.from_bytes(fa), O.Bitmap.from_bytes(fb)], s, [list(range(4)), list(range(4))], None, gb)
    assert red == [total] + per_row.tolist() + gb.tolist()
    assert merged == inter.to_bytes()
