"""Two-GPU test of the fused Count + peer-memory all-reduce (needs >= 2 CUDA devices; skipped otherwise).
One process per GPU; mailbox handles are exchanged over gloo; every rank must see the global count, including
when a rank has no shards at all (it still takes part in the exchange) and for the NCCL path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from featurebase_b200 import cluster, datagen as D, executor as X, lib as L
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    h = X.Holder(device=rank)
    idx = h.create_index("i", track_existence=False)
    f = idx.create_field("f")
    n_shards = 12
    lo, hi = cluster.shard_range(rank, world, n_shards)
    mine = np.arange(lo, hi, dtype=np.uint64)
    g = idx.create_field("g")
    for s in mine:
        h.import_roaring("i", "f", X.VIEW_STANDARD, int(s), D.fragment(1, int(s), [0, 1, 2, 3], 0.02))
        h.import_roaring("i", "g", X.VIEW_STANDARD, int(s), D.fragment(2, int(s), [0, 1, 2], 0.05))
    row = lambda r: L.Op(L.OP_ROW, f.id, 0, 0, r, 0, 0, 0)
    pair = [row(0), row(1), L.Op(L.OP_INTERSECT, 0, 0, 2, 0, 0, 0, 0)]
    union = [row(0), row(1), row(2), row(3), L.Op(L.OP_UNION, 0, 0, 4, 0, 0, 0, 0)]
    out = {}
    # 1. NCCL path
    uid = [h.ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    h.ctx.comm_init(world, rank, uid[0])
    out["nccl_pair"] = h.ctx.count(idx.id, pair, mine)
    out["nccl_union"] = h.ctx.count(idx.id, union, mine)
    # count VECTORS are summed by the same ncclAllReduce(uint64, sum): TopN / TopK over explicit ids, GroupBy (BASELINE config 4's merge)
    out["nccl_row_counts"] = [int(x) for x in h.ctx.row_counts(idx.id, f.id, 0, mine, row_ids=[0, 1, 2, 3, 9])]
    out["nccl_row_counts_filtered"] = [int(x) for x in h.ctx.row_counts(idx.id, f.id, 0, mine, row_ids=[0, 1, 2, 3], filter_ops=[L.Op(L.OP_ROW, g.id, 0, 0, 1, 0, 0, 0)])]
    out["nccl_groupby"] = [int(x) for x in np.asarray(h.ctx.groupby(idx.id, [f.id, g.id], [0, 0], [[0, 1, 2, 3], [0, 1, 2]], mine)).reshape(-1)]
    # 2. fused peer-memory path
    handles = [None] * world
    dist.all_gather_object(handles, h.ctx.comm_p2p_handle())
    h.ctx.comm_p2p_open(world, rank, handles)
    for it in range(5):     # several epochs: exercises the parity double-buffering
        out[f"p2p_pair_{it}"] = h.ctx.count(idx.id, pair, mine)
        out[f"p2p_union_{it}"] = h.ctx.count(idx.id, union, mine)
    # rank 1 contributes nothing in this query (no kernel work, exchange only)
    out["p2p_rank0_only"] = h.ctx.count(idx.id, pair, mine if rank == 0 else np.zeros(0, dtype=np.uint64))
    local_pair = None
    if rank == 0:
        h.ctx2 = None
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_count_allreduce_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from featurebase_b200 import cluster, datagen as D
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pair_tot, union_tot, pair_r0 = 0, 0, 0
    rc, rcf, gb = np.zeros(5, dtype=np.uint64), np.zeros(4, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    for s in range(12):
        fr = O.Bitmap.from_bytes(D.fragment(1, s, [0, 1, 2, 3], 0.02))
        gr = O.Bitmap.from_bytes(D.fragment(2, s, [0, 1, 2], 0.05))
        r = [fr.row(k, s) for k in range(4)]
        for k in range(4):
            rc[k] += np.uint64(r[k].count())
            rcf[k] += np.uint64(r[k].intersection_count(gr.row(1, s)))
        gb += O.groupby_shard([fr, gr], s, [[0, 1, 2, 3], [0, 1, 2]])
        c = r[0].intersection_count(r[1])
        pair_tot += c
        if s < cluster.shard_range(0, 2, 12)[1]:
            pair_r0 += c
        union_tot += r[0].union(r[1], r[2], r[3]).count()
    for rank in (0, 1):
        o = res[rank]
        assert o["nccl_pair"] == pair_tot and o["nccl_union"] == union_tot
        assert o["nccl_row_counts"] == [int(x) for x in rc] and o["nccl_row_counts_filtered"] == [int(x) for x in rcf]
        assert o["nccl_groupby"] == [int(x) for x in gb]
        for it in range(5):
            assert o[f"p2p_pair_{it}"] == pair_tot and o[f"p2p_union_{it}"] == union_tot
        assert o["p2p_rank0_only"] == pair_r0
