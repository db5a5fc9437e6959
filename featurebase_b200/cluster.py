"""Shard placement and result reduction across the GPUs of one box.

The reference places shards with fnv64a(index||shard) % 256 -> jump hash over nodes (disco/snapshot.go:69-78,
disco/hasher.go:16-24) and merges per-node results over HTTP (executor.go:6449-6533).  Here GPU g owns the
contiguous shard range [g*S/G, (g+1)*S/G) and the only exchange is one sum all-reduce of the count buffer:
inside libfbgpu (NCCL, fbgpu_comm_init) on GPUs, or through torch.distributed (gloo) in the CPU tests."""
import numpy as np


def shard_range(rank, world, n_shards):
    """contiguous range owned by `rank` (sizes differ by at most one)"""
    lo = n_shards * rank // world
    hi = n_shards * (rank + 1) // world
    return lo, hi


def owner_of(shard, world, n_shards):
    """inverse of shard_range"""
    r = min(world - 1, (int(shard) * world + world - 1) // max(n_shards, 1))
    while r > 0 and shard_range(r, world, n_shards)[0] > shard:
        r -= 1
    while shard_range(r, world, n_shards)[1] <= shard:
        r += 1
    return r


def local_shards(shards, rank, world, n_shards):
    shards = np.asarray(shards, dtype=np.uint64)
    lo, hi = shard_range(rank, world, n_shards)
    return shards[(shards >= lo) & (shards < hi)]


def all_reduce_counts(counts, group=None):
    """reduce = u64 add (executor.go:5880-5883, cache.go:464, executor.go:3728): torch.distributed sum over int64"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(counts, dtype=np.uint64)).view(np.int64).copy())
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy().view(np.uint64)


def merge_rows(row_bytes_per_rank):
    """Row results need no collective: segments are disjoint by shard (Row.Merge row.go:202).  Concatenates the
    per-rank Pilosa-roaring buffers (ascending shard ranges) into one."""
    from . import roaring_io
    conts = []
    for data in row_bytes_per_rank:
        conts.extend(list(roaring_io.containers(data)))
    conts.sort(key=lambda c: c[0])
    out = bytearray()
    out += np.array([roaring_io.MAGIC], dtype="<u4").tobytes() + np.array([len(conts)], dtype="<u4").tobytes()
    for key, typ, n, _ in conts:
        out += np.array([key], dtype="<u8").tobytes() + np.array([typ, n - 1], dtype="<u2").tobytes()
    off = 8 + 16 * len(conts)
    sizes = []
    for _, typ, _, payload in conts:
        size = len(payload) + (2 if typ == roaring_io.RUN else 0)
        out += np.array([off & 0xFFFFFFFF], dtype="<u4").tobytes()
        off += size
        sizes.append(size)
    for _, typ, _, payload in conts:
        if typ == roaring_io.RUN:
            out += np.array([len(payload) // 4], dtype="<u2").tobytes()
        out += payload.tobytes()
    return bytes(out)
