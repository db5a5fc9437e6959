"""fbgpu_node: every GPU of one process behind one handle (include/fbgpu.h).  One C call fans a query out to the devices that
own the listed shards and merges on the host; any number of caller threads may query concurrently (FeatureBase's goroutines,
executor.go:6449-6533, 6742-6812).  With fewer than two CUDA devices the node is built from two contexts on device 0: routing,
fan-out, merge and the concurrency of the host side are the same code.  Every answer is compared with the CPU oracle driven
through the very same calls (tests/oracle_ctx.py)."""
import os
import threading

import numpy as np
import pytest

from featurebase_b200 import datagen as D
from featurebase_b200 import lib as L
from tests.oracle_ctx import OracleCtx

pytestmark = pytest.mark.gpu
ON_EMU = bool(os.environ.get("FBGPU_TEST_ON_EMULATOR"))


def _devices():
    if ON_EMU:
        return [0, 0]
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n >= 2 else [0, 0]


N_SHARDS = 6 if ON_EMU else 24
BLOCK = 2 if ON_EMU else 3          # shard s -> device slot (s // BLOCK) % n_devices
F, G, V = 1, 2, 3                   # field ids: two set fields, one int field (view ids 0 / 0 / 7)


def _row(field, r, view=0):
    return L.Op(L.OP_ROW, field, view, 0, r, 0, 0, 0)


def _nary(op, n):
    return L.Op(op, 0, 0, n, 0, 0, 0, 0)


@pytest.fixture(scope="module")
def world():
    node, ora = L.Node(_devices(), BLOCK), OracleCtx()
    shards = np.arange(N_SHARDS, dtype=np.uint64)
    for s in shards:
        for fld, seed, rows, p in ((F, 11, list(range(6)), 0.02), (G, 12, list(range(4)), 0.08)):
            data = D.fragment(seed, int(s), rows, p, mode=int(s) % 2)       # odd shards: clustered generator (run containers)
            node.load_fragment(0, fld, 0, int(s), data)
            ora.load_fragment(0, fld, 0, int(s), data)
        data = D.bsi_fragment(13, int(s), 1 << 16, 12, -2000, 2000)
        node.load_fragment(0, V, 7, int(s), data)
        ora.load_fragment(0, V, 7, int(s), data)
    node.commit()
    yield node, ora, shards
    node.close()


def _queries(shards):
    some = shards[1::2]
    q = []
    q.append(("count", lambda c: c.count(0, [_row(F, 0), _row(F, 1), _nary(L.OP_INTERSECT, 2)], shards)))
    q.append(("count_union", lambda c: c.count(0, [_row(F, 0), _row(F, 1), _row(F, 2), _nary(L.OP_UNION, 3), _row(G, 1), _nary(L.OP_INTERSECT, 2)], shards)))
    q.append(("count_some", lambda c: c.count(0, [_row(F, 3), _row(G, 0), _nary(L.OP_DIFFERENCE, 2)], some)))
    q.append(("count_per_shard", lambda c: [int(x) for x in c.count(0, [_row(G, 2), _row(F, 4), _nary(L.OP_XOR, 2)], shards[::-1], per_shard=True)[1]]))
    q.append(("topn", lambda c: [int(x) for x in c.row_counts(0, F, 0, shards, row_ids=[5, 0, 3, 9])]))
    q.append(("topn_filtered", lambda c: [int(x) for x in c.row_counts(0, F, 0, shards, row_ids=[0, 1, 2, 3, 4, 5], filter_ops=[_row(G, 1)])]))
    q.append(("pairs", lambda c: [int(x) for x in c.count_pairs(0, F, 0, [0, 1, 2], G, 0, [0, 1, 2], shards)]))
    q.append(("groupby", lambda c: np.asarray(c.groupby(0, [F, G], [0, 0], [list(range(6)), list(range(4))], shards)).reshape(-1).tolist()))
    q.append(("groupby_filtered", lambda c: np.asarray(c.groupby(0, [F, G], [0, 0], [[0, 2, 4], [1, 3]], some, filter_ops=[_row(F, 1)])).reshape(-1).tolist()))
    q.append(("bsi_count", lambda c: c.count(0, [L.Op(L.OP_BSI_RANGE, V, 7, 0, 12, L.CMP[">"], 100, 0)], shards)))
    q.append(("any", lambda c: (c.any(0, [_row(F, 5)], shards), c.any(0, [_row(F, 77)], shards), c.any(0, [_row(F, 0), _row(G, 0), _nary(L.OP_INTERSECT, 2)], some))))
    q.append(("row", lambda c: c.row(0, [_row(F, 0), _row(G, 3), _nary(L.OP_UNION, 2)], shards)))
    q.append(("row_some", lambda c: c.row(0, [_row(F, 2), _row(F, 3), _nary(L.OP_INTERSECT, 2)], some)))
    return q


def test_every_call_matches_the_oracle(world):
    node, ora, shards = world
    assert node.n_devices >= 2
    owners = {node.owner(int(s)) for s in shards}
    assert len(owners) == node.n_devices                     # every device slot holds some of the shards
    for name, q in _queries(shards):
        assert q(node) == q(ora), name
    st = node.stats()
    assert st["fragments"] == 3 * N_SHARDS


def test_concurrent_callers_never_mix_results(world):
    """8 threads x mixed Count / TopN / GroupBy / Row queries, each in its own order: every single answer equals the oracle's"""
    node, ora, shards = world
    qs = _queries(shards)
    want = [q(ora) for _, q in qs]
    errors = []
    rounds = 1 if ON_EMU else 6

    def worker(t):
        rng = np.random.default_rng(t)
        try:
            for _ in range(rounds):
                for k in rng.permutation(len(qs)):
                    got = qs[k][1](node)
                    if got != want[k]:
                        errors.append((t, qs[k][0]))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(2 if ON_EMU else 8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors[:5]


def test_a_failing_device_fails_only_that_call(world):
    node, ora, shards = world
    with pytest.raises(L.FbgpuError) as e:
        node.count(0, [_nary(L.OP_INTERSECT, 0)], shards)            # Intersect() without children: the reference's error (executor.go:5362)
    assert e.value.code == L.E_QUERY
    with pytest.raises(L.FbgpuError):
        node.count(0, [_nary(L.OP_INTERSECT, 0)], shards[:0])       # ... also when no shard is listed
    name, q = _queries(shards)[0]
    assert q(node) == q(ora)                                       # the node keeps answering


@pytest.mark.skipif(ON_EMU, reason="the interpreted kernels run a launch to completion on the calling thread: no concurrent peer to wait for")
def test_fused_exchange_wait_is_bounded(monkeypatch):
    """the in-kernel wait for a peer's count ends after FBGPU_P2P_TIMEOUT_MS with FBGPU_E_COMM instead of hanging the GPU"""
    monkeypatch.setenv("FBGPU_P2P_TIMEOUT_MS", "150")
    a, b = L.Context(0), L.Context(0)
    try:
        data = D.fragment(21, 0, [0, 1], 0.01)
        for c in (a, b):
            c.load_fragment(0, F, 0, 0, data)
            c.commit()
        prog = [_row(F, 0), _row(F, 1), _nary(L.OP_INTERSECT, 2)]
        alone = a.count(0, prog, [0])
        L.p2p_open_local([a, b])
        out = {}
        tb = threading.Thread(target=lambda: out.setdefault("b", b.count(0, prog, [0])))
        tb.start()
        out["a"] = a.count(0, prog, [0])
        tb.join()
        assert out["a"] == out["b"] == 2 * alone                    # both ranks took part: the sum
        with pytest.raises(L.FbgpuError) as e:                      # rank 1 never shows up for the second exchange
            a.count(0, prog, [0])
        assert e.value.code == L.E_COMM
    finally:
        a.close()
        b.close()
