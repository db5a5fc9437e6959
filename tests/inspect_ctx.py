"""InspectCtx: the product's host side end to end on a box without a GPU.  Fragments are loaded into an inspection-only
library context (fbgpu_init(FBGPU_DEVICE_NONE)); Count / Row queries are answered by (1) the library's own program
compiler (fbgpu_debug_compile), (2) containers located by the library's own resolve() over its own store tables
(fbgpu_debug_container), and (3) a Python model of the kernels' stack machine on 2^16-bit units.  Only step 3 stands in
for CUDA code.  TEST INFRASTRUCTURE; row_counts / count_pairs / groupby fall back to the oracle-backed context."""
import numpy as np

from featurebase_b200 import lib as L
from oracle import oracle as O
from tests.oracle_ctx import OracleCtx

(D_PUSH_ROW, D_PUSH_EMPTY, D_OR_ROW, D_AND_ROW, D_ANDNOT_ROW, D_XOR_ROW, D_ORAND_ROW, D_ORANDNOT_ROW, D_AND, D_OR, D_ANDNOT, D_XOR,
 D_SWAP, D_POP) = range(1, 15)
NO_VIEW = 0xFFFFFFFF


class InspectCtx(OracleCtx):
    def __init__(self):
        super().__init__()
        self.lib = L.Context(L.DEVICE_NONE)

    def load_fragment(self, index, field, view, shard, data):
        super().load_fragment(index, field, view, shard, data)
        self.lib.load_fragment(index, field, view, shard, data)

    def load_rbf(self, index, shard, data, names, fields, views, wal=b""):
        super().load_rbf(index, shard, data, names, fields, views, wal)
        return self.lib.load_rbf(index, shard, data, names, fields, views, wal)

    def _unit(self, fv, shard, row, slot):
        """65536 bools of one (shard, slot) stripe of a row, from the library's store"""
        out = np.zeros(65536, dtype=bool)
        if fv == NO_VIEW:
            return out
        found = self.lib.debug_container(0xFFFFFFFF, fv, 0, shard, row, slot)
        if found is None:
            return out
        typ, card, runs, payload = found
        if typ == 1:
            out[np.frombuffer(payload, dtype="<u2")[:card]] = True
        elif typ == 2:
            out[:] = np.unpackbits(np.frombuffer(payload, dtype=np.uint8), bitorder="little").astype(bool)
        else:
            for s, l in np.frombuffer(payload, dtype="<u2")[: 2 * runs].reshape(-1, 2).tolist():
                out[s:l + 1] = True
        assert int(out.sum()) == card
        return out

    def _run(self, prog, depth, shard, slot):
        st = []
        for op, fv, row in prog:
            x = self._unit(fv, shard, row, slot) if op in (D_PUSH_ROW, D_OR_ROW, D_AND_ROW, D_ANDNOT_ROW, D_XOR_ROW, D_ORAND_ROW, D_ORANDNOT_ROW) else None
            if op == D_PUSH_ROW:
                st.append(x)
            elif op == D_PUSH_EMPTY:
                st.append(np.zeros(65536, dtype=bool))
            elif op == D_OR_ROW:
                st[-1] = st[-1] | x
            elif op == D_AND_ROW:
                st[-1] = st[-1] & x
            elif op == D_ANDNOT_ROW:
                st[-1] = st[-1] & ~x
            elif op == D_XOR_ROW:
                st[-1] = st[-1] ^ x
            elif op == D_ORAND_ROW:
                st[-2] = st[-2] | (st[-1] & x)
            elif op == D_ORANDNOT_ROW:
                st[-2] = st[-2] | (st[-1] & ~x)
            elif op == D_SWAP:
                st[-1], st[-2] = st[-2], st[-1]
            elif op == D_POP:
                st.pop()
            else:
                b = st.pop()
                st[-1] = {D_AND: st[-1] & b, D_OR: st[-1] | b, D_ANDNOT: st[-1] & ~b, D_XOR: st[-1] ^ b}[op]
            assert len(st) <= depth
        return st[-1] if st else np.zeros(65536, dtype=bool)

    def count(self, index, ops, shards, per_shard=False):
        self.programs.append(list(ops))
        prog, depth = self.lib.debug_compile(index, ops)
        per = np.array([sum(int(self._run(prog, depth, int(s), slot).sum()) for slot in range(16)) for s in shards], dtype=np.uint64)
        exp = OracleCtx.count(self, index, ops, shards, per_shard=True)[1]
        assert np.array_equal(per, exp), "library compile + store + stack model disagrees with the oracle"
        self.programs.pop()
        return (int(per.sum()), per) if per_shard else int(per.sum())

    def row(self, index, ops, shards):
        self.programs.append(list(ops))
        prog, depth = self.lib.debug_compile(index, ops)
        vals = []
        for s in sorted(set(int(s) for s in shards)):
            for slot in range(16):
                bits = np.flatnonzero(self._run(prog, depth, s, slot)).astype(np.uint64)
                vals.append(bits + np.uint64((s * 16 + slot) << 16))
        out = O.Bitmap.from_values(np.concatenate(vals) if vals else np.zeros(0, dtype=np.uint64))
        return out.to_bytes(), out.count()
