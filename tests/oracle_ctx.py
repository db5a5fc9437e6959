"""OracleCtx: a stand-in for featurebase_b200.lib.Context that interprets the same post-order fbgpu_op programs with
the CPU oracle.  TEST INFRASTRUCTURE ONLY (`-m "not gpu"` host-logic tests): it lets the host mirror
(featurebase_b200/executor.py: PQL -> program, baseValue clamping, TopN/TopK/GroupBy reduction) run here, where there
is no GPU, through exactly the calls it makes on the C ABI.  The opcode semantics follow include/fbgpu.h's table."""
import numpy as np

from featurebase_b200 import lib as L
from oracle import oracle as O

_CMP = {v: k for k, v in L.CMP.items()}


class OracleCtx:
    def __init__(self):
        self.frags = {}                  # (index, field, view) -> {shard: oracle Bitmap}
        self.programs = []               # every program handed to count()/row(), for assertions on the compile step

    # ---- residency
    def load_fragment(self, index, field, view, shard, data):
        slot = self.frags.setdefault((index, field, view), {})
        if int(shard) in slot:                               # (a replaced fragment leaves its bytes behind, as in the library's arena)
            self._dead = getattr(self, "_dead", 0) + 2 * len(bytes(data))
        slot[int(shard)] = O.Bitmap.from_bytes(bytes(data))

    def apply_containers(self, index, field, view, shard, data=b"", removed_keys=()):
        slot = self.frags.setdefault((index, field, view), {})
        cur = slot.get(int(shard))
        vals = cur.slice() if cur is not None else np.zeros(0, dtype=np.uint64)
        put = O.Bitmap.from_bytes(bytes(data)).slice() if data else np.zeros(0, dtype=np.uint64)
        gone = np.concatenate([np.unique(put >> np.uint64(16)), np.asarray(list(removed_keys), dtype=np.uint64)])
        if len(set(np.unique(put >> np.uint64(16)).tolist()) & set(int(k) for k in removed_keys)):
            raise L.FbgpuError(L.E_INVALID, "container key both written and removed")
        self._dead = getattr(self, "_dead", 0) + 1
        vals = np.sort(np.concatenate([vals[~np.isin(vals >> np.uint64(16), gone)], put]))
        if len(vals):
            slot[int(shard)] = O.Bitmap.from_values(vals)
        else:
            slot.pop(int(shard), None)

    def commit(self):
        pass

    def load_rbf(self, index, shard, data, names, fields, views, wal=b""):
        """same contract as lib.Context.load_rbf; the file is read by the product's rbf_reader.h through the g++ harness"""
        from tests import test_rbf as TR
        try:
            found = TR.dump(TR.harness(), bytes(data), bytes(wal or b""))
        except ValueError as e:
            raise L.FbgpuError(L.E_FORMAT, str(e))
        n = 0
        for name, field, view in zip(names, fields, views):
            if name not in found:
                continue
            vals = [TR.cell_values(*cell) for cell in found[name]]
            self.frags.setdefault((index, int(field), int(view)), {})[int(shard)] = O.Bitmap.from_values(np.concatenate(vals) if vals else [])
            n += 1
        return n

    def stats(self):
        import struct
        frs = [b for d in self.frags.values() for b in d.values()]
        payload = 0
        for b in frs:
            raw = b.to_bytes()
            n = struct.unpack_from("<I", raw, 4)[0]
            for i in range(n):
                _, typ, n1 = struct.unpack_from("<QHH", raw, 8 + 12 * i)
                off = struct.unpack_from("<I", raw, 8 + 12 * n + 4 * i)[0]
                payload += 2 * (n1 + 1) if typ == 1 else 8192 if typ == 2 else 4 * struct.unpack_from("<H", raw, off)[0]
        return {"fragments": len(frs), "containers": sum(struct.unpack_from("<I", b.to_bytes(), 4)[0] for b in frs), "payload_bytes": payload,
                "dead_bytes": getattr(self, "_dead", 0), "device_bytes": 0}

    def compact(self):
        self._dead = 0

    def _frag(self, index, field, view, shard):
        return self.frags.get((index, field, view), {}).get(int(shard))

    def _row(self, index, field, view, row, shard):
        f = self._frag(index, field, view, shard)
        return f.row(int(row), shard) if f is not None else O.Bitmap()

    # ---- program interpreter (one shard)
    def _eval(self, index, ops, shard):
        st = []
        for op in ops:
            code, argc = op.opcode, op.argc
            if code == L.OP_ROW or code == L.OP_ALL:
                st.append(self._row(index, op.field, op.view, op.a, shard))
            elif code == L.OP_EMPTY:
                st.append(O.Bitmap())
            elif code == L.OP_NOT:
                if argc != 1 or not st:
                    raise L.FbgpuError(L.E_INVALID, "malformed program")
                st.append(self._row(index, op.field, op.view, op.a, shard).difference(st.pop()))
            elif code == L.OP_BSI_RANGE:
                f = self._frag(index, op.field, op.view, shard)
                st.append(f.range_op(_CMP[op.b], int(op.a), int(op.lo), int(op.hi), shard=shard) if f is not None else O.Bitmap())
            elif code in (L.OP_INTERSECT, L.OP_UNION, L.OP_DIFFERENCE, L.OP_XOR):
                if argc == 0:
                    if code == L.OP_INTERSECT:
                        raise L.FbgpuError(L.E_QUERY, "empty Intersect query is currently not supported")
                    if code == L.OP_DIFFERENCE:
                        raise L.FbgpuError(L.E_QUERY, "empty Difference query is currently not supported")
                    st.append(O.Bitmap())
                    continue
                if argc > len(st):
                    raise L.FbgpuError(L.E_INVALID, "malformed program")
                args = st[len(st) - argc:]
                del st[len(st) - argc:]
                out = args[0]
                for a in args[1:]:
                    out = {L.OP_INTERSECT: out.intersect, L.OP_UNION: out.union, L.OP_DIFFERENCE: out.difference,
                           L.OP_XOR: out.xor}[code](a)
                st.append(out)
            else:
                raise L.FbgpuError(L.E_INVALID, "unknown opcode")
        if len(st) != 1:
            raise L.FbgpuError(L.E_INVALID, "malformed program")
        return st[0]

    # ---- queries (signatures of lib.Context)
    def count(self, index, ops, shards, per_shard=False):
        self.programs.append(list(ops))
        per = np.array([self._eval(index, ops, s).count() for s in shards], dtype=np.uint64)
        return (int(per.sum()), per) if per_shard else int(per.sum())

    def any(self, index, ops, shards):
        if len(shards) == 0:
            self._eval(index, ops, 0)                       # the program is still validated
            return False
        return any(self._eval(index, ops, s).count() > 0 for s in shards)

    def row(self, index, ops, shards):
        self.programs.append(list(ops))
        out = O.Bitmap()
        for s in sorted(int(s) for s in shards):
            out = out.union(self._eval(index, ops, s))
        return out.to_bytes(), out.count()


    def columns(self, index, ops, shards, offset=0, limit=None):
        from featurebase_b200 import roaring_io
        data, n = self.row(index, ops, shards)
        cols = np.asarray(roaring_io.decode(data), dtype=np.uint64)
        return (cols[offset:] if limit is None else cols[offset:offset + limit]), n

    def extract(self, index, field, view, bit_depth, shards, filter_ops=None, offset=0, limit=None):
        from featurebase_b200 import roaring_io
        cols, vals = [], []
        for s in sorted(set(int(x) for x in shards)):
            f = self._frag(index, field, view, s)
            if f is None:
                continue
            plane = lambda r: np.asarray(roaring_io.decode(f.row(r, s).to_bytes()), dtype=np.uint64)
            keep = f.row(0, s)
            if filter_ops:
                keep = keep.intersect(self._eval(index, filter_ops, s))
            c = np.asarray(roaring_io.decode(keep.to_bytes()), dtype=np.uint64)
            v = np.zeros(len(c), dtype=np.int64)
            for b in range(bit_depth):
                v |= np.isin(c, plane(2 + b)).astype(np.int64) << b
            v = np.where(np.isin(c, plane(1)), -v, v)
            cols.append(c)
            vals.append(v)
        cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.uint64)
        vals = np.concatenate(vals) if vals else np.zeros(0, dtype=np.int64)
        end = None if limit is None else offset + limit
        return cols[offset:end], vals[offset:end], len(cols)

    def bsi_sum(self, index, field, view, bit_depth, shards, filter_ops=None):
        _, vals, _ = self.extract(index, field, view, bit_depth, shards, filter_ops=filter_ops)
        t = int(vals.astype(object).sum()) if len(vals) else 0
        t &= (1 << 64) - 1
        return (t - (1 << 64) if t >> 63 else t), len(vals)

    def bsi_minmax(self, index, field, view, bit_depth, shards, want_max, filter_ops=None):
        _, vals, _ = self.extract(index, field, view, bit_depth, shards, filter_ops=filter_ops)
        if len(vals) == 0:
            return 0, 0
        v = int(vals.max() if want_max else vals.min())
        return v, int((vals == v).sum())

    def row_counts(self, index, field, view, shards, row_ids=None, filter_ops=None, cap=1 << 20):
        tot = {}
        for s in shards:
            f = self._frag(index, field, view, s)
            if f is None:
                continue
            filt = self._eval(index, filter_ops, s) if filter_ops else None
            rows, cnts = f.row_counts(s, filt)
            for r, c in zip(rows.tolist(), cnts.tolist()):
                tot[r] = tot.get(r, 0) + c
        if row_ids is not None:
            return np.array([tot.get(int(r), 0) for r in row_ids], dtype=np.uint64)
        pairs = sorted(((r, c) for r, c in tot.items() if c), key=lambda kv: (-kv[1], kv[0]))      # cap sizes buffers only: the list is never truncated
        return (np.array([p[0] for p in pairs], dtype=np.uint64), np.array([p[1] for p in pairs], dtype=np.uint64))

    def row_counts_per_shard(self, index, field, view, shards, row_ids, filter_ops=None):
        return np.stack([self.row_counts(index, field, view, [s], row_ids=row_ids, filter_ops=filter_ops) for s in shards]) if len(shards) else np.zeros((0, len(row_ids)), dtype=np.uint64)

    def count_pairs(self, index, field_a, view_a, rows_a, field_b, view_b, rows_b, shards):
        out = np.zeros(len(rows_a), dtype=np.uint64)
        for i, (ra, rb) in enumerate(zip(rows_a, rows_b)):
            for s in shards:
                out[i] += self._row(index, field_a, view_a, ra, s).intersection_count(self._row(index, field_b, view_b, rb, s))
        return out

    def groupby(self, index, fields, views, row_ids, shards, filter_ops=None):
        shape = [len(r) for r in row_ids]
        out = np.zeros(int(np.prod(shape)), dtype=np.uint64)
        for s in shards:
            frags = [self._frag(index, f, v, s) for f, v in zip(fields, views)]
            filt = self._eval(index, filter_ops, s) if filter_ops else None
            O.groupby_shard(frags, s, [list(map(int, r)) for r in row_ids], filt, out)
        return out.reshape(shape)
