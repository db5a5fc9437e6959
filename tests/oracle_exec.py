"""Evaluates pql.Call trees with the CPU oracle, shard by shard, following the reference's executor semantics
(executor.go:1782-1816, 5120-5602, 5839-5892).  TEST INFRASTRUCTURE: the checker for the GPU path."""
import numpy as np

from featurebase_b200 import executor as X
from featurebase_b200 import pql
from oracle import oracle as O


class OracleIndex:
    """fragments[(field name, view)][shard] = oracle Bitmap (fragment-relative keys)"""

    def __init__(self, idx):
        self.idx = idx
        self.frags = {}

    def load(self, field, view, shard, data):
        self.frags.setdefault((field, view), {})[int(shard)] = O.Bitmap.from_bytes(data)

    def frag(self, field, view, shard):
        return self.frags.get((field, view), {}).get(int(shard))

    def row(self, field, view, row, shard):
        f = self.frag(field, view, shard)
        return f.row(row, shard) if f is not None else O.Bitmap()

    def eval_shard(self, c, shard):
        n = c.name
        if n in ("Row", "Range"):
            key = [k for k in c.args if not k.startswith("_") and k not in ("from", "to")][0]
            fld = self.idx.fields[key]
            v = c.args[key]
            if fld.type == "int" or isinstance(v, pql.Condition):
                return self._bsi(fld, v if isinstance(v, pql.Condition) else pql.Condition("==", v), shard)
            if "from" in c.args or "to" in c.args:                # executeRowShard :5209-5241: union over the covering time views
                from featurebase_b200 import timeq
                names = fld.views_by_time_range(timeq.parse_time(c.args["from"]) if "from" in c.args else None,
                                                timeq.parse_time(c.args["to"]) if "to" in c.args else None)
                out = O.Bitmap()
                for name in names:
                    if fld.view_id(name) is not None:
                        out = out.union(self.row(key, fld.view_id(name), int(v), shard))
                return out
            return self.row(key, X.VIEW_STANDARD, int(v), shard)
        if n == "Intersect":
            if not c.children:
                raise X.QueryError("empty Intersect query is currently not supported")
            out = self.eval_shard(c.children[0], shard)
            for ch in c.children[1:]:
                out = out.intersect(self.eval_shard(ch, shard))
            return out
        if n == "Union":
            rows = [self.eval_shard(ch, shard) for ch in c.children]
            if not rows:
                return O.Bitmap()
            return rows[0] if len(rows) == 1 else rows[0].union(*rows[1:])
        if n == "Difference":
            if not c.children:
                raise X.QueryError("empty Difference query is currently not supported")
            out = self.eval_shard(c.children[0], shard)
            for ch in c.children[1:]:
                out = out.difference(self.eval_shard(ch, shard))
            return out
        if n == "Xor":
            out = O.Bitmap()
            for i, ch in enumerate(c.children):
                r = self.eval_shard(ch, shard)
                out = r if i == 0 else out.xor(r)
            return out
        if n == "Not":
            ex = self.row(X.EXISTENCE_FIELD, X.VIEW_STANDARD, 0, shard)
            return ex.difference(self.eval_shard(c.children[0], shard))
        if n == "All":
            return self.row(X.EXISTENCE_FIELD, X.VIEW_STANDARD, 0, shard)
        if n == "ConstRow":                                      # executeConstRowShard :5674-5694
            cols = [int(x) for x in c.args["columns"] if int(x) // X.SHARD_WIDTH == shard]
            out = O.Bitmap.from_values(cols)
            return out.intersect(self.row(X.EXISTENCE_FIELD, X.VIEW_STANDARD, 0, shard)) if self.idx.track_existence else out
        if n == "UnionRows":                                     # plain Rows(f) children only: every row of the shard's fragment (fragment.unionRows)
            out = O.Bitmap()
            for ch in c.children:
                assert ch.name == "Rows" and set(ch.args) <= {"_field", "field"}
                fr = self.frag(ch.args.get("_field", ch.args.get("field")), X.VIEW_STANDARD, shard)
                for r in (fr.rows().tolist() if fr is not None else []):
                    out = out.union(fr.row(r, shard))
            return out
        raise KeyError(n)

    def _bsi(self, fld, cond, shard):
        frag = self.frag(fld.name, X.VIEW_BSI, shard)
        op, value = cond.op, cond.value
        not_null = lambda: frag.row(0, shard) if frag is not None else O.Bitmap()
        if value is None:                       # getNullRowShard / getNonNullRowShard executor.go:5056-5118
            if op == "!=":
                return not_null()
            return self.row(X.EXISTENCE_FIELD, X.VIEW_STANDARD, 0, shard).difference(not_null())
        if frag is None:
            return O.Bitmap()
        if op == "><":
            lo, hi, oor = fld.base_value_between(int(value[0]), int(value[1]))
            if oor:
                return O.Bitmap()
            if value[0] <= fld.min and value[1] >= fld.max:
                return not_null()
            return frag.range_op("><", fld.bit_depth, lo, hi, shard=shard)
        value = int(value)
        bv, oor = fld.base_value(op, value)
        if oor and op != "!=":
            return O.Bitmap()
        if (op == "<" and value > fld.max) or (op == "<=" and value >= fld.max) or (op == ">" and value < fld.min) or (op == ">=" and value <= fld.min):
            return not_null()
        if oor and op == "!=":
            return not_null()
        return frag.range_op(op, fld.bit_depth, bv, shard=shard)

    def eval_row(self, c, shards):
        """merged Row over shards (Row.Merge row.go:202) -> oracle Bitmap with absolute keys"""
        out = O.Bitmap()
        for s in sorted(shards):
            out = out.union(self.eval_shard(c, s))
        return out

    def count(self, c, shards):
        return sum(self.eval_shard(c, s).count() for s in shards)

    # ---- BSI aggregates the way the reference runs them: per shard, then reduced (executor.go:1119-1300)
    def _agg_shards(self, c, shards, fn):
        name = c.args.get("field", c.args.get("_field"))
        fld = self.idx.fields[name]
        for s in shards:
            filt = self.eval_shard(c.children[0], s) if c.children else None
            frag = self.frag(name, X.VIEW_BSI, s)
            if frag is None:
                yield fld, 0, 0                                  # ValCount{} (executeSumCountShard :2192-2195, field.go:1579-1582)
                continue
            v, n = fn(frag, fld.bit_depth, filt, s)
            yield fld, v, n

    def sum(self, c, shards):
        val = cnt = 0
        for fld, v, n in self._agg_shards(c, shards, O.bsi_sum):
            val += v + n * fld.base                              # :2203-2206, reduced with ValCount.Add :8438
            cnt += n
        return (X._i64(val), cnt) if cnt else (0, 0)

    def minmax(self, c, shards, want_max):
        best = None                                              # ValCount.Smaller / Larger :8446-8470,8526-8550
        for fld, v, n in self._agg_shards(c, shards, O.bsi_max if want_max else O.bsi_min):
            cur = (v + fld.base, n) if n else (0, 0)
            if best is None or best[1] == 0 or (cur[1] > 0 and (cur[0] > best[0] if want_max else cur[0] < best[0])):
                best = cur
            elif cur[0] == best[0]:
                best = (best[0], best[1] + cur[1])
        return best if best and best[1] else (0, 0)


class Pair:
    """a GPU Holder/Executor and an oracle index loaded with the same fragments"""

    def __init__(self, name="i", track_existence=True, ctx=None):
        self.holder = X.Holder(ctx=ctx)
        self.idx = self.holder.create_index(name, track_existence)
        self.ex = X.Executor(self.holder)
        self.ora = OracleIndex(self.idx)
        self.name = name

    def field(self, name, ftype="set", **kw):
        return self.idx.create_field(name, ftype, **kw)

    def load(self, field, view, shard, data):
        self.holder.import_roaring(self.name, field, view, shard, data)
        self.ora.load(field, view, shard, data)

    def apply(self, field, view, shard, put_values=(), removed_keys=()):
        """incremental refresh on both sides: put_values = fragment-relative bit positions of the WRITTEN containers (whole containers),
        removed_keys = deleted container keys; the product gets only the delta, the oracle the resulting fragment"""
        put = np.sort(np.asarray(put_values, dtype=np.uint64))
        data = O.Bitmap.from_values(put).to_bytes() if len(put) else b""
        self.holder.apply_containers(self.name, field, view, shard, data, list(removed_keys))
        cur = self.ora.frag(field, view, shard)
        vals = cur.slice() if cur is not None else np.zeros(0, dtype=np.uint64)
        gone = np.concatenate([np.unique(put >> np.uint64(16)), np.asarray(list(removed_keys), dtype=np.uint64)])
        vals = np.sort(np.concatenate([vals[~np.isin(vals >> np.uint64(16), gone)], put]))
        self.ora.frags.setdefault((field, view), {})[int(shard)] = O.Bitmap.from_values(vals)

    def sync_pending(self):
        """push Holder.set_bit/set_value staged bits to both sides"""
        from featurebase_b200 import roaring_io
        for (index, field, view, shard), bits in self.holder._pending.items():
            data = roaring_io.encode(np.fromiter(bits, dtype=np.uint64, count=len(bits)))
            self.load(field, view, shard, data)
        self.holder._pending = {}

    def shards(self):
        return sorted(self.idx.shards)

    def check_row(self, q, shards=None):
        shards = self.shards() if shards is None else shards
        call = pql.parse(q)[0]
        got = self.ex.execute(self.name, q, shards)[0]
        exp = self.ora.eval_row(call, shards)
        assert got.count == exp.count(), q
        assert got.roaring == exp.to_bytes(), q      # canonical bytes (Bitmap.WriteTo, roaring.go:1730)
        return got

    def check_count(self, q, shards=None):
        shards = self.shards() if shards is None else shards
        call = pql.parse(q)[0]
        got = self.ex.execute(self.name, q, shards)[0]
        exp = self.ora.count(call.children[0], shards)
        assert got == exp, (q, got, exp)
        return got
