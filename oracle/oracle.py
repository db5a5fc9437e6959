"""ctypes binding of the CPU oracle (oracle/fb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (featurebase_b200) never imports this module.

Parity pinning: the reference is Go and cannot run here; this oracle is pinned by the reference's own
golden vectors (tests/golden/*, restated from /root/reference test files) and cross-checked against
oracle/naive.py (an independent Python-set model, the roaring/naive.go idea).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ARRAY, BITMAP, RUN = 1, 2, 3
OP_EQ, OP_NEQ, OP_LT, OP_LTE, OP_GT, OP_GTE, OP_BETWEEN = 1, 2, 3, 4, 5, 6, 7
OPS = {"==": OP_EQ, "!=": OP_NEQ, "<": OP_LT, "<=": OP_LTE, ">": OP_GT, ">=": OP_GTE, "><": OP_BETWEEN}


def build(force=False):
    so = os.path.join(_HERE, "libfboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fb_oracle.c", "fb_bench.c", "fb_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.exists(x) and os.path.getmtime(so) < os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libfboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, u64, i64, i32, u16p = C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.POINTER(C.c_uint16)
        sig = {
            "fbo_c_array": (vp, [vp, i32]), "fbo_c_bitmap": (vp, [vp]), "fbo_c_run": (vp, [vp, i32]),
            "fbo_c_clone": (vp, [vp]), "fbo_c_free": (None, [vp]), "fbo_c_n": (i32, [vp]),
            "fbo_c_contains": (C.c_int, [vp, C.c_uint16]), "fbo_c_count_runs": (i32, [vp]),
            "fbo_c_optimize": (vp, [vp]), "fbo_c_convert": (vp, [vp, C.c_int]), "fbo_c_to_words": (None, [vp, vp]),
            "fbo_intersect": (vp, [vp, vp]), "fbo_union": (vp, [vp, vp]), "fbo_difference": (vp, [vp, vp]),
            "fbo_xor": (vp, [vp, vp]), "fbo_flip": (vp, [vp]), "fbo_intersection_count": (i32, [vp, vp]),
            "fbo_c_count_range": (i32, [vp, i32, i32]),
            "fbo_b_new": (vp, []), "fbo_b_free": (None, [vp]), "fbo_b_clone": (vp, [vp]),
            "fbo_b_put": (None, [vp, u64, vp]), "fbo_b_get": (vp, [vp, u64]), "fbo_b_add": (C.c_int, [vp, u64]),
            "fbo_b_add_many": (None, [vp, vp, i64]), "fbo_b_contains": (C.c_int, [vp, u64]),
            "fbo_b_count": (u64, [vp]), "fbo_b_any": (C.c_int, [vp]), "fbo_b_slice": (u64, [vp, vp, u64]),
            "fbo_b_intersect": (vp, [vp, vp]), "fbo_b_union": (vp, [vp, vp]), "fbo_b_union_n": (vp, [vp, vp, C.c_int]),
            "fbo_b_difference": (vp, [vp, vp]), "fbo_b_xor": (vp, [vp, vp]),
            "fbo_b_intersection_count": (u64, [vp, vp]), "fbo_b_optimize": (None, [vp]),
            "fbo_b_offset_range": (vp, [vp, u64, u64, u64]),
            "fbo_b_write": (u64, [vp, vp, u64, C.c_int]), "fbo_b_read": (vp, [vp, u64]),
            "fbo_frag_row": (vp, [vp, u64, u64]),
            "fbo_frag_range_op": (vp, [vp, u64, C.c_int, u64, i64, i64]),
            "fbo_frag_row_counts": (i64, [vp, u64, vp, vp, vp, i64]),
            "fbo_frag_rows": (i64, [vp, vp, i64]),
            "fbo_groupby_shard": (C.c_int, [vp, C.c_int, u64, vp, vp, vp, vp]),
            "fbo_frag_row_view": (vp, [vp, u64, u64]),
            "fbo_pool_create": (vp, [C.c_int]), "fbo_pool_create2": (vp, [C.c_int, C.c_int]), "fbo_pool_destroy": (None, [vp]), "fbo_pool_threads": (C.c_int, [vp]),
            "fbo_bench_union_intersect_count": (u64, [vp, vp, vp, i64, vp, C.c_int, vp, C.c_int, vp]),
            "fbo_bench_union_intersect_per_shard": (u64, [vp, vp, vp, i64, vp, C.c_int, vp, C.c_int, vp]),
            "fbo_bench_pair_counts": (u64, [vp, vp, vp, i64, vp, vp, C.c_int, C.c_int, vp, vp]),
            "fbo_bench_range_count": (u64, [vp, vp, vp, i64, C.c_int, u64, i64, i64, vp]),
            "fbo_bench_groupby": (C.c_int, [vp, vp, C.c_int, vp, i64, vp, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


class _CStruct(C.Structure):
    _fields_ = [("typ", C.c_uint8), ("n", C.c_int32), ("len", C.c_int32), ("data", C.c_void_p)]


class Container:
    """Owning handle of an fbo_container (None pointer == empty/nil container)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().fbo_c_free(self.ptr)
            self.ptr = None

    @staticmethod
    def array(vals):
        a = np.ascontiguousarray(np.asarray(vals, dtype=np.uint16))
        return Container(lib().fbo_c_array(a.ctypes.data, len(a)))

    @staticmethod
    def bitmap(words):
        w = np.ascontiguousarray(np.asarray(words, dtype=np.uint64))
        assert len(w) == 1024
        return Container(lib().fbo_c_bitmap(w.ctypes.data))

    @staticmethod
    def run(intervals):
        iv = np.ascontiguousarray(np.asarray(intervals, dtype=np.uint16).reshape(-1, 2))
        return Container(lib().fbo_c_run(iv.ctypes.data, len(iv)))

    @staticmethod
    def from_values(vals, typ):
        """container of an explicit encoding holding the given sorted unique values"""
        vals = np.asarray(sorted(set(int(v) for v in vals)), dtype=np.int64)
        if typ == ARRAY:
            return Container.array(vals)
        w = np.zeros(1024, dtype=np.uint64)
        if len(vals):
            np.bitwise_or.at(w, vals >> 6, np.uint64(1) << (vals & 63).astype(np.uint64))
        b = Container.bitmap(w)
        if typ == BITMAP:
            return b
        return b.convert(RUN)

    @property
    def typ(self):
        return 0 if not self.ptr else C.cast(self.ptr, C.POINTER(_CStruct)).contents.typ

    @property
    def n(self):
        return lib().fbo_c_n(self.ptr)

    def words(self):
        w = np.zeros(1024, dtype=np.uint64)
        lib().fbo_c_to_words(self.ptr, w.ctypes.data)
        return w

    def values(self):
        w = self.words()
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")
        return np.nonzero(bits)[0].astype(np.int64)

    def count_runs(self):
        return lib().fbo_c_count_runs(self.ptr)

    def convert(self, typ):
        return Container(lib().fbo_c_convert(self.ptr, typ))

    def optimized(self):
        return Container(lib().fbo_c_optimize(lib().fbo_c_clone(self.ptr)))

    def count_range(self, start, end):
        return lib().fbo_c_count_range(self.ptr, start, end)

    def _bin(self, fn, other):
        return Container(getattr(lib(), fn)(self.ptr, other.ptr))

    def intersect(self, o):
        return self._bin("fbo_intersect", o)

    def union(self, o):
        return self._bin("fbo_union", o)

    def difference(self, o):
        return self._bin("fbo_difference", o)

    def xor(self, o):
        return self._bin("fbo_xor", o)

    def flip(self):
        return Container(lib().fbo_flip(self.ptr))

    def intersection_count(self, o):
        return lib().fbo_intersection_count(self.ptr, o.ptr)


class Bitmap:
    """Owning handle of an fbo_bitmap (roaring.Bitmap)."""

    def __init__(self, ptr=None):
        self.ptr = ptr if ptr is not None else lib().fbo_b_new()

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().fbo_b_free(self.ptr)
            self.ptr = None

    @staticmethod
    def from_values(vals):
        b = Bitmap()
        b.add_many(vals)
        return b

    @staticmethod
    def from_bytes(data):
        data = bytes(data)
        p = lib().fbo_b_read(data, len(data))
        if not p:
            raise ValueError("unreadable roaring data")
        return Bitmap(p)

    def add(self, v):
        return bool(lib().fbo_b_add(self.ptr, int(v)))

    def add_many(self, vals):
        a = np.ascontiguousarray(np.asarray(vals, dtype=np.uint64))
        if len(a):
            lib().fbo_b_add_many(self.ptr, a.ctypes.data, len(a))

    def put(self, key, container):
        lib().fbo_b_put(self.ptr, int(key), lib().fbo_c_clone(container.ptr))

    def contains(self, v):
        return bool(lib().fbo_b_contains(self.ptr, int(v)))

    def count(self):
        return int(lib().fbo_b_count(self.ptr))

    def any(self):
        return bool(lib().fbo_b_any(self.ptr))

    def slice(self):
        n = self.count()
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().fbo_b_slice(self.ptr, out.ctypes.data, n)
        return out[:n]

    def to_bytes(self, optimize=True):
        need = lib().fbo_b_write(self.ptr, None, 0, 1 if optimize else 0)
        buf = (C.c_uint8 * max(need, 1))()
        lib().fbo_b_write(self.ptr, buf, need, 0)
        return bytes(buf[:need])

    def clone(self):
        return Bitmap(lib().fbo_b_clone(self.ptr))

    def _bin(self, fn, o):
        return Bitmap(getattr(lib(), fn)(self.ptr, o.ptr))

    def intersect(self, o):
        return self._bin("fbo_b_intersect", o)

    def union(self, *others):
        if len(others) == 1:
            return self._bin("fbo_b_union", others[0])
        if not others:
            return self.clone()
        arr = (C.c_void_p * len(others))(*[o.ptr for o in others])
        return Bitmap(lib().fbo_b_union_n(self.ptr, arr, len(others)))

    def difference(self, o):
        return self._bin("fbo_b_difference", o)

    def xor(self, o):
        return self._bin("fbo_b_xor", o)

    def intersection_count(self, o):
        return int(lib().fbo_b_intersection_count(self.ptr, o.ptr))

    def offset_range(self, offset, start, end):
        return Bitmap(lib().fbo_b_offset_range(self.ptr, offset, start, end))

    # ---- fragment-level views (self is a fragment: pos = row<<20 | col&(2^20-1)) ----
    def row(self, row, shard=0):
        return Bitmap(lib().fbo_frag_row(self.ptr, int(row), int(shard)))

    def range_op(self, op, bit_depth, predicate, predicate_max=0, shard=0):
        code = OPS[op] if isinstance(op, str) else op
        return Bitmap(lib().fbo_frag_range_op(self.ptr, int(shard), code, int(bit_depth), int(predicate), int(predicate_max)))

    def row_counts(self, shard=0, filt=None):
        cap = 1 << 16
        while True:
            rows = np.zeros(cap, dtype=np.uint64)
            cnts = np.zeros(cap, dtype=np.uint64)
            n = lib().fbo_frag_row_counts(self.ptr, int(shard), filt.ptr if filt is not None else None,
                                          rows.ctypes.data, cnts.ctypes.data, cap)
            if n <= cap:
                return rows[:n], cnts[:n]
            cap = int(n)

    def rows(self):
        cap = 1 << 16
        while True:
            rows = np.zeros(cap, dtype=np.uint64)
            n = lib().fbo_frag_rows(self.ptr, rows.ctypes.data, cap)
            if n <= cap:
                return rows[:n]
            cap = int(n)


# ---- BSI aggregates of one fragment (restated with the C oracle's row / set operations; rows 0 exists, 1 sign, 2+i bit i) ----
_M64 = (1 << 64) - 1


def _i64(x):
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


def bsi_sum(frag, bit_depth, filt=None, shard=0):
    """fragment.sum fragment.go:722-748 + BitmapBSICountFilter roaring/filter.go:1106-1165: (sum, count) with
    positive = (filter ∩ exists) minus sign, negative = filter ∩ exists ∩ sign; psum/nsum are uint64 accumulators of
    count << i, the total is int64(psum) - int64(nsum).  filt is a Bitmap with absolute keys, or None (no filter).
    The reference walks every value row stored in the fragment; well-formed fragments have none beyond bit_depth."""
    pos = frag.row(0, shard)
    if filt is not None:
        pos = pos.intersect(filt)
    count = pos.count()
    sign = frag.row(1, shard)
    neg = pos.intersect(sign)
    pos = pos.difference(sign)
    psum = nsum = 0
    for i in range(int(bit_depth)):
        plane = frag.row(2 + i, shard)
        psum = (psum + (pos.intersection_count(plane) << i)) & _M64
        nsum = (nsum + (neg.intersection_count(plane) << i)) & _M64
    return _i64(_i64(psum) - _i64(nsum)), count


def _bsi_min_unsigned(frag, filt, bit_depth, shard):
    """fragment.minUnsigned fragment.go:788-807"""
    mn, count = 0, filt.count()
    for i in range(int(bit_depth) - 1, -1, -1):
        row = filt.difference(frag.row(2 + i, shard))
        count = row.count()
        if count > 0:
            filt = row
        else:
            mn += 1 << i
            if i == 0:
                count = filt.count()
    return mn, count


def _bsi_max_unsigned(frag, filt, bit_depth, shard):
    """fragment.maxUnsigned fragment.go:841-860"""
    mx, count = 0, filt.count()
    for i in range(int(bit_depth) - 1, -1, -1):
        row = frag.row(2 + i, shard).intersect(filt)
        count = row.count()
        if count > 0:
            mx += 1 << i
            filt = row
        elif i == 0:
            count = filt.count()
    return mx, count


def bsi_min(frag, bit_depth, filt=None, shard=0):
    """fragment.min fragment.go:752-785 -> (min, count)"""
    consider = frag.row(0, shard)
    if filt is not None:
        consider = consider.intersect(filt)
    if consider.count() == 0:
        return 0, 0
    neg = frag.row(1, shard).intersect(consider)
    if neg.any():
        mx, count = _bsi_max_unsigned(frag, neg, bit_depth, shard)
        return -mx, count
    return _bsi_min_unsigned(frag, consider, bit_depth, shard)


def bsi_max(frag, bit_depth, filt=None, shard=0):
    """fragment.max fragment.go:811-838 -> (max, count)"""
    consider = frag.row(0, shard)
    if filt is not None:
        consider = consider.intersect(filt)
    if not consider.any():
        return 0, 0
    pos = consider.difference(frag.row(1, shard))
    if not pos.any():
        mn, count = _bsi_min_unsigned(frag, consider, bit_depth, shard)
        return -mn, count
    return _bsi_max_unsigned(frag, pos, bit_depth, shard)


def fragment_top(frag, shard=0, n=0, src=None, row_ids=None, min_threshold=0, tanimoto_threshold=0):
    """fragment.top (fragment.go:1317-1437) over one fragment whose rank cache holds every row (the exact case, SURVEY
    Appendix D): candidates = rows with a non-zero count, largest first (topBitmapPairs :1439-1480; ties pinned id ascending);
    cut-offs on the row's own count `cnt` and on `count` = |Src ∩ row| by MinThreshold or by the Tanimoto band / coefficient
    (:1329-1338, 1351-1362, 1378-1388); the first N pairs go to a min-heap, later rows only while their `cnt` can still
    beat the smallest kept count (:1401-1421).  Returns [(row, count)], count descending, ties id ascending."""
    import heapq
    import math
    rows, cnts = frag.row_counts(shard)
    pairs = [(int(r), int(c)) for r, c in zip(rows.tolist(), cnts.tolist()) if c > 0]
    if row_ids is not None and len(row_ids) > 0:
        have = dict(pairs)
        pairs = [(int(r), have[int(r)]) for r in row_ids if int(r) in have]
        n = 0                                                    # :1325-1327 ids given: no truncation
    pairs.sort(key=lambda p: (-p[1], p[0]))
    tan = 0
    src_count = 0
    if tanimoto_threshold > 0 and src is not None:
        tan = int(tanimoto_threshold)
        src_count = src.count()
        min_tan = float(src_count * tan) / 100
        max_tan = float(src_count * 100) / float(tan)
    heap = []                                                    # pairHeap: min-heap on Count (cache.go:395-431)
    for row, cnt in pairs:
        if tan > 0:
            if float(cnt) <= min_tan or float(cnt) >= max_tan:
                continue
        elif cnt < min_threshold:
            continue
        if n == 0 or len(heap) < n:
            count = cnt
            if src is not None:
                count = src.intersection_count(frag.row(row, shard))
            if count == 0:
                continue
            if tan > 0:
                if math.ceil(float(count * 100) / float(cnt + src_count - count)) <= float(tan):
                    continue
            elif count < min_threshold:
                continue
            heapq.heappush(heap, (count, row))
            if n > 0 and len(heap) == n and src is None:
                break
            continue
        threshold = heap[0][0]
        if threshold < min_threshold or cnt < threshold:
            break
        count = src.intersection_count(frag.row(row, shard))
        if count < threshold:
            continue
        heapq.heappush(heap, (count, row))
    return sorted(((r, c) for c, r in heap), key=lambda p: (-p[1], p[0]))


def groupby_shard(frags, shard, row_ids, filt=None, out=None):
    """frags: list of Bitmap-or-None; row_ids: list of lists; returns dense count array (row-major)."""
    n_rows = np.asarray([len(r) for r in row_ids], dtype=np.int32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint64) for r in row_ids]) if len(row_ids) else np.zeros(0, np.uint64))
    if out is None:
        out = np.zeros(int(np.prod(n_rows)), dtype=np.uint64)
    arr = (C.c_void_p * len(frags))(*[(f.ptr if f is not None else None) for f in frags])
    rc = lib().fbo_groupby_shard(arr, len(frags), int(shard), flat.ctypes.data, n_rows.ctypes.data,
                                 filt.ptr if filt is not None else None, out.ctypes.data)
    assert rc == 0
    return out


class Pool:
    """Long-lived pinned worker threads of the CPU baseline (fb_bench.c): created once, reused by every bench call."""

    def __init__(self, n_threads=None, pin=True):
        self.n = int(n_threads or len(os.sched_getaffinity(0)) or 1)
        self.pin = bool(pin)
        self.ptr = lib().fbo_pool_create2(self.n, 1 if pin else 0)

    def close(self):
        if self.ptr:
            lib().fbo_pool_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _frag_array(frags):
    return (C.c_void_p * len(frags))(*[(f.ptr if f is not None else None) for f in frags])


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def bench_union_intersect_count(pool, frags, shards, rows_a, rows_b):
    """CPU baseline: Count(Intersect(Union(rows_a), Union(rows_b))) over the given fragments. Returns (count, seconds)."""
    arr, sh, ra, rb = _frag_array(frags), _u64(shards), _u64(rows_a), _u64(rows_b)
    secs = C.c_double(0)
    tot = lib().fbo_bench_union_intersect_count(pool.ptr, arr, sh.ctypes.data, len(frags), ra.ctypes.data, len(ra),
                                                rb.ctypes.data, len(rb), C.byref(secs))
    return int(tot), secs.value


def union_intersect_per_shard(pool, frags, shards, rows_a, rows_b):
    """Count(Intersect(Union(rows_a), Union(rows_b))) of EVERY shard (threaded). Returns the per-shard vector."""
    arr, sh, ra, rb = _frag_array(frags), _u64(shards), _u64(rows_a), _u64(rows_b)
    per = np.zeros(len(frags), dtype=np.uint64)
    tot = lib().fbo_bench_union_intersect_per_shard(pool.ptr, arr, sh.ctypes.data, len(frags), ra.ctypes.data, len(ra), rb.ctypes.data, len(rb), per.ctypes.data)
    assert int(tot) == int(per.sum())
    return per


def bench_pair_counts(pool, frags, shards, rows_a, rows_b, materialise=True):
    """CPU baseline: Count(Intersect(Row(a_k), Row(b_k))) for every pair k. Returns (per-pair counts, seconds)."""
    arr, sh, ra, rb = _frag_array(frags), _u64(shards), _u64(rows_a), _u64(rows_b)
    out = np.zeros(len(ra), dtype=np.uint64)
    secs = C.c_double(0)
    tot = lib().fbo_bench_pair_counts(pool.ptr, arr, sh.ctypes.data, len(frags), ra.ctypes.data, rb.ctypes.data, len(ra),
                                      1 if materialise else 0, out.ctypes.data, C.byref(secs))
    assert int(tot) == int(out.sum())
    return out, secs.value


def bench_range_count(pool, frags, shards, op, bit_depth, predicate, predicate_max=0):
    """CPU baseline: Count(Row(v <op> predicate)) over BSI fragments. Returns (count, seconds)."""
    arr, sh = _frag_array(frags), _u64(shards)
    secs = C.c_double(0)
    tot = lib().fbo_bench_range_count(pool.ptr, arr, sh.ctypes.data, len(frags), OPS[op] if isinstance(op, str) else int(op),
                                      int(bit_depth), int(predicate), int(predicate_max), C.byref(secs))
    return int(tot), secs.value


def bench_groupby(pool, frags_per_field, shards, row_ids):
    """CPU baseline: GroupBy over set fields; frags_per_field[f][s]. Returns (dense counts, seconds)."""
    nf, ns = len(frags_per_field), len(shards)
    flat_frags = [f for per in frags_per_field for f in per]
    arr, sh = _frag_array(flat_frags), _u64(shards)
    n_rows = np.asarray([len(r) for r in row_ids], dtype=np.int32)
    flat = _u64(np.concatenate([np.asarray(r, dtype=np.uint64) for r in row_ids]))
    out = np.zeros(int(np.prod(n_rows)), dtype=np.uint64)
    secs = C.c_double(0)
    rc = lib().fbo_bench_groupby(pool.ptr, arr, nf, sh.ctypes.data, ns, flat.ctypes.data, n_rows.ctypes.data, out.ctypes.data, C.byref(secs))
    assert rc == 0
    return out, secs.value
