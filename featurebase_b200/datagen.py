"""ctypes wrapper of tools/libfbdatagen.so — synthetic fragments as Pilosa-roaring bytes (tests + bench)."""
import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "tools", "libfbdatagen.so")
        if not os.path.exists(path):
            from . import build
            build.build_datagen()
        L = C.CDLL(path)
        vp, u32, u64, i64, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64, C.c_double
        L.fbdg_fragment.argtypes, L.fbdg_fragment.restype = [u64, u32, u64, vp, C.c_int, dbl, C.c_int, dbl, C.POINTER(u64)], vp
        L.fbdg_bsi_fragment.argtypes, L.fbdg_bsi_fragment.restype = [u64, u32, u64, u64, C.c_int, i64, i64, i64, dbl, C.POINTER(u64)], vp
        L.fbdg_bsi_value.argtypes, L.fbdg_bsi_value.restype = [u64, u32, u64, u64, i64, i64, dbl, C.POINTER(i64)], C.c_int
        L.fbdg_groupby_fragments.argtypes = [u64, u32, u32, u64, dbl, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
        L.fbdg_groupby_fragments.restype = C.c_int
        L.fbdg_fragments.argtypes, L.fbdg_fragments.restype = [u64, u32, vp, i64, vp, C.c_int, dbl, C.c_int, dbl, C.c_int, vp], vp
        L.fbdg_free.argtypes, L.fbdg_free.restype = [vp], None
        _LIB = L
    return _LIB


def _take(ptr, n):
    data = C.string_at(ptr, n)
    _lib().fbdg_free(ptr)
    return data


def fragment(field, shard, rows, p, mode=0, mean_run=64.0, seed=0):
    rows = np.ascontiguousarray(np.asarray(rows, dtype=np.uint64))
    n = C.c_uint64(0)
    ptr = _lib().fbdg_fragment(seed, field, int(shard), rows.ctypes.data, len(rows), float(p), mode, float(mean_run), C.byref(n))
    return _take(ptr, n.value)


class Bulk:
    """n_shards fragments in one C buffer (kept alive by this object); .buf is a uint8 numpy view"""

    def __init__(self, ptr, offsets):
        self.ptr, self.offsets = ptr, offsets
        total = int(offsets[-1])
        self.buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total]

    def fragment_bytes(self, i):
        return bytes(self.buf[int(self.offsets[i]):int(self.offsets[i + 1])])

    def __del__(self):
        if getattr(self, "ptr", None):
            _lib().fbdg_free(self.ptr)
            self.ptr = None


def fragments(field, shards, rows, p, mode=0, mean_run=64.0, seed=0, threads=None):
    shards = np.ascontiguousarray(np.asarray(shards, dtype=np.uint64))
    rows = np.ascontiguousarray(np.asarray(rows, dtype=np.uint64))
    offsets = np.zeros(len(shards) + 1, dtype=np.uint64)
    threads = threads or os.cpu_count() or 1
    ptr = _lib().fbdg_fragments(seed, field, shards.ctypes.data, len(shards), rows.ctypes.data, len(rows), float(p), mode,
                                float(mean_run), int(threads), offsets.ctypes.data)
    return Bulk(ptr, offsets)


def bsi_fragment(field, shard, n_cols, bit_depth, lo, hi, base=0, null_frac=0.0, seed=0):
    n = C.c_uint64(0)
    ptr = _lib().fbdg_bsi_fragment(seed, field, int(shard), int(n_cols), bit_depth, lo, hi, base, float(null_frac), C.byref(n))
    return _take(ptr, n.value)


def bsi_value(field, shard, col, lo, hi, null_frac=0.0, seed=0):
    v = C.c_int64(0)
    ok = _lib().fbdg_bsi_value(seed, field, int(shard), int(col), lo, hi, float(null_frac), C.byref(v))
    return v.value if ok else None


def groupby_fragments(field_a, field_b, shard, p_rec, na, nb, seed=0):
    pa, pb, la, lb = C.c_void_p(), C.c_void_p(), C.c_uint64(0), C.c_uint64(0)
    _lib().fbdg_groupby_fragments(seed, field_a, field_b, int(shard), float(p_rec), na, nb, C.byref(pa), C.byref(la), C.byref(pb), C.byref(lb))
    return _take(pa, la.value), _take(pb, lb.value)
