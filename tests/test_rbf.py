"""featurebase_b200/csrc/rbf_reader.h — the host-side RBF walker behind fbgpu_load_rbf.

Pinned by the reference's own RBF fixture bytes (tests/golden/vectors.py:RBF_FIXTURE_PAGES); everything larger comes from
tests/rbf_writer.py, whose output is first checked byte for byte against the same fixture."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from featurebase_b200 import datagen as D
from oracle import oracle as O
from tests import rbf_writer as W
from tests.golden import vectors as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAGE = 8192


_HARNESS = None


def harness():
    """g++-built wrapper around rbf_reader.h (host code only), compiled once per process into a temp dir"""
    global _HARNESS
    if _HARNESS is None:
        import tempfile
        out = os.path.join(tempfile.mkdtemp(prefix="rbf_check_"), "librbf_check.so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "featurebase_b200", "csrc"),
                               os.path.join(ROOT, "tests", "native", "rbf_check.cpp"), "-o", out])
        L = C.CDLL(out)
        L.rbf_dump.restype = C.c_longlong
        L.rbf_dump.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong]
        _HARNESS = L
    return _HARNESS


@pytest.fixture(scope="module")
def lib():
    return harness()


def cell_values(key, typ, elem_n, bit_n, payload):
    """absolute bit positions (key << 16 | low) of one dumped leaf cell"""
    if typ == W.C_ARRAY:
        v = np.frombuffer(payload, dtype="<u2").astype(np.uint64)
    elif typ == W.C_RLE:
        r = np.frombuffer(payload, dtype="<u2").reshape(-1, 2)
        v = np.concatenate([np.arange(s, l + 1, dtype=np.uint64) for s, l in r.tolist()])
    else:
        v = np.flatnonzero(np.unpackbits(np.frombuffer(payload, dtype=np.uint8), bitorder="little")).astype(np.uint64)
    assert len(v) == bit_n
    return v + np.uint64(key << 16)


def dump(L, data, wal=b""):
    """-> {name: [(key, type, elemN, bitN, payload bytes)]}; raises ValueError with the reader's message"""
    cap = 1 << 20
    while True:
        buf = C.create_string_buffer(cap)
        n = L.rbf_dump(data, len(data), wal or None, len(wal), buf, cap)
        if n < 0:
            raise ValueError(buf.value.decode())
        if n <= cap:
            break
        cap = int(n)
    out, cur = {}, None
    for line in buf.value.decode().splitlines():
        f = line.split(" ")
        if f[0] == "B":
            cur = out.setdefault(bytes.fromhex(f[1]).decode(), [])
        else:
            cur.append((int(f[1]), int(f[2]), int(f[3]), int(f[4]), bytes.fromhex(f[5]) if len(f) > 5 else b""))
    return out


def fixture(name):
    return b"".join(bytes.fromhex(h).ljust(PAGE, b"\0") for h in V.RBF_FIXTURE_PAGES[name])


def test_reference_fixture_bytes(lib):
    good = fixture("bad-freelist")
    assert dump(lib, good) == {"x": [(0, W.C_ARRAY, 1, 1, struct.pack("<H", 100))]}
    with pytest.raises(ValueError):                       # leaf page flagged as a branch: the walk must fail, not invent cells
        dump(lib, fixture("bad-bitmap"))
    with pytest.raises(ValueError, match="magic"):
        dump(lib, b"\0" * PAGE)
    with pytest.raises(ValueError):
        dump(lib, good[:100])
    # the test writer reproduces the reference's bytes for the same content (page 2 = freelist is the damaged one)
    mine = W.build({"x": [(0, "array", np.array([100], dtype=np.uint16))]}, wal_id=4)
    assert len(mine) == len(good)
    for pg in (0, 1, 3):
        assert mine[pg * PAGE:(pg + 1) * PAGE] == good[pg * PAGE:(pg + 1) * PAGE], pg
    assert mine[2 * PAGE:2 * PAGE + 4] == good[2 * PAGE:2 * PAGE + 4]


def _expected_cells(conts):
    exp = []
    for key, kind, payload in conts:
        if (kind == "array" and len(payload) > W.ARRAY_MAX) or (kind == "run" and len(payload) > W.RLE_MAX):
            kind, payload = "bitmap", W._to_bitmap(kind, payload)
        if kind == "array":
            exp.append((key, W.C_ARRAY, len(payload), len(payload), payload.astype("<u2").tobytes()))
        elif kind == "run":
            n = int((payload[:, 1].astype(np.int64) - payload[:, 0] + 1).sum())
            exp.append((key, W.C_RLE, len(payload), n, payload.astype("<u2").tobytes()))
        else:
            exp.append((key, W.C_BITMAP_PTR, 0, W._popcount(payload), payload.astype("<u8").tobytes()))
    return exp


def _fragment_containers(seed, shard):
    parts = [D.fragment(seed, shard, [0, 1, 2], 0.01), D.fragment(seed, shard, [3], 0.3), D.fragment(seed, shard, [5], 0.2, mode=1, mean_run=200.0),
             D.fragment(seed, shard, [6], 0.9, mode=1, mean_run=5000.0), D.fragment(seed, shard, [9], 0.0622), D.fragment(seed, shard, [40], 0.0001)]
    merged = O.Bitmap()
    for d in parts:
        merged = merged.union(O.Bitmap.from_bytes(d))
    return merged, W.cells_from_pilosa(merged.to_bytes())


def test_roundtrip_mixed_fragments(lib):
    """array / run / bitmap containers of several 'fields' of one shard, incl. arrays of 4080..4095 elements, which RBF
    stores as bitmap pages (ArrayMaxSize 4079)"""
    bitmaps, oracle_bm = {}, {}
    for k, name in enumerate(("~f;standard<", "~g;standard<", "~v;bsig_v<")):
        bm, conts = _fragment_containers(30 + k, 7)
        oracle_bm[name], bitmaps[name] = bm, conts
    assert any(kind == "array" and len(p) > W.ARRAY_MAX for _, kind, p in bitmaps["~f;standard<"])       # row 9 @6.22 %: ~4076 +- 60
    data = W.build(bitmaps)
    got = dump(lib, data)
    assert sorted(got) == sorted(bitmaps)
    for name, conts in bitmaps.items():
        assert got[name] == _expected_cells(conts), name
        # and, independently of the writer's bookkeeping: the decoded bits are the fragment's bits
        bits = []
        for cell in got[name]:
            bits.append(cell_values(*cell))
        assert np.array_equal(np.concatenate(bits), oracle_bm[name].slice())


def test_multi_level_tree_and_root_overflow(lib):
    """> 454 leaf pages under one bitmap (two branch levels) and enough bitmaps to overflow the root-record page"""
    rng = np.random.default_rng(3)
    n = 150000
    keys = np.sort(rng.choice(1 << 30, n, replace=False))
    big = [(int(k), "array", np.sort(rng.choice(65536, int(rng.integers(1, 4)), replace=False)).astype(np.uint16)) for k in keys]
    bitmaps = {"~big;standard<": big}
    for i in range(600):
        bitmaps["~field_with_a_long_name_%04d;standard<" % i] = [(i, "array", np.array([i], dtype=np.uint16))]
    data = W.build(bitmaps)
    got = dump(lib, data)
    assert len(got) == 601
    assert got["~big;standard<"] == _expected_cells(big)
    assert got["~field_with_a_long_name_0599;standard<"] == [(599, W.C_ARRAY, 1, 1, struct.pack("<H", 599))]
    # depth check: the root of the big bitmap is a branch whose children are branches
    w = W.Writer()
    root = w.add_bitmap("b", big)
    pages = w.finish()
    flags = lambda pg: struct.unpack_from(">I", pages[pg], 4)[0]
    child = struct.unpack_from("<I", pages[root], struct.unpack_from(">H", pages[root], 10)[0] + 12)[0]
    assert flags(root) == W.T_BRANCH and flags(child) == W.T_BRANCH


def test_wal_overlay(lib):
    """committed WAL pages replace data pages (incl. raw bitmap pages behind bitmap headers and a new meta page that grows
    the file); pages after the last meta page belong to an unfinished transaction and are ignored"""
    _, conts_a = _fragment_containers(50, 2)
    _, conts_b = _fragment_containers(51, 2)
    wa = W.Writer()
    wa.add_bitmap("~f;standard<", conts_a)
    old = wa.finish(wal_id=1)
    wb = W.Writer()
    wb.add_bitmap("~f;standard<", conts_b)
    wb.add_bitmap("~g;standard<", conts_a)
    new = wb.finish(wal_id=9)
    wal = W.wal_between(old, new, wb.raw)
    exp = dump(lib, b"".join(new))
    assert dump(lib, b"".join(old), wal) == exp
    assert dump(lib, b"".join(old)) != exp
    junk = bytearray(PAGE)
    struct.pack_into(">II", junk, 0, 3, W.T_LEAF)                         # an uncommitted rewrite of page 3
    assert dump(lib, b"".join(old), wal + bytes(junk)) == exp
    # a WAL without any meta page commits nothing
    assert dump(lib, b"".join(old), wal[:PAGE]) == dump(lib, b"".join(old))


def test_corrupted_and_truncated_files_fail_cleanly(lib):
    """random byte flips (biased to page headers, cell indexes and cell headers) and truncations: the walker either reports
    an error or returns cells whose payload views lie inside the file — it must never read out of bounds"""
    rng = np.random.default_rng(1)
    _, conts = _fragment_containers(70, 1)
    data = W.build({"~f;standard<": conts, "~g;standard<": conts[:50]})
    pages = len(data) // PAGE
    errors = 0
    for _ in range(600):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            pg = int(rng.integers(0, pages))
            off = pg * PAGE + int(rng.choice([rng.integers(0, 64), rng.integers(0, 2048), rng.integers(0, PAGE)]))
            b[off] = int(rng.integers(0, 256))
        try:
            out = dump(lib, bytes(b))
        except ValueError:
            errors += 1
            continue
        for cells in out.values():
            for key, typ, elem_n, bit_n, payload in cells:
                assert len(payload) == (2 * elem_n if typ == W.C_ARRAY else 4 * elem_n if typ == W.C_RLE else PAGE)
    assert errors > 20
    for cut in range(0, len(data), 4096):
        try:
            dump(lib, data[:cut])
        except ValueError:
            pass
