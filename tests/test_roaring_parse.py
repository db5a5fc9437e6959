"""featurebase_b200/csrc/roaring_parse.h on the CPU: the product's reader of serialised fragments (Pilosa format and the
official RoaringBitmap format with the reference's quirks) against the reference's golden images, against the oracle's
reader on random fragments, and against truncated / corrupted inputs (must fail cleanly; every payload view it accepts must
lie inside the buffer)."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from featurebase_b200 import datagen as D
from oracle import oracle as O
from tests.golden import vectors as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

@pytest.fixture(scope="module")
def pc():
    out = os.path.join(tempfile.mkdtemp(prefix="parse_check_"), "libparse_check.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "featurebase_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "parse_check.cpp"), "-o", out])
    L = C.CDLL(out)
    L.parse_dump.restype = C.c_longlong
    L.parse_dump.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_ulonglong]
    return L


def parse(pc, data):
    """-> [(key, typ, n, cnt, official_run, payload)] or raises ValueError(message)"""
    buf = C.create_string_buffer(bytes(data), len(data)) if len(data) else C.create_string_buffer(1)
    cap = 1 << 16
    while True:
        out = C.create_string_buffer(cap)
        n = pc.parse_dump(buf, len(data), out, cap)
        if n < 0:
            raise ValueError(out.value.decode())
        if n <= cap:
            break
        cap = int(n)
    res = []
    for line in out.value.decode().splitlines():
        f = line.split(" ")
        res.append((int(f[0]), int(f[1]), int(f[2]), int(f[3]), f[4] == "1", bytes.fromhex(f[5]) if len(f) > 5 else b""))
    return res


def values_of(conts):
    out = []
    for key, typ, n, cnt, official, payload in conts:
        if typ == 1:
            v = np.frombuffer(payload, dtype="<u2").astype(np.uint64)
        elif typ == 2:
            v = np.flatnonzero(np.unpackbits(np.frombuffer(payload, dtype=np.uint8), bitorder="little")).astype(np.uint64)
        else:
            r = np.frombuffer(payload, dtype="<u2").reshape(-1, 2).astype(np.int64)
            if official:                                   # (start, length-1) as stored; the loader converts to (start, last)
                r = np.stack([r[:, 0], r[:, 0] + r[:, 1]], axis=1)
            v = np.concatenate([np.arange(s, l + 1, dtype=np.uint64) for s, l in r.tolist()])
        assert len(v) == n, (key, typ, n, len(v))
        out.append(v + np.uint64(key << 16))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint64)


def test_official_format_goldens(pc):
    """roaring_internal_test.go:3793-3853"""
    for hx, exp in V.OFFICIAL_HEX:
        assert values_of(parse(pc, bytes.fromhex(hx))).tolist() == exp
    name, count = V.OFFICIAL_FILE
    data = open(os.path.join(GOLD, name), "rb").read()
    conts = parse(pc, data)
    assert [(c[0], c[1], c[2]) for c in conts] == [(0, 2, 9999), (1, 1, 1)]
    assert len(values_of(conts)) == count
    assert np.array_equal(values_of(conts), O.Bitmap.from_bytes(data).slice())
    for hx in V.OFFICIAL_ZERO_CONTAINER_ERRORS:                                          # the reference rejects zero-container images
        with pytest.raises(ValueError):
            parse(pc, bytes.fromhex(hx))
    assert parse(pc, bytes.fromhex(V.PILOSA_EMPTY_OK)) == []
    for junk in (b"", b"\x3c", b"\x3c\x30\x00", b"\x00" * 8, b"\x3c\x30\x01\x00\x00\x00\x00\x00"):
        with pytest.raises(ValueError):
            parse(pc, junk)


def test_pilosa_fragments_match_oracle_reader(pc):
    for seed, dens, mode in ((1, 0.01, 0), (2, 0.3, 0), (3, 0.2, 1), (4, 0.0001, 0), (5, 0.9, 1)):
        data = D.fragment(seed, 5, [0, 3, 7, 200], dens, mode=mode, mean_run=300.0)
        conts = parse(pc, data)
        ob = O.Bitmap.from_bytes(data)
        assert np.array_equal(values_of(conts), ob.slice())
        assert [c[0] for c in conts] == sorted(c[0] for c in conts)
        # header cardinalities are what the file says (N-1 as u16)
        n_hdr, = struct.unpack_from("<I", data, 4)
        assert len(conts) == n_hdr


def test_truncated_and_corrupted_inputs_fail_cleanly(pc):
    rng = np.random.default_rng(2)
    merged = O.Bitmap()
    for d in (D.fragment(7, 1, [0, 1], 0.01), D.fragment(7, 1, [2], 0.3), D.fragment(7, 1, [3], 0.2, mode=1, mean_run=200.0)):
        merged = merged.union(O.Bitmap.from_bytes(d))
    data = merged.to_bytes()
    good = values_of(parse(pc, data))
    assert np.array_equal(good, merged.slice())
    cuts = sorted(set([0, 1, 7, 8, 9, 19, 20] + rng.integers(0, len(data), 300).tolist() + [len(data) - 1]))
    for cut in cuts:
        try:
            conts = parse(pc, data[:cut])
        except ValueError:
            continue
        values_of(conts)                                   # whatever parses must be internally consistent and in bounds
    name, _ = V.OFFICIAL_FILE
    off = open(os.path.join(GOLD, name), "rb").read()
    for cut in sorted(set([0, 4, 8, 12, 16, 20, 23, 24, 8215, 8216, 8217] + rng.integers(0, len(off), 100).tolist())):
        try:
            parse(pc, off[:cut])
        except ValueError:
            pass
    for _ in range(300):                                   # random byte flips in the header / offset tables
        b = bytearray(data)
        hdr = 8 + 16 * struct.unpack_from("<I", data, 4)[0]
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, hdr))] = int(rng.integers(0, 256))
        try:
            conts = parse(pc, bytes(b))
        except ValueError:
            continue
        for key, typ, n, cnt, official, payload in conts:  # accepted => every payload view lay inside the buffer
            assert len(payload) == (2 * n if typ == 1 else 8192 if typ == 2 else 4 * cnt)
