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


def _fnv32a(*parts):
    h = 2166136261
    for p in parts:
        for b in p:
            h = ((h ^ b) * 16777619) & 0xFFFFFFFF
    return h


def _op(typ, value=0, values=None, roaring=None, opn=0):
    """op.WriteTo (roaring.go:6325-6366): type, value / length, fnv32a checksum, payload"""
    if typ in (0, 1):
        head, tail = struct.pack("<BQ", typ, value), b""
    elif typ in (2, 3):
        head, tail = struct.pack("<BQ", typ, len(values)), b"".join(struct.pack("<Q", v) for v in values)
    else:
        head, tail = struct.pack("<BQ", typ, len(roaring)), struct.pack("<I", opn) + roaring
    return head + struct.pack("<I", _fnv32a(head, tail)) + tail


def test_ops_log_is_replayed_by_the_oracle_and_refused_by_the_product(pc):
    """Bytes after the last container are an ops log (unmarshal_binary.go:66-92).  The oracle replays it like the reference (the op
    list of TestOpLogWriteUnmarshal, roaring_internal_test.go:4007-4060, plus the two roaring op types); the product's loader
    refuses such an image instead of loading the containers without the log."""
    base_vals = [1, 6, 28, 44, 70000, (5 << 16) + 3] + list(range(200000, 200000 + 5000))
    base = O.Bitmap.from_values(base_vals)
    add_r = O.Bitmap.from_values([9, 70001, (9 << 16) + 1]).to_bytes()
    rem_r = O.Bitmap.from_values([70000, 200001, 999999]).to_bytes()
    ops = [(0, 27), (1, 28), (2, [1, 2, 6, 19]), (3, [1, 2, 6, 19, 22, 44]), (2, [51234567890]), (3, [51234567890]), (0, 0), (1, 0),
           (2, [0]), (3, [0]), (2, []), (3, []), (4, add_r), (5, rem_r), (0, 200001)]
    def replay(start):
        model, log = set(start), b""
        for typ, arg in ops:
            if typ in (0, 1):
                log += _op(typ, value=arg)
                (model.add if typ == 0 else model.discard)(arg)
            elif typ in (2, 3):
                log += _op(typ, values=arg)
                model = model | set(arg) if typ == 2 else model - set(arg)
            else:
                log += _op(typ, roaring=arg, opn=3)
                vals = set(O.Bitmap.from_bytes(arg).slice().tolist())
                model = model | vals if typ == 4 else model - vals
        return sorted(model), log
    for start in (base_vals, []):                              # zero containers and a log: still replayed (roaring.go:1994-2001)
        image = O.Bitmap.from_values(start).to_bytes()
        model, log = replay(start)
        assert O.Bitmap.from_bytes(image + log).slice().tolist() == model
        with pytest.raises(ValueError, match="ops log"):
            parse(pc, image + log)
        assert values_of(parse(pc, image)).tolist() == sorted(start)
    _, log = replay(base_vals)
    data = base.to_bytes() + log
    bad = bytearray(data); bad[len(base.to_bytes()) + 9] ^= 1      # checksum of the first op
    with pytest.raises(ValueError):
        O.Bitmap.from_bytes(bytes(bad))
    with pytest.raises(ValueError):
        O.Bitmap.from_bytes(data[:-3])                             # truncated op
    with pytest.raises(ValueError):
        O.Bitmap.from_bytes(base.to_bytes() + bytes([9]) + bytes(12))   # unknown op type


def test_official_format_header_strictness(pc):
    """readOfficialHeader (roaring.go:6960-6996): the no-run cookie is a 32-bit compare; more than 2^16 containers is refused.
    The reference Put()s containers by key (a repeated key replaces, order is free): the oracle does the same, the product
    refuses images whose keys are not strictly ascending (no writer emits them; the descriptor tables index by key rank)."""
    # two array containers, no-run cookie: key 3 {1,2}, key 3 {7}  -> the second replaces the first in the reference
    def norun(keys_cards_payloads):
        n = len(keys_cards_payloads)
        hdr = b"".join(struct.pack("<HH", k, len(v) - 1) for k, v in keys_cards_payloads)
        off, offs, body = 8 + 8 * n, b"", b""
        for _, v in keys_cards_payloads:
            offs += struct.pack("<I", off + len(body)); body += b"".join(struct.pack("<H", x) for x in v)
        return struct.pack("<II", 12346, n) + hdr + offs + body
    ok = norun([(1, [5, 6]), (3, [1, 2])])
    assert values_of(parse(pc, ok)).tolist() == [(1 << 16) + 5, (1 << 16) + 6, (3 << 16) + 1, (3 << 16) + 2]
    assert O.Bitmap.from_bytes(ok).slice().tolist() == [(1 << 16) + 5, (1 << 16) + 6, (3 << 16) + 1, (3 << 16) + 2]
    dup, unordered = norun([(3, [1, 2]), (3, [7])]), norun([(3, [1, 2]), (1, [7])])
    assert O.Bitmap.from_bytes(dup).slice().tolist() == [(3 << 16) + 7]
    assert O.Bitmap.from_bytes(unordered).slice().tolist() == [(1 << 16) + 7, (3 << 16) + 1, (3 << 16) + 2]
    for img in (dup, unordered):
        with pytest.raises(ValueError, match="ascending"):
            parse(pc, img)
    hi = bytearray(ok); hi[2] = 1                              # 0x0001303A: low 16 bits say no-run, the cookie does not
    with pytest.raises(ValueError):
        parse(pc, bytes(hi))
    with pytest.raises(ValueError):
        O.Bitmap.from_bytes(bytes(hi))
    many = struct.pack("<II", 12346, (1 << 16) + 1) + bytes(8 * ((1 << 16) + 1) + 64)
    with pytest.raises(ValueError):
        parse(pc, many)
    with pytest.raises(ValueError):
        O.Bitmap.from_bytes(many)
