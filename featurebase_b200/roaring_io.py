"""Pilosa-roaring (de)serialisation on the host side of the C ABI, numpy only.

encode(): what roaring.Bitmap.WriteTo emits (reference roaring/roaring.go:1730-1817) for a set of uint64 values,
with canonical optimize() encodings (:3412-3426).  decode(): what Row.Columns() needs (row.go:471): all values of
a Pilosa-roaring buffer.  Used by the host mirror (executor.py) to import test data and to expand Row results."""
import numpy as np

MAGIC = 12348
ARRAY, BITMAP, RUN = 1, 2, 3


def _container_payload(lows):
    """lows: sorted unique uint16 values of one container -> (typ, n, payload bytes)"""
    n = len(lows)
    l64 = lows.astype(np.int64)
    brk = np.nonzero(np.diff(l64) != 1)[0]
    runs = len(brk) + 1
    if runs <= 2048 and runs <= n // 2:
        starts = np.concatenate([[l64[0]], l64[brk + 1]]).astype("<u2")
        lasts = np.concatenate([l64[brk], [l64[-1]]]).astype("<u2")
        body = np.stack([starts, lasts], axis=1).tobytes()
        return RUN, n, np.array([runs], dtype="<u2").tobytes() + body
    if n < 4096:
        return ARRAY, n, lows.astype("<u2").tobytes()
    w = np.zeros(1024, dtype=np.uint64)
    np.bitwise_or.at(w, l64 >> 6, np.uint64(1) << (l64 & 63).astype(np.uint64))
    return BITMAP, n, w.astype("<u8").tobytes()


def encode(values):
    v = np.unique(np.asarray(values, dtype=np.uint64))
    keys = v >> np.uint64(16)
    conts = []
    if len(v):
        bounds = np.concatenate([[0], np.nonzero(np.diff(keys))[0] + 1, [len(v)]])
        for a, b in zip(bounds[:-1], bounds[1:]):
            typ, n, payload = _container_payload((v[a:b] & np.uint64(0xFFFF)).astype(np.uint16))
            conts.append((int(keys[a]), typ, n, payload))
    out = bytearray()
    out += np.array([MAGIC], dtype="<u4").tobytes() + np.array([len(conts)], dtype="<u4").tobytes()
    for key, typ, n, _ in conts:
        out += np.array([key], dtype="<u8").tobytes() + np.array([typ, n - 1], dtype="<u2").tobytes()
    off = 8 + 16 * len(conts)
    for _, _, _, payload in conts:
        out += np.array([off & 0xFFFFFFFF], dtype="<u4").tobytes()
        off += len(payload)
    for _, _, _, payload in conts:
        out += payload
    return bytes(out)


def containers(data):
    """yields (key, typ, n, payload view) of a Pilosa-roaring buffer"""
    buf = np.frombuffer(data, dtype=np.uint8)
    if len(buf) < 8 or int(buf[:2].view("<u2")[0]) != MAGIC:
        raise ValueError("not pilosa roaring data")
    cnt = int(buf[4:8].view("<u4")[0])
    hdr = buf[8:8 + 12 * cnt]
    offs = buf[8 + 12 * cnt:8 + 16 * cnt].view("<u4")
    chunk, prev = 0, 0
    for i in range(cnt):
        key = int(hdr[12 * i:12 * i + 8].view("<u8")[0])
        typ, n1 = (int(x) for x in hdr[12 * i + 8:12 * i + 12].view("<u2"))
        o = int(offs[i])
        if o < prev:
            chunk += 1 << 32
        prev = o
        o += chunk
        n = n1 + 1
        if typ == ARRAY:
            yield key, typ, n, buf[o:o + 2 * n]
        elif typ == BITMAP:
            yield key, typ, n, buf[o:o + 8192]
        else:
            rc = int(buf[o:o + 2].view("<u2")[0])
            yield key, typ, n, buf[o + 2:o + 2 + 4 * rc]


def decode(data):
    """all values, ascending (Row.Columns, row.go:471)"""
    parts = []
    for key, typ, n, payload in containers(data):
        base = np.uint64(key) << np.uint64(16)
        if typ == ARRAY:
            lows = payload.view("<u2").astype(np.uint64)
        elif typ == BITMAP:
            bits = np.unpackbits(payload, bitorder="little")
            lows = np.nonzero(bits)[0].astype(np.uint64)
        else:
            r = payload.view("<u2").astype(np.int64).reshape(-1, 2)
            lows = np.concatenate([np.arange(s, l + 1) for s, l in r]).astype(np.uint64) if len(r) else np.zeros(0, np.uint64)
        parts.append(base + lows)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint64)
