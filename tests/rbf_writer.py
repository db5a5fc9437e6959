"""Minimal RBF (roaring b-tree file) writer — TEST INFRASTRUCTURE for featurebase_b200/csrc/rbf_reader.h.

The Go toolchain is absent, so the reference cannot produce RBF files here; this restates the on-disk layout of
rbf/rbf.go (page header :185-205, root records :229-287, leaf cells :586-594, branch cells :620-625, dataOffset :203,
meta page :120-141) and the cell-type choice of ConvertToLeafArgs (rbf/cursor.go:1299-1341: array > 4079 elements or
run > 2039 intervals => bitmap page).  Pinned against the reference's own fixture bytes
(rbf/testdata/check/bad-freelist/data, tests/golden/vectors.py:RBF_FIXTURE_PAGES) in tests/test_rbf.py.

Pages are packed the simple way (fill a leaf until the next cell does not fit); the reference's b-tree splits pages
differently while inserting, which changes which cells share a page but not the format a reader sees."""
import struct

import numpy as np

PAGE = 8192
MAGIC = b"\xffRBF"
T_ROOT, T_LEAF, T_BRANCH, T_BITMAP_HEADER = 1, 2, 4, 8
C_ARRAY, C_RLE, C_BITMAP_PTR = 1, 2, 4
ARRAY_MAX, RLE_MAX = 4079, 2039
MAX_BRANCH_CELLS = (PAGE - 10) // (2 + 16)


def align8(n):
    return (n + 7) & ~7


def data_offset(n):
    return align8(10 + 2 * n)


def cells_from_pilosa(data):
    """Pilosa roaring bytes (one fragment) -> [(key, 'array'|'run'|'bitmap', payload ndarray)] in key order"""
    magic, = struct.unpack_from("<H", data, 0)
    assert magic == 12348
    n, = struct.unpack_from("<I", data, 4)
    out = []
    for i in range(n):
        key, typ, n1 = struct.unpack_from("<QHH", data, 8 + 12 * i)
        off, = struct.unpack_from("<I", data, 8 + 12 * n + 4 * i)
        card = n1 + 1
        if typ == 1:
            out.append((key, "array", np.frombuffer(data, dtype="<u2", count=card, offset=off).copy()))
        elif typ == 2:
            out.append((key, "bitmap", np.frombuffer(data, dtype="<u8", count=1024, offset=off).copy()))
        else:
            rn, = struct.unpack_from("<H", data, off)
            out.append((key, "run", np.frombuffer(data, dtype="<u2", count=2 * rn, offset=off + 2).copy().reshape(-1, 2)))
    return out


def _to_bitmap(kind, payload):
    w = np.zeros(1024, dtype=np.uint64)
    if kind == "array":
        v = payload.astype(np.uint64)
        np.bitwise_or.at(w, (v >> np.uint64(6)).astype(np.int64), np.uint64(1) << (v & np.uint64(63)))
    else:
        for s, l in payload.tolist():
            v = np.arange(s, l + 1, dtype=np.uint64)
            np.bitwise_or.at(w, (v >> np.uint64(6)).astype(np.int64), np.uint64(1) << (v & np.uint64(63)))
    return w


def _popcount(words):
    return int(np.unpackbits(words.view(np.uint8)).sum())


class Writer:
    def __init__(self):
        self.pages = [None, None, None]          # 0 meta, 1 root records, 2 freelist (an empty leaf)
        self.roots = {}
        self.raw = set()                         # page numbers of raw bitmap pages (no page header)

    def _alloc(self, page=None, raw=False):
        self.pages.append(page)
        if raw:
            self.raw.add(len(self.pages) - 1)
        return len(self.pages) - 1

    def add_bitmap(self, name, containers):
        """containers: [(key, kind, payload)] ascending keys; returns the root page number"""
        cells = []
        for key, kind, payload in containers:
            if kind == "array" and len(payload) > ARRAY_MAX or kind == "run" and len(payload) > RLE_MAX:
                kind, payload = "bitmap", _to_bitmap(kind, payload)           # ConvertToLeafArgs cursor.go:1308,1326
            if kind == "array":
                cells.append((key, C_ARRAY, len(payload), len(payload), np.asarray(payload, dtype="<u2").tobytes()))
            elif kind == "run":
                p = np.asarray(payload, dtype="<u2").reshape(-1, 2)
                bit_n = int((p[:, 1].astype(np.int64) - p[:, 0].astype(np.int64) + 1).sum())
                cells.append((key, C_RLE, len(p), bit_n, p.tobytes()))
            else:
                words = np.asarray(payload, dtype="<u8")
                pg = self._alloc(words.tobytes(), raw=True)                    # raw bitmap page, no header
                cells.append((key, C_BITMAP_PTR, 0, _popcount(words), struct.pack("<I", pg)))
        # leaf level
        level = []                                                            # [(left key, pgno)]
        i = 0
        while True:
            take, used = [], 0
            while i < len(cells):
                sz = align8(18 + len(cells[i][4]))
                if data_offset(len(take) + 1) + used + sz > PAGE:
                    break
                take.append(cells[i])
                used += sz
                i += 1
            pg = self._alloc()
            self.pages[pg] = self._leaf_page(pg, take)
            level.append((take[0][0] if take else 0, pg))
            if i >= len(cells):
                break
        while len(level) > 1:                                                 # branch levels
            nxt = []
            for j in range(0, len(level), MAX_BRANCH_CELLS):
                grp = level[j:j + MAX_BRANCH_CELLS]
                pg = self._alloc()
                self.pages[pg] = self._branch_page(pg, grp)
                nxt.append((grp[0][0], pg))
            level = nxt
        self.roots[name] = level[0][1]
        return level[0][1]

    @staticmethod
    def _leaf_page(pgno, cells):
        p = bytearray(PAGE)
        struct.pack_into(">IIH", p, 0, pgno, T_LEAF, len(cells))
        off = data_offset(len(cells))
        for i, (key, typ, elem_n, bit_n, data) in enumerate(cells):
            struct.pack_into(">H", p, 10 + 2 * i, off)
            struct.pack_into("<QIH", p, off, key, typ, elem_n)
            struct.pack_into("<I", p, off + 14, bit_n)
            p[off + 18:off + 18 + len(data)] = data
            off += align8(18 + len(data))
        assert off <= PAGE
        return bytes(p)

    @staticmethod
    def _branch_page(pgno, children):
        p = bytearray(PAGE)
        struct.pack_into(">IIH", p, 0, pgno, T_BRANCH, len(children))
        off = data_offset(len(children))
        for i, (key, child) in enumerate(children):
            struct.pack_into(">H", p, 10 + 2 * i, off)
            struct.pack_into("<QII", p, off, key, 0, child)
            off += 16
        assert off <= PAGE
        return bytes(p)

    def finish(self, wal_id=0):
        """-> list of pages (bytes); page 0 is the meta page"""
        names = sorted(self.roots)
        root_pages, cur, buf = [1], 1, b""
        recs = {1: b""}
        for name in names:
            rec = struct.pack(">IH", self.roots[name], len(name.encode())) + name.encode()
            if 12 + len(recs[cur]) + len(rec) > PAGE:
                cur = self._alloc()
                root_pages.append(cur)
                recs[cur] = b""
            recs[cur] += rec
        for k, pg in enumerate(root_pages):
            p = bytearray(PAGE)
            nxt = root_pages[k + 1] if k + 1 < len(root_pages) else 0
            struct.pack_into(">III", p, 0, pg, T_ROOT, nxt)
            p[12:12 + len(recs[pg])] = recs[pg]
            self.pages[pg] = bytes(p)
        self.pages[2] = self._leaf_page(2, [])
        meta = bytearray(PAGE)
        meta[0:4] = MAGIC
        struct.pack_into(">IqII", meta, 8, len(self.pages), wal_id, 1, 2)
        self.pages[0] = bytes(meta)
        return list(self.pages)


def build(bitmaps, wal_id=0):
    """bitmaps: {name: [(key, kind, payload)]} -> data file bytes"""
    w = Writer()
    for name in sorted(bitmaps):
        w.add_bitmap(name, bitmaps[name])
    return b"".join(w.finish(wal_id))


def wal_between(old_pages, new_pages, raw_bitmap_pages=()):
    """WAL bytes that turn the data file `old_pages` into `new_pages`: every changed / added page (raw bitmap pages behind
    a bitmap-header page, rbf/db.go:329-338), then the new meta page as the commit marker."""
    out = []
    for pg in range(1, len(new_pages)):
        if pg < len(old_pages) and old_pages[pg] == new_pages[pg]:
            continue
        if pg in raw_bitmap_pages:
            hdr = bytearray(PAGE)
            struct.pack_into(">II", hdr, 0, pg, T_BITMAP_HEADER)
            out.append(bytes(hdr))
        out.append(new_pages[pg])
    out.append(new_pages[0])
    return b"".join(out)
