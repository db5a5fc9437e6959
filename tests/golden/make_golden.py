#!/usr/bin/env python3
"""Extracts the reference's literal golden vectors into JSON fixtures (run in the build container, where
/root/reference exists; the GPU box only sees the committed fixtures).

Sources (all under /root/reference):
  roaring/roaring_internal_test.go:2974-3761  TestContainerCombinations table  -> container_combinations.json
  roaring/testdata/bitmapcontainer.roaringbitmap (official-format file, 10,000 bits) -> bitmapcontainer.roaringbitmap
Only test DATA (op, archetype names, expected archetype) is extracted - no reference code.
"""
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def combinations():
    lines = open(os.path.join(REF, "roaring/roaring_internal_test.go")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("func TestContainerCombinations"))
    end = next(i for i in range(start, len(lines)) if lines[i].startswith("	for _, testOp := range testOps"))
    pat = re.compile(r'^\s*\{(\w+),\s*"(\w+)",\s*"(\w+)",\s*"(\w+)"\},')
    pat1 = re.compile(r'^\s*\{(\w+),\s*"(\w+)",\s*"",\s*"(\w+)"\},')
    rows = []
    for i in range(start, end):
        m = pat.match(lines[i])
        if m:
            rows.append({"op": m.group(1), "x": m.group(2), "y": m.group(3), "exp": m.group(4), "line": i + 1})
            continue
        m = pat1.match(lines[i])
        if m:
            rows.append({"op": m.group(1), "x": m.group(2), "y": "", "exp": m.group(3), "line": i + 1})
    return rows


KERNEL_TABLE_FUNCS = ["TestIntersectArrayRun", "TestIntersectRunRun", "TestUnionInterval16InPlace", "TestUnionRunRun", "TestUnionArrayRun",
                      "TestDifferenceArrayRun", "TestDifferenceRunArray", "TestDifferenceRunRun", "TestXorArrayRun", "TestXorRunRun"]


def _parse_literal(text):
    """[]uint16{..} | []uint16(nil) | []Interval16{{Start: a, Last: b}|{a, b}, ..} | NewContainerArray(..) | NewContainerRun(..) | int"""
    text = text.strip().rstrip(",")
    m = re.match(r"^NewContainer(Array|Run)\((.*)\)$", text)
    if m:
        return _parse_literal(m.group(2))
    if text.startswith("[]uint16"):
        body = text[len("[]uint16"):]
        if body.startswith("(nil)") or body in ("{}",):
            return {"kind": "array", "values": []}
        return {"kind": "array", "values": [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]}
    if text.startswith("[]Interval16"):
        body = text[len("[]Interval16"):]
        if body.startswith("(nil)") or body in ("{}",):
            return {"kind": "runs", "values": []}
        pairs = re.findall(r"\{(?:Start:\s*)?(\d+),\s*(?:Last:\s*)?(\d+)\}", body)
        return {"kind": "runs", "values": [[int(a), int(b)] for a, b in pairs]}
    m = re.match(r"^(?:bitmap|Make)(LastBitSet|Full|Empty|OddBitsSet|EvenBitsSet|FirstBitSet)\(\)$", text)
    if m:                                                        # helper-built bitmaps of roaring_helpers_test.go:77-135: named by archetype
        name = m.group(1)
        return {"kind": "archetype", "values": name[0].lower() + name[1:]}
    m = re.match(r"^MakeBitmap\((.*)\)$", text)
    if m:
        return _parse_literal(m.group(1))
    m = re.match(r"^(?:\[\]uint64|\[bitmapN\]uint64)(.*)$", text)
    if m:                                                        # leading words of a 1024-word bitmap
        return {"kind": "bitmap", "values": [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]}
    if re.match(r"^-?\d+$", text):
        return {"kind": "int", "values": int(text)}
    return None


# second batch: table tests whose operands include bitmap words / ranges / conversions (same literal-table shape)
KERNEL_TABLE_FUNCS2 = ["TestIntersectBitmapRunBitmap", "TestIntersectBitmapRunArray", "TestUnionBitmapRun", "TestDifferenceRunBitmap", "TestDifferenceBitmapRun",
                       "TestDifferenceBitmapArray", "TestDifferenceBitmapBitmap", "TestXorBitmapRun", "TestIntersectArrayBitmap", "TestIntersectionCountArrayBitmap2",
                       "TestBitmapCountRuns", "TestArrayCountRuns", "TestArrayToBitmap", "TestBitmapToArray", "TestRunToBitmap", "TestBitmapToRun", "TestArrayToRun",
                       "TestRunToArray", "TestBitmapSetRange", "TestBitmapZeroRange", "TestBitmapXorRange"]


def kernel_tables(funcs=None):
    lines = open(os.path.join(REF, "roaring/roaring_internal_test.go")).read().split("\n")
    out = []
    for fn in (funcs or KERNEL_TABLE_FUNCS):
        start = next(i for i, l in enumerate(lines) if l.startswith(f"func {fn}("))
        end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("func "))
        cur, cur_line = {}, None
        for i in range(start, end):
            m = re.match(r"^\s*(\w+):\s*(.+?),?\s*(?://.*)?$", lines[i])
            if m and m.group(1) != "name":
                lit = _parse_literal(m.group(2))
                if lit is not None:
                    if not cur:
                        cur_line = i + 1
                    cur[m.group(1)] = lit
            elif re.match(r"^\s*\},?\s*(\{\s*)?$", lines[i]) and cur:
                out.append({"func": fn, "line": cur_line, "fields": cur})
                cur = {}
    return out


if __name__ == "__main__":
    kt = kernel_tables()
    with open(os.path.join(HERE, "kernel_tables.json"), "w") as f:
        json.dump({"source": "roaring/roaring_internal_test.go (table tests listed in KERNEL_TABLE_FUNCS)", "cases": kt}, f, indent=0)
    byf = {}
    for c in kt:
        byf[c["func"]] = byf.get(c["func"], 0) + 1
    print(len(kt), byf)
    kt2 = kernel_tables(KERNEL_TABLE_FUNCS2)
    with open(os.path.join(HERE, "kernel_tables2.json"), "w") as f:
        json.dump({"source": "roaring/roaring_internal_test.go (table tests listed in KERNEL_TABLE_FUNCS2)", "cases": kt2}, f, indent=0)
    byf = {}
    for c in kt2:
        byf[c["func"]] = byf.get(c["func"], 0) + 1
    print(len(kt2), byf)
    rows = combinations()
    with open(os.path.join(HERE, "container_combinations.json"), "w") as f:
        json.dump({"source": "roaring/roaring_internal_test.go:2974-3761 (TestContainerCombinations)", "rows": rows}, f, indent=0)
    shutil.copyfile(os.path.join(REF, "roaring/testdata/bitmapcontainer.roaringbitmap"),
                    os.path.join(HERE, "bitmapcontainer.roaringbitmap"))
    ops = {}
    for r in rows:
        ops[r["op"]] = ops.get(r["op"], 0) + 1
    print(len(rows), ops)
