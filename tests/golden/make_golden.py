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


if __name__ == "__main__":
    rows = combinations()
    with open(os.path.join(HERE, "container_combinations.json"), "w") as f:
        json.dump({"source": "roaring/roaring_internal_test.go:2974-3761 (TestContainerCombinations)", "rows": rows}, f, indent=0)
    shutil.copyfile(os.path.join(REF, "roaring/testdata/bitmapcontainer.roaringbitmap"),
                    os.path.join(HERE, "bitmapcontainer.roaringbitmap"))
    ops = {}
    for r in rows:
        ops[r["op"]] = ops.get(r["op"], 0) + 1
    print(len(rows), ops)
