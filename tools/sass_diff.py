#!/usr/bin/env python
"""Which kernels of the current build differ, instruction for instruction, from the build of an earlier commit?
    python tools/sass_diff.py 5f6bb73          # 5f6bb73 = the last commit whose library ran on a B200 in round 1
Builds that commit's csrc/ in a scratch directory with the same nvcc line, dumps both libraries with cuobjdump -sass and
compares the instruction streams per kernel (addresses and encodings stripped; template arguments that did not exist yet
are matched by kernel name).  No GPU needed.  A kernel reported IDENTICAL is, bit for bit in its instructions, the one that
was parity-tested and timed on the device."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if m and cur:
            res[cur].append(m.group(1).strip())
    return res


def short(name):
    m = re.match(r"_ZN5fbgpu\d+([a-z_0-9]+?)(I(Lb[01])E)?E", name)
    return (m.group(1), m.group(3) or "") if m else (name, "")


def main():
    commit = sys.argv[1]
    tmp = tempfile.mkdtemp(prefix="sassdiff_")
    tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "featurebase_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    old = os.path.join(tmp, "libold.so")
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(tmp, "include"), "-o", old, os.path.join(tmp, "featurebase_b200/csrc/fbgpu.cu"), "-ldl"], check=True, stdout=subprocess.DEVNULL)
    a, b = kernels(old), kernels(os.path.join(ROOT, "featurebase_b200", "libfbgpu.so"))
    by_name = {}
    for k, v in a.items():
        by_name.setdefault(short(k)[0], []).append((short(k)[1], v))
    for k, v in b.items():
        name, targ = short(k)
        cands = by_name.get(name, [])
        hit = [x for x in cands if x[0] == targ] or [x for x in cands if not x[0] and targ in ("", "Lb0")]
        label = name + ("<%s>" % ("true" if targ == "Lb1" else "false") if targ else "")
        if not hit:
            print(f"NEW        {label:28s} {len(v):5d} instructions")
        else:
            print(f"{'IDENTICAL' if hit[0][1] == v else 'CHANGED  '}  {label:28s} {len(hit[0][1]):5d} -> {len(v):5d} instructions")


if __name__ == "__main__":
    main()
