"""AddressSanitizer + UBSan fuzz of the host-only readers (csrc/rbf_reader.h, csrc/roaring_parse.h): tests/native/fuzz_asan.cpp
mutates valid files (byte flips biased to headers, truncations, WAL overlays made of arbitrary pages) and touches every byte
of every payload view the readers return.  Skipped when the sanitizer runtimes are not installed."""
import os
import subprocess
import tempfile

import pytest

from featurebase_b200 import datagen as D
from oracle import oracle as O
from tests import rbf_writer as W
from tests import test_rbf as TR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readers_under_asan_ubsan():
    tmp = tempfile.mkdtemp(prefix="fuzz_asan_")
    exe = os.path.join(tmp, "fuzz_asan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "featurebase_b200", "csrc"),
                        os.path.join(ROOT, "tests", "native", "fuzz_asan.cpp"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-200:])
    _, conts = TR._fragment_containers(70, 1)
    rbf = os.path.join(tmp, "f.rbf")
    open(rbf, "wb").write(W.build({"~f;standard<": conts, "~g;standard<": conts[:50]}))
    merged = O.Bitmap()
    for d in (D.fragment(7, 1, [0, 1], 0.01), D.fragment(7, 1, [2], 0.3), D.fragment(7, 1, [3], 0.2, mode=1, mean_run=200.0)):
        merged = merged.union(O.Bitmap.from_bytes(d))
    pil = os.path.join(tmp, "f.roaring")
    open(pil, "wb").write(merged.to_bytes())
    official = os.path.join(ROOT, "tests", "golden", "bitmapcontainer.roaringbitmap")
    for roaring_file in (pil, official):
        out = subprocess.run([exe, rbf, roaring_file, "800"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "fuzz_asan done" in out.stdout, out.stderr[-2000:]
