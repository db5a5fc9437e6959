"""CPU-only: the C-ABI library builds, loads and exports every symbol include/fbgpu.h declares; host-side logic
(PQL mirror, program shapes, roaring_io) — no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from featurebase_b200 import build, executor as X, lib as L, pql, roaring_io
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    return build.build_fbgpu()


def test_header_symbols_exported(so):
    hdr = open(os.path.join(ROOT, "include", "fbgpu.h")).read()
    declared = set(re.findall(r"\b(fbgpu_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"fbgpu_last_error"} - {"fbgpu_last_error"}
    dll = ctypes.CDLL(so)
    missing = [s for s in sorted(declared) if not hasattr(dll, s)]
    assert not missing, missing
    assert dll.fbgpu_abi_version() == 2
    assert set(L.EXPORTS) <= declared


def test_no_oracle_linkage(so):
    """the product library must not reference the oracle"""
    data = open(so, "rb").read()
    assert b"fbo_" not in data and b"libfboracle" not in data
    for f in os.listdir(os.path.join(ROOT, "featurebase_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "featurebase_b200", f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f


def test_no_kernel_interpreter_in_product(so):
    """tests/emu/ (the CPU kernel interpreter) is test infrastructure: nothing the product ships may name it, and the shipped
    library is the nvcc build (it carries sm_100a device code, the interpreted build carries none)"""
    data = open(so, "rb").read()
    assert b"fbgpu_emu_switch" not in data and b"kernel emulation" not in data and b".nv_fatbin" in data
    for sub in ("featurebase_b200", os.path.join("featurebase_b200", "csrc"), "include", "."):
        for f in os.listdir(os.path.join(ROOT, sub)):
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(ROOT, sub, f)).read()
                assert "tests/emu" not in src and "FBGPU_EMU" not in src and "libfbgpu_emu" not in src, f


def test_init_without_gpu_fails_loudly(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.FbgpuError):
        L.Context(0)


def test_roaring_io_matches_oracle_writer():
    rng = np.random.default_rng(1)
    v = np.unique(np.concatenate([rng.integers(0, 1 << 21, 20000), np.arange(7 << 16, (7 << 16) + 50000),
                                  (3 << 16) + rng.choice(1 << 16, 30000, replace=False)])).astype(np.uint64)
    d = roaring_io.encode(v)
    assert d == O.Bitmap.from_values(v).to_bytes()
    assert np.array_equal(roaring_io.decode(d), v)


def test_pql_subset_parser():
    c = pql.parse("Count(Intersect(Row(f=10), Union(Row(g=11), Row(g=12))))")[0]
    assert c.name == "Count" and c.children[0].name == "Intersect" and c.children[0].children[1].children[1].args == {"g": 12}
    c = pql.parse("Row(-5 < v <= 10)")[0]
    assert c.args["v"].op == "><" and c.args["v"].value == [-4, 10]
    c = pql.parse("TopN(f, Row(other=10), n=5, ids=[0,10,30])")[0]
    assert c.args["_field"] == "f" and c.args["ids"] == [0, 10, 30] and c.children[0].name == "Row"


def test_bsigroup_base_value_rules():
    """field.go:2412-2463 edge rules (restated), foo in [-990, 1000] as in executor_test.go:3007"""
    f = X.Field(1, "foo", "int", min=-990, max=1000)
    assert f.base == 0 and f.bit_depth == 10
    assert f.base_value("<", 2000) == (1024, False)      # clamp to bitDepthMax, +1 for LT
    assert f.base_value(">", -5000) == (-1024, False)
    assert f.base_value(">", 5000)[1] and f.base_value("<", -5000)[1]
    assert f.base_value("==", 5000)[1]
    assert f.base_value_between(-5000, 5000) == (-1023, 1023, False)
    assert f.base_value_between(10, 5)[2]
    g = X.Field(2, "pos", "int", min=100, max=200)
    assert g.base == 100 and g.bit_depth == 7


def test_header_is_plain_c():
    """include/fbgpu.h is what cgo compiles: it must be valid C99 on its own (no C++-isms, every type declared)"""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "h.c")
        open(src, "w").write('#include "fbgpu.h"\nint main(void) { return sizeof(fbgpu_op) == 48 ? 0 : 1; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), src, "-o", os.path.join(d, "h")])
        assert subprocess.call([os.path.join(d, "h")]) == 0          # the 48-byte op layout INTEGRATION.md's Go struct mirrors
