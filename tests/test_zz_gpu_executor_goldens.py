"""GPU twin of tests/test_oracle.py::test_executor_bsi_goldens_through_host_mirror: the reference's end-to-end BSI
goldens (executor_test.go:3007-3289) through PQL -> host mirror -> C ABI -> CUDA kernels, compared with the literal
expected columns AND with the oracle's canonical bytes.  (Sorted last on purpose: written after the round's GPU budget
was spent, so it has only been exercised through its CPU twin.)"""
import numpy as np
import pytest

from featurebase_b200 import executor as X
from tests.golden import vectors as V
from tests.oracle_exec import Pair

pytestmark = pytest.mark.gpu


def test_executor_bsi_goldens_on_gpu():
    p = Pair(track_existence=True)
    p.field("f")
    for name, (lo, hi) in V.BSI_EXEC_SETUP["ranges"].items():
        p.field(name, "int", min=lo, max=hi, bit_depth=(63 if hi > (1 << 40) else None))
    for name, bits in V.BSI_EXEC_SETUP["set"].items():
        for r, c in bits:
            p.holder.set_bit("i", name, r, c)
    for name, vals in V.BSI_EXEC_SETUP["int"].items():
        for c, v in vals:
            p.holder.set_value("i", name, c, v)
    p.sync_pending()
    for q, exp in V.BSI_EXEC_CASES:
        got = p.check_row(q)                      # bytes == oracle canonical bytes
        assert [int(c) for c in got.columns()] == exp, q
    with pytest.raises(X.QueryError):
        p.ex.execute("i", "Row(bad_field >= 20)")
