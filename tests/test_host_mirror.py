"""Host-logic twin of the executor-level GPU tests (`-m "not gpu"`).

The bodies of the GPU tests that go PQL -> featurebase_b200.executor -> C ABI are re-run here with the library context
replaced by tests/oracle_ctx.OracleCtx, which interprets the very same fbgpu_op programs / row-count / group-by calls
with the CPU oracle.  What this pins on a box without a GPU: the PQL parser, the program the mirror emits for every
call shape (bsiGroup.baseValue clamping, null rows, whole-range short cuts, Not/All over the existence field), and the
TopN/TopK/GroupBy reductions — against the reference's literal expected results (executor_test.go cases in
tests/golden/vectors.py).  The CUDA kernels and the C-side program compiler are NOT exercised here; they are covered by
the `-m gpu` runs of the same bodies."""
import pytest

from featurebase_b200 import executor as X
from featurebase_b200 import lib as L
from tests import oracle_exec
from tests import test_gpu_parity as G
from tests import test_zz_gpu_executor_goldens as Z
from tests import test_zz_gpu_experimental as E
from tests.golden import vectors as V
from tests.oracle_ctx import OracleCtx


@pytest.fixture
def oracle_backed(monkeypatch):
    def make(*a, **kw):
        return oracle_exec.Pair(*a, ctx=OracleCtx(), **kw)
    monkeypatch.setattr(G, "Pair", make)
    monkeypatch.setattr(Z, "Pair", make)
    monkeypatch.setattr(E, "Pair", make)


def test_setop_goldens_and_edge_semantics(oracle_backed):
    G.test_executor_goldens_and_edge_semantics()
    G.test_config1_single_shard_plumbing()


def test_bsi_goldens(oracle_backed):
    G.test_bsi_range_goldens_on_gpu()
    Z.test_executor_bsi_goldens_on_gpu()


@pytest.mark.parametrize("signed", [False, True])
def test_bsi_diagonal(oracle_backed, signed):
    G.test_bsi_diagonal_exhaustive_on_gpu(signed)


def test_topk_topn_groupby(oracle_backed):
    G.test_executor_topk_topn_groupby_goldens_on_gpu()
    G.test_topk_topn_rowcounts()
    G.test_groupby_two_and_three_fields()


def test_bsi_aggregates(oracle_backed):
    """Sum / Min / Max: the mirror's whole-batch composition (counts per value row, bit sweep over the batch) against
    the reference's per-shard evaluation + ValCount reduce, and the literal expectations of executor_test.go"""
    E.test_bsi_aggregate_goldens()
    E.test_bsi_aggregates_random()


def test_fragment_top_goldens(oracle_backed):
    E.test_fragment_top_goldens()
    E.test_topn_cutoff_goldens()
    E.test_topn_cutoffs_random()
    E.test_row_counts_per_shard_entry_point()
    E.test_filter_sample_goldens()


def test_groupby_postprocessing_goldens(oracle_backed):
    E.test_groupby_postprocessing_goldens()


def test_time_quantum(oracle_backed):
    import datetime as dt
    from featurebase_b200 import timeq
    for a, b, q, exp in V.VIEWS_BY_TIME_RANGE:                   # time_internal_test.go:107-186, literal
        got = timeq.views_by_time_range("F", dt.datetime.strptime(a, "%Y-%m-%d %H:%M"), dt.datetime.strptime(b, "%Y-%m-%d %H:%M"), q)
        assert got == exp, (a, b, q)
    assert timeq.views_by_time("F", dt.datetime(2000, 1, 2, 3, 4), "YMDH") == ["F_2000", "F_200001", "F_20000102", "F_2000010203"]
    E.test_time_quantum_rows()


def test_embedded_rows(oracle_backed):
    E.test_embedded_rows_constrow_unionrows()
    E.test_shift_and_includes_column()
    E.test_all_with_limit_offset()
    E.test_min_max_row()


def test_various_queries(oracle_backed):
    E.test_various_queries_goldens()
    E.test_distinct_random()
    E.test_columns_entry_point()
    E.test_extract_entry_point()
    E.test_extract_table_golden()
    E.test_sort_goldens()
    E.test_field_value_and_options()
    E.test_topk_time_range()
    E.test_arena_compaction()
    E.test_incremental_container_refresh()


def test_percentile(oracle_backed):
    E.test_percentile_vs_reference_helper()


def test_bench_archetype_matrix_plumbing(oracle_backed):
    E.test_bench_archetype_matrix()
    E.test_kernel_table_goldens_on_device()
    E.test_mixed_container_goldens_on_device()
    E.test_bitmap_level_goldens_on_device()


def test_rbf_import_plumbing(oracle_backed):
    """Holder.import_rbf naming ("~field;view<") and the body of the opt-in GPU test, with the RBF bytes read by the
    product's reader (g++ harness) into an oracle-backed context"""
    E.test_rbf_loader_matches_fragment_loader()


def test_emitted_programs():
    """the exact programs the mirror hands to the C ABI for the BSI short cuts (executor.go:5249-5354)"""
    ctx = OracleCtx()
    h = X.Holder(ctx=ctx)
    idx = h.create_index("i", track_existence=True)
    idx.create_field("f")
    v = idx.create_field("v", "int", min=-990, max=1000)
    assert (v.base, v.bit_depth, v.bit_depth_min(), v.bit_depth_max()) == (0, 10, -1023, 1023)
    pos = idx.create_field("p", "int", min=100, max=227)
    assert (pos.base, pos.bit_depth) == (100, 7)                  # bsiBase: min > 0 -> base = min (field.go:2384)
    neg = idx.create_field("n", "int", min=-300, max=-10)
    assert (neg.base, neg.bit_depth) == (-10, 9)                  # max < 0 -> base = max
    ex = X.Executor(h)

    def prog(q):
        ctx.programs.clear()
        ex.execute("i", q, [0])
        return [(o.opcode, o.field, o.view, o.argc, o.a, o.b, o.lo, o.hi) for o in ctx.programs[-1]]

    not_null = (L.OP_ROW, v.id, X.VIEW_BSI, 0, 0, 0, 0, 0)
    assert prog("Row(v != null)") == [not_null]
    assert prog("Row(v == null)") == [(L.OP_ALL, 0, 0, 0, 0, 0, 0, 0), not_null, (L.OP_DIFFERENCE, 0, 0, 2, 0, 0, 0, 0)]
    assert prog("Row(v > 5)") == [(L.OP_BSI_RANGE, v.id, X.VIEW_BSI, 0, 10, L.CMP[">"], 5, 0)]
    assert prog("Row(v > 2000)") == [(L.OP_EMPTY, 0, 0, 0, 0, 0, 0, 0)]          # above bitDepthMax: out of range
    assert prog("Row(v > 1010)") == [(L.OP_BSI_RANGE, v.id, X.VIEW_BSI, 0, 10, L.CMP[">"], 1010, 0)]   # > max but inside depth
    assert prog("Row(v < 1001)") == [not_null]                                   # LT above max: whole range
    assert prog("Row(v <= 1000)") == [not_null]
    assert prog("Row(v >= -990)") == [not_null]
    assert prog("Row(v > -991)") == [not_null]
    assert prog("Row(v < -5000)") == [(L.OP_EMPTY, 0, 0, 0, 0, 0, 0, 0)]
    assert prog("Row(v != 5000)") == [not_null]                                  # NEQ out of range: everything not null
    assert prog("Row(v == 5000)") == [(L.OP_EMPTY, 0, 0, 0, 0, 0, 0, 0)]
    assert prog("Row(v >< [-2000,2000])") == [not_null]
    assert prog("Row(v >< [-2000,7])") == [(L.OP_BSI_RANGE, v.id, X.VIEW_BSI, 0, 10, L.CMP["><"], -1023, 7)]
    assert prog("Row(v >< [7,3])") == [(L.OP_EMPTY, 0, 0, 0, 0, 0, 0, 0)]
    assert prog("Row(p >= 130)") == [(L.OP_BSI_RANGE, pos.id, X.VIEW_BSI, 0, 7, L.CMP[">="], 30, 0)]   # base-adjusted
    assert prog("Row(n == -20)") == [(L.OP_BSI_RANGE, neg.id, X.VIEW_BSI, 0, 9, L.CMP["=="], -10, 0)]
    assert prog("Count(Not(Row(f=3)))") == [(L.OP_ROW, idx.fields["f"].id, 0, 0, 3, 0, 0, 0), (L.OP_NOT, 0, 0, 1, 0, 0, 0, 0)]
    with pytest.raises(X.QueryError):
        ex.execute("i", "Row(f > 3)")                    # condition on a set field
    with pytest.raises(X.QueryError):
        ex.execute("i", "Count()")
    with pytest.raises(X.QueryError):
        ex.execute("i", "Count(Row(f=1), Row(f=2))")
    with pytest.raises(X.QueryError):
        ex.execute("nope", "Row(f=1)")
