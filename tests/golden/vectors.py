"""Literal golden vectors restated by hand from the reference's test files (paths under /root/reference).
Each entry cites the test it comes from.  DATA ONLY."""

# roaring/roaring_internal_test.go:259-282 TestBitmapCountRange: (start, end, first words, expected)
BITMAP_COUNT_RANGE = [
    (0, 1, [1], 1),
    (2, 7, [0xFFFFFFFFFFFFFF18], 2),
    (67, 68, [0, 0x8], 1),
    (1, 68, [0x3, 0x8, 0xF], 2),
    (1, 258, [0xF, 0x8, 0xA, 0x4, 0xFFFFFFFFFFFFFFFF], 9),
    (66, 71, [0xF, 0xFFFFFFFFFFFFFF18], 2),
    (63, 64, [0x8000000000000000], 1),
]

# roaring_internal_test.go:398-406 TestIntersectionCountArrayRun
ICOUNT_ARRAY_RUN = [([1, 5, 10, 11, 12], [(2, 10), (12, 13), (15, 16)], 3)]

# roaring_internal_test.go:408-426 TestIntersectionCountBitmapRun: (bitmap words, runs, expected)
ICOUNT_BITMAP_RUN = [
    ([1 << 63], [(63, 64)], 1),
    ([0xF0000001, 0xFF00000000000000, 0xFF000000000000F0, 0x0F0000], [(29, 31), (125, 134), (191, 197), (200, 300)], 14),
]

# roaring_internal_test.go:428-473 TestIntersectionCountRunRun
ICOUNT_RUN_RUN = [
    ([], [(3, 8)], 0),
    ([(2, 10)], [(3, 8)], 6),
    ([(2, 10)], [(1, 11)], 9),
    ([(2, 10)], [(0, 2)], 1),
    ([(2, 10)], [(1, 10)], 9),
    ([(2, 10)], [(5, 12)], 6),
    ([(2, 10)], [(10, 99)], 1),
    ([(2, 10), (44, 99)], [(12, 14)], 0),
    ([(2, 10), (12, 13)], [(2, 10), (12, 13)], 11),
    ([(8, 12), (15, 19)], [(9, 9), (11, 17)], 6),
]

# roaring_internal_test.go:3793-3813 TestUnmarshalRoaringWithNoErrors (official RoaringBitmap format images)
OFFICIAL_HEX = [
    ("3A300000020000000000020001000000180000001E0000000100020003000100", [1, 2, 3, 65537]),
    ("3B3001000100000900010000000100010009000100", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 65537]),
]
OFFICIAL_FILE = ("bitmapcontainer.roaringbitmap", 10000)
# roaring_internal_test.go:3839-3853 TestUnmarshalRoaringWithErrors: zero-container images
OFFICIAL_ZERO_CONTAINER_ERRORS = ["3A30000000000000", "3B30000000000000"]
PILOSA_EMPTY_OK = "3C30000000000000"

# fragment_internal_test.go:606-916 TestFragment_Range: (values {col: value}, bitDepth, [(op, predicate(s), expected cols)])
BSI_RANGE_CASES = [
    ({1000: 382, 2000: 300, 3000: 2818, 4000: 300}, 16, [("==", 300, [2000, 4000]), ("!=", 300, [1000, 3000])]),
    ({1000: 0, 2000: 1}, 1, [("==", 3, []), ("==", 4, [])]),  # EQOversizeRegression
    ({1000: 382, 2000: 300, 3000: 2817, 4000: 301, 5000: 1, 6000: 0}, 16, [
        ("<", 301, [2000, 5000, 6000]), ("<", 300, [5000, 6000]),
        ("<=", 301, [2000, 4000, 5000, 6000]), ("<=", 300, [2000, 5000, 6000]),
        (">", 300, [1000, 3000, 4000]), (">", 301, [1000, 3000]),
        (">=", 300, [1000, 2000, 3000, 4000]), (">=", 301, [1000, 3000, 4000]),
        ("><", (300, 2817), [1000, 2000, 3000, 4000]), ("><", (301, 2817), [1000, 3000, 4000]),
        ("><", (301, 2816), [1000, 4000]), ("><", (300, 2816), [1000, 2000, 4000]),
    ]),
    ({1: 1}, 1, [("<", 2, [1])]),                      # LTRegression
    ({1: 3, 2: 0}, 2, [("<", 3, [2])]),                # LTMaxRegression
    ({1: 0, 2: 1}, 2, [(">", 0, [2])]),                # GTMinRegression
    ({1: 0, 2: 1}, 2, [(">", 4, [])]),                 # GTOversizeRegression
    ({1: 0xf0, 2: 0xf1}, 64, [("><", (0xf0, 0xf1), [1, 2])]),  # BetweenCommonBitsRegression
]

# executor_test.go:1236-1373 set-op end-to-end goldens. Field "general", ShardWidth = 2^20.
SW = 1 << 20
EXEC_SETOPS = {
    # TestExecutor_Execute_Difference: Set(1,general=10) Set(2,general=10) Set(3,general=10) Set(2,general=11) Set(4,general=11)
    "difference": ({10: [1, 2, 3], 11: [2, 4]}, "Difference(Row(general=10), Row(general=11))", [1, 3]),
    # TestExecutor_Execute_Intersect: row10 = {1, SW+1, SW+2}, row11 = {1, 2, SW+2}
    "intersect": ({10: [1, SW + 1, SW + 2], 11: [1, 2, SW + 2]}, "Intersect(Row(general=10), Row(general=11))", [1, SW + 2]),
    # TestExecutor_Execute_Union: row10 = {0, SW+1, SW+2}, row11 = {2, SW+2}
    "union": ({10: [0, SW + 1, SW + 2], 11: [2, SW + 2]}, "Union(Row(general=10), Row(general=11))", [0, 2, SW + 1, SW + 2]),
    # TestExecutor_Execute_Xor: row10 = {0, SW+1, SW+2}, row11 = {2, SW+2}
    "xor": ({10: [0, SW + 1, SW + 2], 11: [2, SW + 2]}, "Xor(Row(general=10), Row(general=11))", [0, 2, SW + 1]),
    # TestExecutor_Execute_Count: row10 = {3, SW+1, SW+2} -> Count = 3
    "count": ({10: [3, SW + 1, SW + 2]}, "Count(Row(general=10))", 3),
}

# executor_test.go:1758-1809 TestExecutor_ExecuteTopK: bits (row, col); TopK(f, k=2) -> [(10,4),(0,3)]
TOPK_BITS = [(0, 0), (0, SW + 2), (10, 2), (10, SW), (10, 2 * SW), (10, SW + 1), (20, SW), (0, 1)]
TOPK_EXPECT = [(10, 4), (0, 3)]
# executor_test.go:1846-1889 TestExecutor_Execute_TopN: TopN(f, n=2) -> [(0,5),(10,2)]
TOPN_BITS = [(0, 0), (0, 1), (0, SW), (0, SW + 2), (0, 5 * SW + 100), (10, 0), (10, SW), (20, SW)]
TOPN_EXPECT = [(0, 5), (10, 2)]
# executor_test.go:6033-6120 TestExecutor_Execute_GroupBy: Basic / Filter
GROUPBY_GENERAL = [(10, 0), (10, 1), (10, SW + 1), (11, 2), (11, SW + 2), (12, 2), (12, SW + 2)]
GROUPBY_SUB = [(100, 0), (100, 1), (100, 3), (100, SW + 1), (110, 2), (110, 0)]
GROUPBY_BASIC = [((10, 100), 3), ((10, 110), 1), ((11, 110), 1), ((12, 110), 1)]
GROUPBY_FILTER_GENERAL_10 = [((10, 100), 3), ((10, 110), 1)]

# executor_test.go:3007-3289 TestExecutor_Execute_Row_BSIGroup: fields foo in [-990,1000], other/bar int64, edge in [-900,1000];
# existence tracking on.  (query, expected columns)
BSI_EXEC_SETUP = {
    "set": {"f": [(0, 0), (0, SW + 1)]},
    "int": {"foo": [(50, 20), (SW, 30), (SW + 2, 10), (5 * SW + 100, 20), (SW + 1, 60)], "bar": [(50, 2000)], "other": [(0, 1000)],
            "edge": [(0, 100), (1, -100)]},
    "ranges": {"foo": (-990, 1000), "bar": (-(1 << 63), (1 << 63) - 1), "other": (-(1 << 63), (1 << 63) - 1), "edge": (-900, 1000)},
}
BSI_EXEC_CASES = [
    ("Row(other == null)", [1, 50, SW, SW + 1, SW + 2, 5 * SW + 100]),
    ("Row(foo == 20)", [50, 5 * SW + 100]),
    ("Row(other != null)", [0]),
    ("Row(foo != 20)", [SW, SW + 1, SW + 2]),
    ("Row(other != -20)", [0]),
    ("Row(foo < 20)", [SW + 2]),
    ("Row(foo <= 20)", [50, SW + 2, 5 * SW + 100]),
    ("Row(foo > 20)", [SW, SW + 1]),
    ("Row(foo >= 20)", [50, SW, SW + 1, 5 * SW + 100]),
    ("Row(0 <= other <= 1000)", [0]),
    ("Row(foo == 0)", []),
    ("Row(foo == 200)", []),
    ("Row(edge < 200)", [0, 1]),
    ("Row(edge > -1000)", [0, 1]),
]

# ---------------------------------------------------------------------------------------------------
# BSI aggregates.  fragment_internal_test.go:452-521 (TestFragment_Sum: bitDepth 16; the ClearValue step is the same
# fragment without column 1000) and :524-603 (TestFragment_MinMax).  Entries: (filter columns or None, value, count).
# ---------------------------------------------------------------------------------------------------
FRAG_SUM_VALUES = {1000: 382, 2000: 300, 2500: -600, 3000: 2818, 4000: 300}
FRAG_SUM_CASES = [(None, 382 + 300 - 600 + 2818 + 300, 5), ([2000, 4000, 5000], 300 + 300, 2)]
FRAG_SUM_CLEARED = (1000, 3800 - 382 - 600, 4)          # after clearValue(1000): sum 2818, count 4
FRAG_MINMAX_VALUES = {1000: 382, 2000: 300, 3000: 2818, 4000: 300, 5000: 2818, 6000: 2817, 7000: 0}
FRAG_MIN_CASES = [(None, 0, 1), ([2000, 4000, 5000], 300, 2), ([2000, 4000], 300, 2), ([1], 0, 0), ([1000], 382, 1), ([7000], 0, 1)]
FRAG_MAX_CASES = [(None, 2818, 2), ([2000, 4000, 5000], 2818, 1), ([2000, 4000], 300, 2), ([1], 0, 0), ([1000], 382, 1), ([7000], 0, 1)]

# executor_test.go:2192-2286 (MinMax WithOffset/Int): field range (min, max), one value set at column 10 -> Min == Max == (value, 1)
EXEC_MINMAX_OFFSET = [(10, 20, 11), (-10, 20, 11), (-10, 20, -9), (-20, -10, -11)]
# executor_test.go:2508-2567 (MinMax ColumnID): set field x, int field f in [-1100, 1000]; Min cases :2545-2567.  The Max
# expectations are the ColumnKey twin's (:2629-2655): same values and filters, columns named by keys instead of ids.
EXEC_MINMAX_SETUP = {
    "set": {"x": [(0, 0), (0, 3), (0, SW + 1), (1, 1), (2, SW + 2)]},
    "int": {"f": [(0, 20), (1, -5), (2, -5), (3, 10), (SW, 30), (SW + 2, 40), (5 * SW + 100, 50), (SW + 1, 60)]},
    "ranges": {"f": (-1100, 1000)},
}
EXEC_MIN_CASES = [("Min(field=f)", (-5, 2)), ("Min(Row(x=0), field=f)", (10, 1)), ("Min(Row(x=1), field=f)", (-5, 1)), ("Min(Row(x=2), field=f)", (40, 1))]
EXEC_MAX_CASES = [("Max(field=f)", (60, 1)), ("Max(Row(x=0), field=f)", (60, 1)), ("Max(Row(x=1), field=f)", (-5, 1)), ("Max(Row(x=2), field=f)", (40, 1))]
# executor_test.go:2782-2869 (Sum ColumnID / Integer)
EXEC_SUM_SETUP = {
    "set": {"x": [(0, 0), (0, SW + 1)]},
    "int": {"foo": [(0, 20), (SW, 30), (SW + 2, 40), (5 * SW + 100, 50), (SW + 1, 60)], "bar": [(0, 2000)], "other": [(0, 1000)]},
    "ranges": {"foo": (-990, 1000), "bar": (-(1 << 63), (1 << 63) - 1), "other": (-(1 << 63), (1 << 63) - 1)},
}
EXEC_SUM_CASES = [("Sum(field=foo)", (200, 5)), ('Sum(field="foo")', (200, 5)), ("Sum(foo)", (200, 5)), ("Sum(Row(x=0), field=foo)", (80, 2)),
                  ("Sum(foo, Row(x=0))", (80, 2)), ("Sum(field=bar)", (2000, 1)), ("Sum(Row(x=1), field=foo)", (0, 0))]

# ---------------------------------------------------------------------------------------------------
# RBF fixtures: the non-zero prefix of every 8 KiB page of rbf/testdata/check/bad-freelist/data and .../bad-bitmap/data
# (used by rbf/tx_test.go:1280-1305; 4 pages each, the rest of each page is zero).  One bitmap "x" holding the single
# bit 100 (leaf cell key 0, array, elemN 1, bitN 1).  In "bad-freelist" only the freelist page (2) is damaged (flags 4
# instead of 2), the bitmap itself is intact; in "bad-bitmap" the bitmap's leaf page (3) carries branch flags.
# ---------------------------------------------------------------------------------------------------
RBF_FIXTURE_PAGES = {
    "bad-freelist": ["ff524246000000000000000400000000000000040000000100000002", "00000001000000010000000000000003000178",
                     "0000000200000004", "0000000300000002000100100000000000000000000000000100000001000100000064"],
    "bad-bitmap": ["ff524246000000000000000400000000000000040000000100000002", "00000001000000010000000000000003000178",
                   "0000000200000002", "0000000300000004000100100000000000000000000000000100000001000100000064"],
}

# ---------------------------------------------------------------------------------------------------
# fragment.top exact cases (the rank cache holds every row at these sizes).  fragment_internal_test.go:1150-1171 (Top),
# :1174-1199 (TopN_Intersect, Src = columns 1,2,3), :1202-1249 (TopN_Intersect_Large: row i holds columns 0..i-1 for
# i < 1000, Src = columns 980..999), :1252-1272 (TopN_IDs), :1513-1537 (Zero_Tanimoto == no threshold).
# (rows {row: [cols]}, Src columns or None, n, ids or None, expected pairs)
# ---------------------------------------------------------------------------------------------------
FRAG_TOP_CASES = [
    ({100: [1, 3, 200], 101: [1], 102: [1, 2]}, None, 2, None, [(100, 3), (102, 2)]),
    ({100: [1, 10, 11, 12], 101: [1, 2, 3, 4], 102: [1, 2, 4, 5, 6], 103: [1000, 1001, 1002]}, [1, 2, 3], 3, None, [(101, 3), (102, 2), (100, 1)]),
    ("large", list(range(980, 1000)), 10, None, [(999 - k, 19 - k) for k in range(10)]),
    ({100: [1, 2, 3], 101: [4, 5, 6, 7], 102: [8, 9, 10, 11, 12]}, None, 0, [100, 101, 200], [(101, 4), (100, 3)]),
    ({100: [1, 3, 2, 200], 101: [1, 3], 102: [1, 2, 10, 12]}, [1, 2, 3], 0, None, [(100, 3), (101, 2), (102, 2)]),
]

# fragment.top with cut-offs: fragment_internal_test.go:1490-1511 (TestFragment_Tanimoto: threshold 50, Src = columns 1,2,3 ->
# rows 100 (coefficient 75) and 101 (67) stay, 102 (40) goes) and :1514-1537 (threshold 0 == no cut-off).  The MinThreshold
# rows are worked from fragment.go:1357-1362,1384-1388 on the same bits (cnt = 4, 2, 4; |Src ∩ row| = 3, 2, 2).
# (rows, Src columns or None, n, ids or None, MinThreshold, TanimotoThreshold, expected pairs)
FRAG_TOP_THRESHOLD_ROWS = {100: [1, 3, 2, 200], 101: [1, 3], 102: [1, 2, 10, 12]}
FRAG_TOP_THRESHOLD_CASES = [
    (FRAG_TOP_THRESHOLD_ROWS, [1, 2, 3], 0, None, 0, 50, [(100, 3), (101, 2)]),
    (FRAG_TOP_THRESHOLD_ROWS, [1, 2, 3], 0, None, 0, 0, [(100, 3), (101, 2), (102, 2)]),
    (FRAG_TOP_THRESHOLD_ROWS, None, 0, None, 3, 0, [(100, 4), (102, 4)]),                 # cnt < MinThreshold: row 101 out
    (FRAG_TOP_THRESHOLD_ROWS, [1, 2, 3], 0, None, 3, 0, [(100, 3)]),                      # 101 out on cnt, 102 out on |Src ∩ row| = 2
    (FRAG_TOP_THRESHOLD_ROWS, [1, 2, 3], 0, [100, 101, 102], 2, 0, [(100, 3), (101, 2), (102, 2)]),
    (FRAG_TOP_THRESHOLD_ROWS, None, 0, None, 0, 50, [(100, 4), (102, 4), (101, 2)]),      # Tanimoto without Src: ignored (:1333)
]

# ---------------------------------------------------------------------------------------------------
# roaring/filter_internal_test.go:24-41 sample fragment: for every slot i in 1..15 the column (i << 16) + i is set in
# rows 0, i, 2i, ... < 100.  :78-86 TestBaseFilter (all rows 0..99 present), :88-99 TestColumnFilter (rows holding
# column (i<<16)+i are the multiples of i), :101-114 TestRowsFilter (row set {0,1,2,3} ∩ rows holding column (2<<16)+2 =
# {0, 2}; limit 1 -> {0}), :116-138 TestRowsUnion (rows 7 ∪ 11 = columns {1<<16+1, 7<<16+7, 11<<16+11}, also in shard 2).
# ---------------------------------------------------------------------------------------------------
FILTER_SAMPLE_ROWS = 100


def filter_sample_bits():
    """[(row, column)]"""
    return [(row, (i << 16) + i) for i in range(1, 16) for row in range(0, FILTER_SAMPLE_ROWS, i)]


FILTER_ROWS_UNION = ([7, 11], [(1 << 16) + 1, (7 << 16) + 7, (11 << 16) + 11])
FILTER_ROWSET = ([0, 1, 2, 3], (2 << 16) + 2, [0, 2])

# ---------------------------------------------------------------------------------------------------
# executor_test.go:6033-6386 TestExecutor_Execute_GroupBy beyond Basic / Filter: aggregate=Sum (:6121-6129; int field v:
# column 0 -> 10, 1 -> 100, SW+10 -> 100), previous / limit paging (:6164-6181), "tricky data" (:6194-6201), wrapping
# iterators (:6203-6258), rows in different shards (:6260-6294, 6321-6331), paging through 64 groups (:6335-6386).
# Bits are (row, column); results ((row ids...), count[, agg]).
# ---------------------------------------------------------------------------------------------------
GB_V_VALUES = [(0, 10), (1, 100), (SW + 10, 100)]
GB_FIELDS = {
    "a": [(0, 1), (1, SW + 1)], "b": [(0, SW + 1), (1, 1)],
    "wa": [(0, 0), (0, 1), (0, 2), (1, 1), (2, 0), (2, 2), (3, 3)],
    "ma": [(0, 0), (1, SW), (2, 0), (3, SW)],
    "na": [(0, 0), (0, SW), (1, 0), (1, SW)],
    "ppa": [(0, 0), (1, 0), (2, 0), (3, 0), (3, 91000), (3, SW), (3, 2 * SW), (3, 3 * SW)],
}
for _src, _dsts in (("wa", ("wb", "wc")), ("ma", ("mb",)), ("na", ("nb",)), ("ppa", ("ppb", "ppc"))):
    for _d in _dsts:
        GB_FIELDS[_d] = GB_FIELDS[_src]
GB_CASES = [
    ("GroupBy(Rows(field=general), Rows(sub))", GROUPBY_BASIC),                                   # BasicLegacy
    ("GroupBy(Rows(general), Rows(sub), aggregate=Sum(field=v))", [((10, 100), 2, 110), ((10, 110), 1, 10)]),
    ("GroupBy(Rows(general, previous=10))", [((11,), 2), ((12,), 2)]),
    ("GroupBy(Rows(general, previous=10), limit=1)", [((11,), 2)]),
    ("GroupBy(Rows(a), Rows(b), limit=1)", [((0, 1), 1)]),                                           # tricky data
    ("GroupBy(Rows(wa), Rows(wb), Rows(wc, previous=1), limit=3)", [((0, 0, 2), 2), ((0, 1, 0), 1), ((0, 1, 1), 1)]),
    ("GroupBy(Rows(wa, previous=3), Rows(wb, previous=3), Rows(wc, previous=3), limit=3)", []),
    ("GroupBy(Rows(wa), Rows(wb, previous=2), Rows(wc, previous=2), limit=1)", [((1, 0, 0), 1)]),
    ("GroupBy(Rows(ma), Rows(mb), limit=5)", [((0, 0), 1), ((0, 2), 1), ((1, 1), 1), ((1, 3), 1), ((2, 0), 1)]),
    ("GroupBy(Rows(ma), Rows(mb, limit=2), limit=5)", [((0, 0), 1), ((1, 1), 1), ((2, 0), 1), ((3, 1), 1)]),
    ("GroupBy(Rows(na), Rows(nb))", [((0, 0), 2), ((0, 1), 2), ((1, 0), 2), ((1, 1), 2)]),
]
GB_PAGING_EXPECT = [((i // 16, (i % 16) // 4, i % 4), 5 if i == 63 else 1) for i in range(64)]

# ---------------------------------------------------------------------------------------------------
# time_internal_test.go:107-186 TestViewsByTimeRange: (start, end, quantum, views for name "F")
# ---------------------------------------------------------------------------------------------------
VIEWS_BY_TIME_RANGE = [
    ("2000-01-01 00:00", "2002-01-01 00:00", "Y", ["F_2000", "F_2001"]),
    ("2000-11-01 00:00", "2003-03-01 00:00", "YM", ["F_200011", "F_200012", "F_2001", "F_2002", "F_200301", "F_200302"]),
    ("2001-10-31 00:00", "2003-04-01 00:00", "YM", ["F_200110", "F_200111", "F_200112", "F_2002", "F_200301", "F_200302", "F_200303"]),
    ("1999-12-31 00:00", "2000-04-01 00:00", "YM", ["F_199912", "F_200001", "F_200002", "F_200003"]),
    ("2000-01-31 00:00", "2001-04-01 00:00", "YM", ["F_2000", "F_200101", "F_200102", "F_200103"]),
    ("2000-11-28 00:00", "2003-03-02 00:00", "YMD", ["F_20001128", "F_20001129", "F_20001130", "F_200012", "F_2001", "F_2002", "F_200301", "F_200302", "F_20030301"]),
    ("2000-11-28 22:00", "2002-03-01 03:00", "YMDH", ["F_2000112822", "F_2000112823", "F_20001129", "F_20001130", "F_200012", "F_2001", "F_200201", "F_200202",
                                                       "F_2002030100", "F_2002030101", "F_2002030102"]),
    ("2000-01-01 00:00", "2000-03-01 00:00", "M", ["F_200001", "F_200002"]),
    ("2000-11-29 00:00", "2002-02-03 00:00", "MD", ["F_20001129", "F_20001130", "F_200012"] + ["F_2001%02d" % m for m in range(1, 13)] + ["F_200201", "F_20020201", "F_20020202"]),
    ("2000-11-29 22:00", "2002-03-02 03:00", "MDH", ["F_2000112922", "F_2000112923", "F_20001130", "F_200012"] + ["F_2001%02d" % m for m in range(1, 13)]
     + ["F_200201", "F_200202", "F_20020301", "F_2002030200", "F_2002030201", "F_2002030202"]),
    ("2000-01-01 00:00", "2000-01-04 00:00", "D", ["F_20000101", "F_20000102", "F_20000103"]),
    ("2000-01-01 22:00", "2000-03-01 02:00", "DH", ["F_2000010122", "F_2000010123"] + ["F_200001%02d" % d for d in range(2, 32)] + ["F_200002%02d" % d for d in range(1, 30)]
     + ["F_2000030100", "F_2000030101"]),
    ("2000-01-01 00:00", "2000-01-01 02:00", "H", ["F_2000010100", "F_2000010101"]),
]
# executor_test.go:470-515 (quantum YMDH) and :982-1010 (quantum YMD): timestamped Set()s of field f, then Row(f=1, from, to)
TIME_BITS = [(1, 2, "1999-12-31T00:00"), (1, 3, "2000-01-01T00:00"), (1, 4, "2000-01-02T00:00"), (1, 5, "2000-02-01T00:00"), (1, 6, "2001-01-01T00:00"),
             (1, 7, "2002-01-01T02:00"), (1, 2, "1999-12-30T00:00"), (1, 2, "2002-02-01T00:00"), (10, 2, "2001-01-01T00:00")]
TIME_ROW_CASES = {
    "YMDH": [("Row(f=1, from=1999-12-31T00:00, to=2002-01-01T03:00)", [2, 3, 4, 5, 6, 7]), ("Row(f=1, from=1999-12-31T00:00)", [2, 3, 4, 5, 6, 7]),
             ("Row(f=1, to=2002-01-01T02:00)", [2, 3, 4, 5, 6]), ("Row(f=1, from=946598400, to=1009854000)", [2, 3, 4, 5, 6, 7])],
    "YMD": [("Row(f=1, from=1999-12-31T00:00, to=2003-01-01T03:00)", [2, 3, 4, 5, 6, 7]), ("Row(f=1, from=2002-01-01T00:00, to=2002-01-02T00:00)", [7]),
            ("Row(f=10, from=1999-12-31T00:00, to=2003-01-01T03:00)", [2])],
}

# ---------------------------------------------------------------------------------------------------
# roaring/roaring_test.go Bitmap-level literal cases (multi-container operands).  Operand specs: ("vals", [...]),
# ("range", start, stop, step), ("cat", spec, spec, ...).  (cite, a, b, op, ("count", n) | ("slice", [...]))
# ---------------------------------------------------------------------------------------------------
_EVEN10K = ("range", 0, 10000, 2)
BITMAP_LEVEL_CASES = [
    ("TestBitmap_Intersection :483", ("vals", [0, 2683177]), ("range", 628, 2683301, 1), "intersect", ("count", 1)),
    ("TestBitmap_Difference :1041", ("vals", [0, 2683177]), ("range", 628, 2683301, 1), "difference", ("count", 1)),
    ("TestBitmap_Difference2 :1053", ("vals", [0, 1, 2, 131072, 262144, SW + 5, SW + 7]), ("vals", [2, 3, 100000, 262144, 2 * SW + 1]), "difference",
     ("slice", [0, 1, 131072, SW + 5, SW + 7])),
    ("TestBitmap_Difference_Empty :1062", ("vals", [0, 2683177]), ("vals", []), "difference", ("count", 2)),
    ("TestBitmap_DifferenceArrayArray :1071", ("vals", [0, 4, 8, 12, 16, 20]), ("vals", [1, 3, 6, 9, 12, 15, 18]), "difference", ("count", 5)),
    ("TestBitmap_DifferenceArrayRun :1080", ("vals", [0, 4, 8, 12, 16, 20, 36, 40, 44]), ("vals", [1, 2, 3, 4, 5, 6, 7, 8, 9, 30, 31, 32, 33, 34, 35, 36]), "difference", ("count", 6)),
    ("TestBitmap_Union :1091", ("vals", [0, 1000001, 1000002, 1000003]), ("vals", [0, 50000, 1000001, 1000002]), "union", ("count", 5)),
    ("TestBitmap_Xor_ArrayArray :1143", ("vals", [0, 1000001, 1000002, 1000003]), ("vals", [0, 50000, 1000001, 1000002]), "xor", ("count", 2)),
    ("TestBitmap_Xor_Empty :1160", ("vals", [0, 50000, 1000001, 1000002]), ("vals", []), "xor", ("count", 4)),
    ("TestBitmap_Xor_ArrayBitmap :1169", ("vals", [1, 70, 200, 4097, 4098]), _EVEN10K, "xor", ("count", 4999)),
    ("TestBitmap_Xor_ArrayBitmap :1176 (reverse)", _EVEN10K, ("vals", [1, 70, 200, 4097, 4098]), "xor", ("count", 4999)),
    ("TestBitmap_Xor_ArrayBitmap :1188 (empty)", _EVEN10K, ("vals", []), "xor", ("count", 5000)),
    ("TestBitmap_Xor_BitmapBitmap :1199", ("range", 1, 10000, 2), _EVEN10K, "xor", ("count", 10000)),
    ("TestBitmap_IntersectionCount_ArrayArray :1283", ("vals", [0, 1000001, 1000002, 1000003]), ("vals", [0, 50000, 999998, 999999, 1000000, 1000001, 1000002]), "intersect", ("count", 3)),
    ("TestBitmap_IntersectionCount_ArrayRun :1295", ("vals", [0, 1000001, 1000002, 1000003]), ("vals", [0, 1, 2, 3, 4, 5, 1000000, 1000002, 1000003, 1000004, 1000005, 1000006]), "intersect", ("count", 3)),
    ("TestBitmap_IntersectionCount_RunRun :1308", ("vals", [3, 4, 5, 6, 7, 8, 1000001, 1000002, 1000003, 1000004]), ("vals", [0, 1, 2, 3, 4, 5, 1000000, 1000002, 1000003, 1000004, 1000005, 1000006]), "intersect", ("count", 6)),
    ("TestBitmap_IntersectionCount_BitmapRun :1322", ("range", 3, 1000007, 2), ("vals", [0, 1, 2, 3, 4, 5, 1000000, 1000002, 1000003, 1000004, 1000005, 1000006]), "intersect", ("count", 4)),
    ("TestBitmap_IntersectionCount_ArrayBitmap :1338", ("vals", [1, 70, 200, 4097, 4098]), ("range", 0, 10001, 2), "intersect", ("count", 3)),
    ("TestBitmap_IntersectionCount_BitmapBitmap :1353", ("cat", _EVEN10K + (), ("vals", [10000, 1000, 2000])), ("cat", ("range", 1, 10002, 2), ("vals", [1000, 2000])), "intersect", ("count", 2)),
    # row_test.go: shard-spanning rows (Row.Merge is a union of disjoint-or-equal segments, row.go:202)
    ("TestRow_Merge #0 row_test.go:21", ("vals", [1, 2, 3, SW + 1, 2 * SW]), ("vals", [3, 4, 5]), "union", ("count", 7)),
    ("TestRow_Merge #1 row_test.go:26", ("vals", []), ("vals", [2, 66000, 70000, 70001, 70002, 70003, 70004]), "union", ("count", 7)),
    ("TestRow_Xor row_test.go:46", ("vals", [0, 1, SW]), ("vals", [0, 2 * SW]), "xor", ("slice", [1, SW, 2 * SW])),
    ("TestRow_Xor (reverse) row_test.go:58", ("vals", [0, 2 * SW]), ("vals", [0, 1, SW]), "xor", ("slice", [1, SW, 2 * SW])),
    ("TestRow_Union_Segment row_test.go:68", ("vals", [0, 1, SW]), ("vals", [0, 2 * SW]), "union", ("slice", [0, 1, SW, 2 * SW])),
    ("TestRow_Difference_Segment row_test.go:89", ("vals", [0, 1, SW]), ("vals", [0, 2 * SW]), "difference", ("slice", [1, SW])),
    ("TestRow_IsEmpty row_test.go:103", ("vals", [0, 2 * SW]), ("vals", [1, SW]), "intersect", ("count", 0)),
]

# ---------------------------------------------------------------------------------------------------
# fragment_internal_test.go TestFragmentPositionsForValue (BSI bit layout: exists row 0, sign row 1, bit i row 2+i):
# (column, bitDepth, value, positions set).  TestIntLTRegression: value 33 at depth 6, Row(v < 33) is empty.
# ---------------------------------------------------------------------------------------------------
BSI_POSITIONS = [(0, 1, 0, [0]), (0, 3, 0, [0]), (1, 3, 0, [1]), (0, 1, 1, [0, 2 * SW]), (0, 4, 10, [0, 3 * SW, 5 * SW]), (0, 5, 10, [0, 3 * SW, 5 * SW])]
BSI_LT_REGRESSION = (1, 6, 33)


# ---------------------------------------------------------------------------------------------------
# Mixed-encoding container cases of roaring_internal_test.go that are not table tests (so not in kernel_tables*.json).
# Operand specs: ("array", [values]) / ("run", [(start, last), ...]) / ("bitmap", [first words]).
# (cite, op, a, b, expected values, encoding the reference reads the result through: "array" / "run" / "bitmap" / None)
# ---------------------------------------------------------------------------------------------------
_UM_A, _UM_B, _UM_R = ("array", [1, 4, 5, 7, 10, 11, 12]), ("bitmap", [0x3]), ("run", [(5, 10)])
_IM_A, _IM_B, _IM_C = ("run", [(5, 10)]), ("array", [1, 4, 5, 7, 10, 11, 12]), ("bitmap", [0x60])
_DM_A, _DM_B, _DM_C, _DM_D = ("run", [(5, 10)]), ("array", [0, 2, 4, 6, 8, 10, 12]), ("bitmap", [0x64]), ("array", [1, 3, 5, 7, 9, 11, 12])
MIXED_CONTAINER_CASES = [
    # TestUnionMixed :694-735 (results compared as arrays after conversion: encoding not asserted)
    (":712 run-array", "union", _UM_R, _UM_A, [1, 4, 5, 6, 7, 8, 9, 10, 11, 12], None),
    (":713 array-run", "union", _UM_A, _UM_R, [1, 4, 5, 6, 7, 8, 9, 10, 11, 12], None),
    (":714 run-run", "union", _UM_R, _UM_R, [5, 6, 7, 8, 9, 10], None),
    (":716 bitmap-run", "union", _UM_B, _UM_R, [0, 1, 5, 6, 7, 8, 9, 10], None),
    (":717 run-bitmap", "union", _UM_R, _UM_B, [0, 1, 5, 6, 7, 8, 9, 10], None),
    (":718 array-bitmap", "union", _UM_A, _UM_B, [0, 1, 4, 5, 7, 10, 11, 12], None),
    # TestIntersectMixed :918-955
    (":923", "intersect", _IM_A, _IM_B, [5, 7, 10], "array"),
    (":927", "intersect", _IM_B, _IM_A, [5, 7, 10], "array"),
    (":931", "intersect", _IM_A, _IM_A, [5, 6, 7, 8, 9, 10], "run"),
    (":936", "intersect", _IM_C, _IM_A, [5, 6], "array"),
    (":941", "intersect", _IM_A, _IM_C, [5, 6], "array"),
    (":946", "intersect", _IM_B, _IM_C, [5], "array"),
    (":950", "intersect", _IM_C, _IM_B, [5], "array"),
    # TestDifferenceMixed :956-1022
    (":965", "difference", _DM_A, _DM_B, [5, 7, 9], "array"),
    (":971", "difference", _DM_B, _DM_A, [0, 2, 4, 12], "array"),
    (":976", "difference", _DM_A, _DM_A, [], None),
    (":981", "difference", _DM_C, _DM_A, [2], "bitmap"),
    (":986", "difference", _DM_A, _DM_C, [7, 8, 9, 10], "run"),
    (":991", "difference", _DM_B, _DM_C, [0, 4, 8, 10, 12], "array"),
    (":996", "difference", _DM_C, _DM_B, [5], "array"),
    (":1001", "difference", _DM_B, _DM_B, [], None),
    (":1006", "difference", _DM_C, _DM_C, [], None),
    (":1011", "difference", _DM_D, _DM_B, [1, 3, 5, 7, 9, 11], "array"),
    (":1016", "difference", _DM_B, _DM_D, [0, 2, 4, 6, 8, 10], "array"),
    # TestXorRunRun1 :2026-2037
    (":2029", "xor", ("run", [(4, 10)]), ("run", [(5, 10)]), [4], "array"),
    (":2033", "xor", ("run", [(5, 10)]), ("run", [(4, 10)]), [4], "array"),
]
# TestIntersectionCountArrayBitmap3 :284-304 (full containers through bitmap / run encodings: |a ∩ b| = 65536 every way) and
# TestDifferenceInPlace_N :4316-4323 (full run \ full bitmap is empty)
FULL_CONTAINER_ENCODINGS = [("bitmap", "bitmap"), ("bitmap", "run"), ("run", "bitmap"), ("run", "run")]
# TestRunCountRange :144-236: (runs, start, end, expected count of [start, end)); the last run list must count 3 runs
RUN_COUNT_RANGE = [
    ([], 2, 9, 0), ([(5, 7)], 2, 9, 3),
    ([(5, 11)], 4, 8, 3), ([(5, 11)], 5, 8, 3), ([(5, 11)], 6, 8, 2), ([(5, 11)], 3, 9, 4), ([(5, 11)], 9, 14, 3), ([(5, 11)], 8, 10, 2),
    ([(5, 11)], 8, 11, 3), ([(5, 11)], 8, 12, 4), ([(5, 11)], 5, 12, 7), ([(5, 11)], 5, 11, 6),
    ([(5, 11), (17, 19)], 1, 22, 10), ([(5, 11), (13, 14), (17, 19)], 6, 18, 9),
]
