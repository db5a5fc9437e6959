#!/usr/bin/env python3
"""bench_sweep.py — secondary measurements on one GPU (BASELINE configs 3, 4, 5); prints one JSON line per point.

  config 5: density sweep p in {0.01 %..50 %} x {uniform, clustered}, Count(Intersect(Row a, Row b)) over 1024 shards
            (fused pair_count_kernel); row pairs are rotated between steps so that the touched data exceeds L2.
  config 3: BSI Count(Row(v > k)) over 10 M records, 32-bit values (eval_kernel plane sweep).
  config 4: GroupBy(Rows(a), Rows(b)) 256 x 256 over this GPU's share (512 shards) of 100 M records / 4096 shards.
  config X: fbgpu_columns / fbgpu_extract (device-side column-id and int-value expansion), wall clock; R: fbgpu_row.
Every point is spot-checked against the CPU oracle on a few shards (the checker, not the thing measured)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SW = 1 << 20


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(ctx, fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    ms = []
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
        ms.append(ctx.counters()["last_query_gpu_ms"])
    wall = (time.perf_counter() - t0) / steps * 1e3
    return float(np.mean(ms)), float(np.min(ms)), wall


def config5(args, out):
    from featurebase_b200 import datagen as D, executor as X, lib as L
    from oracle import oracle as O
    pk, src = peak()
    S = args.shards
    shards = np.arange(S, dtype=np.uint64)
    for mode, mname in ((0, "uniform"), (1, "clustered")):
        if mname not in args.generators.split(","):
            continue
        for p in [float(x) for x in args.densities.split(",")]:
            per_pair = 2 * S * 16 * max(2 * p * 65536, 16)
            n_pairs = int(min(8, max(2, np.ceil(300e6 / max(per_pair, 1)))))
            if p >= 0.0625 and mode == 0:
                n_pairs = 2
            rows = list(range(2 * n_pairs))
            h = X.Holder()
            idx = h.create_index("i", track_existence=False)
            f = idx.create_field("f")
            bulk = D.fragments(11 + mode, shards, rows, p, mode=mode, mean_run=64.0)
            h.ctx.load_fragments(idx.id, f.id, X.VIEW_STANDARD, shards, bulk.buf, bulk.offsets)
            h.ctx.commit()
            st = h.ctx.stats()
            progs = [[L.Op(L.OP_ROW, f.id, 0, 0, 2 * k, 0, 0, 0), L.Op(L.OP_ROW, f.id, 0, 0, 2 * k + 1, 0, 0, 0), L.Op(L.OP_INTERSECT, 0, 0, 2, 0, 0, 0, 0)] for k in range(n_pairs)]
            counts = {}

            def step(i):
                k = i % n_pairs
                counts[k] = h.ctx.count(idx.id, progs[k], shards)

            ms, ms_min, wall = timed(h.ctx, step, args.steps)
            pay, nc = h.ctx.rows_payload_bytes(idx.id, f.id, X.VIEW_STANDARD, shards, [0, 1])
            algo = pay + 16 * nc + 8
            # oracle spot check: pair 0 on 2 shards
            tot, per = h.ctx.count(idx.id, progs[0], shards, per_shard=True)
            for s in (0, S - 1):
                fr = O.Bitmap.from_bytes(bulk.fragment_bytes(s))
                assert int(per[s]) == fr.row(0, s).intersection_count(fr.row(1, s)), (mname, p, s)
            assert tot == counts[0]
            gbs = algo / (ms * 1e-3) / 1e9
            out({"config": 5, "generator": mname, "density": p, "shards": S, "kernel": "pair_count_kernel", "ms": ms, "ms_min": ms_min, "wall_ms": wall,
                 "set_ops_per_sec": S / (ms * 1e-3), "count_rows_per_sec": S / (ms * 1e-3), "columns_per_sec": S * SW / (ms * 1e-3),
                 "algorithmic_bytes": int(algo), "achieved_gbs": gbs, "peak_gbs": pk, "peak_source": src, "frac": gbs / pk,
                 "l2_note": f"{n_pairs} row pairs rotated, {n_pairs * algo / 1e6:.0f} MB touched per cycle" + (" (< L2: launch/L2-bound point)" if n_pairs * algo < 126e6 else ""),
                 "containers": {"array": st["array_containers"], "bitmap": st["bitmap_containers"], "run": st["run_containers"]}, "count": int(tot)})
            if args.batched:
                # the same Intersect+Count, N independent row pairs fused in one launch (SURVEY §8d "(ii) batched")
                nb = max(n_pairs, 4)
                ra, rb = [2 * (k % n_pairs) for k in range(nb)], [2 * (k % n_pairs) + 1 for k in range(nb)]
                res = {}

                def bstep(i):
                    res[0] = h.ctx.count_pairs(idx.id, f.id, 0, ra, f.id, 0, rb, shards)

                bms, bmin, bwall = timed(h.ctx, bstep, args.steps)
                for k in range(n_pairs):                         # (fewer timed steps than row pairs: count the ones the rotation did not reach)
                    if k not in counts:
                        counts[k] = h.ctx.count(idx.id, progs[k], shards)
                assert int(res[0][0]) == counts[0] and int(res[0].sum()) == sum(counts[k % n_pairs] for k in range(nb))
                bgbs = nb * algo / (bms * 1e-3) / 1e9
                out({"config": "5b", "generator": mname, "density": p, "shards": S, "pairs_per_launch": nb, "kernel": "pair_count_kernel (multi-pair)", "ms": bms, "ms_min": bmin, "wall_ms": bwall,
                     "set_ops_per_sec": nb * S / (bms * 1e-3), "algorithmic_bytes": int(nb * algo), "achieved_gbs": bgbs, "peak_gbs": pk, "peak_source": src, "frac": bgbs / pk,
                     "l2_note": f"{nb} pairs over {n_pairs} distinct row pairs ({n_pairs * algo / 1e6:.0f} MB distinct data)"})
            h.ctx.close()


def config_row(args, out):
    """Row-returning calls (canonical Pilosa-roaring bytes to the host): fbgpu_row end to end, wall clock"""
    from featurebase_b200 import datagen as D, executor as X, pql
    from oracle import oracle as O
    S = args.shards
    shards = np.arange(S, dtype=np.uint64)
    h = X.Holder()
    idx = h.create_index("i", track_existence=False)
    f = idx.create_field("f")
    ex = X.Executor(h)
    bulk = D.fragments(11, shards, [0, 1, 2, 3], 0.01)
    h.ctx.load_fragments(idx.id, f.id, X.VIEW_STANDARD, shards, bulk.buf, bulk.offsets)
    h.ctx.commit()
    idx.shards.update(range(S))
    for q in ("Intersect(Row(f=0), Row(f=1))", "Union(Row(f=0), Row(f=1), Row(f=2), Row(f=3))", "Row(f=0)"):
        ops = ex._bitmap_call(idx, pql.parse(q)[0])
        data, cnt = h.ctx.row(idx.id, ops, shards)               # (also sizes the buffer below)
        buf = np.zeros(len(data) + 4096, dtype=np.uint8)         # caller-owned and reused, as a Go caller would: pages already mapped
        for _ in range(2):
            h.ctx.row_into(idx.id, ops, shards, buf)
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            need, cnt, fits = h.ctx.row_into(idx.id, ops, shards, buf)
        wall = (time.perf_counter() - t0) / n * 1e3
        assert fits and buf[:need].tobytes() == data
        # oracle spot check: the first shard's segment
        fr = O.Bitmap.from_bytes(bulk.fragment_bytes(0))
        call = pql.parse(q)[0]
        from tests.oracle_exec import OracleIndex
        oi = OracleIndex(idx)
        oi.load("f", 0, 0, bulk.fragment_bytes(0))
        sub, _ = h.ctx.row(idx.id, ops, [0])
        assert sub == oi.eval_row(call, [0]).to_bytes()
        out({"config": "R", "query": q, "shards": S, "kernel": "eval_kernel + canon_emit_kernel + host assembly", "ms": wall, "result_bytes": len(data), "result_count": int(cnt),
             "columns_per_sec": S * SW / (wall * 1e-3), "frac": 0.0, "achieved_gbs": 0.0, "note": "wall clock of fbgpu_row() into a reused caller buffer: evaluate, {N,runs} D2H, encoding choice on host, emit, payload D2H, roaring assembly"})
    h.ctx.close()


def config_extract(args, out):
    """Column-id and value expansion on the device: fbgpu_columns over a 1 % row, fbgpu_extract over a 32-bit int field
    (10 M records), wall clock through the C ABI; every result is checked against the data generator"""
    from featurebase_b200 import datagen as D, executor as X, pql, roaring_io
    S = args.shards
    shards = np.arange(S, dtype=np.uint64)
    h = X.Holder()
    idx = h.create_index("i", track_existence=False)
    f = idx.create_field("f")
    v = idx.create_field("v", "int", min=0, max=(1 << 32) - 1)
    ex = X.Executor(h)
    bulk = D.fragments(11, shards, [0, 1], 0.01)
    h.ctx.load_fragments(idx.id, f.id, X.VIEW_STANDARD, shards, bulk.buf, bulk.offsets)
    n_rec = min(10_000_000, S * SW)
    n_sh = (n_rec + SW - 1) // SW
    for s in range(n_sh):
        h.ctx.load_fragment(idx.id, v.id, X.VIEW_BSI, s, D.bsi_fragment(12, s, min(SW, n_rec - s * SW), 32, 0, (1 << 32) - 1))
    h.ctx.commit()
    idx.shards.update(range(S))

    def timed(fn, n=5):
        for _ in range(2):
            r = fn()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        return r, (time.perf_counter() - t0) / n * 1e3

    ops = ex._bitmap_call(idx, pql.parse("Union(Row(f=0), Row(f=1))")[0])
    (cols, total), wall = timed(lambda: h.ctx.columns(idx.id, ops, shards))
    want = roaring_io.decode(h.ctx.row(idx.id, ops, shards)[0])
    assert total == len(want) and np.array_equal(cols, np.asarray(want, dtype=np.uint64))
    out({"config": "X", "query": "columns of Union(Row(f=0), Row(f=1)) at 1 %", "shards": S, "kernel": "eval_kernel + columns_emit_kernel", "ms": wall, "columns": int(total),
         "columns_per_sec": float(total) / (wall * 1e-3), "frac": 0.0, "achieved_gbs": 0.0, "note": "wall clock of fbgpu_columns(): evaluate, N per unit D2H, expand ids on the device, ids D2H"})
    bsh = np.arange(n_sh, dtype=np.uint64)
    (c2, vals, tot2), wall = timed(lambda: h.ctx.extract(idx.id, v.id, X.VIEW_BSI, 32, bsh), n=3)
    assert tot2 == n_rec and len(c2) == n_rec
    for i in (0, 1, n_rec // 2, n_rec - 1):
        assert int(vals[i]) == D.bsi_value(12, int(c2[i]) // SW, int(c2[i]) % SW, 0, (1 << 32) - 1)
    out({"config": "X", "query": "values of a 32-bit int field, all records", "records": n_rec, "kernel": "eval_kernel + columns_emit_kernel + extract_values_kernel", "ms": wall,
         "records_per_sec": n_rec / (wall * 1e-3), "frac": 0.0, "achieved_gbs": 0.0, "note": "wall clock of fbgpu_extract(): 34 planes read once, 16 B per record D2H"})
    # the one-pass aggregates, cross-checked against the extracted value vector
    (tot, cnt), wall = timed(lambda: h.ctx.bsi_sum(idx.id, v.id, X.VIEW_BSI, 32, bsh))
    assert cnt == n_rec and tot == int(vals.astype(object).sum())
    out({"config": "X", "query": "Sum(field=v), 32-bit, all records", "records": n_rec, "kernel": "eval_kernel + bsi_sum_kernel", "ms": wall, "records_per_sec": n_rec / (wall * 1e-3),
         "frac": 0.0, "achieved_gbs": 0.0, "note": "wall clock of fbgpu_bsi_sum()"})
    for want_max in (False, True):
        (val, n), wall = timed(lambda: h.ctx.bsi_minmax(idx.id, v.id, X.VIEW_BSI, 32, bsh, want_max))
        ref = int(vals.max() if want_max else vals.min())
        assert (val, n) == (ref, int((vals == ref).sum()))
        out({"config": "X", "query": ("Max" if want_max else "Min") + "(field=v), 32-bit, all records", "records": n_rec, "kernel": "eval_kernel + bsi_minmax_kernel", "ms": wall,
             "records_per_sec": n_rec / (wall * 1e-3), "frac": 0.0, "achieved_gbs": 0.0, "note": "wall clock of fbgpu_bsi_minmax()"})
    h.ctx.close()


def config3(args, out, n_rec=10_000_000, nf=4):
    """nf fields are rotated between steps so that the touched planes exceed L2 (the 10 M-record config is 42.5 MB)"""
    from featurebase_b200 import datagen as D, executor as X, pql
    from oracle import oracle as O
    from tests.oracle_exec import OracleIndex
    pk, src = peak()
    n_sh = (n_rec + SW - 1) // SW
    shards = np.arange(n_sh, dtype=np.uint64)
    h = X.Holder()
    idx = h.create_index("i", track_existence=False)
    ex = X.Executor(h)
    ora = OracleIndex(idx)
    for k in range(nf):
        fld = idx.create_field(f"v{k}", "int", min=0, max=(1 << 32) - 1)
        for s in range(n_sh):
            ncols = min(SW, n_rec - s * SW)
            data = D.bsi_fragment(20 + k, s, ncols, 32, 0, (1 << 32) - 1)
            h.import_roaring("i", f"v{k}", X.VIEW_BSI, s, data)
            if k == 0 and s in (0, n_sh - 1):
                ora.load("v0", X.VIEW_BSI, s, data)
    h.ctx.commit()
    for kname, kval in (("2^31", 1 << 31), ("0.99*2^32", int(0.99 * (1 << 32)))):
        progs = [ex._bitmap_call(idx, pql.parse(f"Row(v{k} > {kval})")[0]) for k in range(nf)]
        res = {}

        def step(i):
            res[i % nf] = h.ctx.count(idx.id, progs[i % nf], shards)

        ms, ms_min, wall = timed(h.ctx, step, args.steps)
        # planes the compiled program touches: exists, sign, and every bit row down to where the predicate saturates
        pay, nc = h.ctx.rows_payload_bytes(idx.id, idx.fields["v0"].id, X.VIEW_BSI, shards, None)
        algo = pay + 16 * nc + 8
        call = pql.parse(f"Row(v0 > {kval})")[0]
        exp = sum(ora.eval_shard(call, s).count() for s in (0, n_sh - 1))
        tot, per = h.ctx.count(idx.id, progs[0], shards, per_shard=True)
        assert int(per[0]) + int(per[n_sh - 1]) == exp
        gbs = algo / (ms * 1e-3) / 1e9
        out({"config": 3, "query": f"Count(Row(v > {kname}))", "records": n_rec, "shards": n_sh, "kernel": "eval_kernel (BSI plane sweep)", "ms": ms, "ms_min": ms_min, "wall_ms": wall,
             "records_per_sec": n_rec / (ms * 1e-3), "algorithmic_bytes": int(algo), "achieved_gbs": gbs, "peak_gbs": pk, "peak_source": src, "frac": gbs / pk,
             "note": f"algorithmic bytes = all 34 planes of the field (upper bound; the sweep stops early when the predicate saturates); {nf} field(s) rotated",
             "count": int(tot), "selectivity": tot / n_rec})
    h.ctx.close()


def config4(args, out):
    from featurebase_b200 import datagen as D, executor as X
    from oracle import oracle as O
    pk, src = peak()
    S = args.groupby_shards
    p_rec = 100e6 / (4096 * SW)
    shards = np.arange(S, dtype=np.uint64)
    h = X.Holder()
    idx = h.create_index("i", track_existence=False)
    fa, fb = idx.create_field("a"), idx.create_field("b")
    keep = {}
    t0 = time.time()
    for s in range(S):
        da, db = D.groupby_fragments(31, 32, s, p_rec, 256, 256)
        h.import_roaring("i", "a", X.VIEW_STANDARD, s, da)
        h.import_roaring("i", "b", X.VIEW_STANDARD, s, db)
        if s in (0, S - 1):
            keep[s] = (da, db)
    h.ctx.commit()
    rows = list(range(256))
    res = {}

    def step(i):
        res[0] = h.ctx.groupby(idx.id, [fa.id, fb.id], [0, 0], [rows, rows], shards)

    ms, ms_min, wall = timed(h.ctx, step, args.steps)
    pa, na = h.ctx.rows_payload_bytes(idx.id, fa.id, 0, shards, None)
    pb, nb = h.ctx.rows_payload_bytes(idx.id, fb.id, 0, shards, None)
    algo = pa + pb + 16 * (na + nb) + 8 * 65536
    # oracle spot check on two shards
    sub = h.ctx.groupby(idx.id, [fa.id, fb.id], [0, 0], [rows, rows], np.array(sorted(keep), dtype=np.uint64))
    exp = np.zeros(65536, dtype=np.uint64)
    for s, (da, db) in keep.items():
        O.groupby_shard([O.Bitmap.from_bytes(da), O.Bitmap.from_bytes(db)], s, [rows, rows], None, exp)
    assert np.array_equal(sub.reshape(-1), exp)
    total = int(res[0].sum())
    gbs = algo / (ms * 1e-3) / 1e9
    out({"config": 4, "query": "GroupBy(Rows(a), Rows(b)) 256x256", "shards": S, "records": total, "kernel": ("groupby_shard_kernel" if os.environ.get("FBGPU_GROUPBY_HASH") else "groupby_direct_kernel"), "ms": ms, "ms_min": ms_min, "wall_ms": wall,
         "records_per_sec": total / (ms * 1e-3), "group_counts_per_sec": 65536 * S / (ms * 1e-3), "algorithmic_bytes": int(algo), "payload_bytes": int(pa + pb),
         "achieved_gbs": gbs, "peak_gbs": pk, "peak_source": src, "frac": gbs / pk, "nonzero_groups": int((res[0] > 0).sum()),
         "note": f"this GPU's 1/8 share ({S} of 4096 shards) of the 100 M-record config; load {time.time() - t0:.1f}s"})
    h.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="5,3,4")
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--shards", type=int, default=1024)
    ap.add_argument("--groupby-shards", type=int, default=512)
    ap.add_argument("--generators", default="uniform,clustered")
    ap.add_argument("--batched", action="store_true", help="also time the multi-pair launch (config 5b)")
    ap.add_argument("--densities", default="0.0001,0.001,0.01,0.03,0.0625,0.125,0.25,0.5")
    args = ap.parse_args()

    def out(d):
        print(json.dumps(d), flush=True)

    for c in args.configs.split(","):
        c = c.strip()
        if c == "R":
            config_row(args, out)
        elif c == "X":
            config_extract(args, out)
        elif c == "3L":     # the same BSI query at 256 shards (268 M records, 1.1 GB of planes): shows the kernel away from the launch-bound regime
            config3(args, out, n_rec=256 * SW, nf=1)
        else:
            {"5": config5, "3": config3, "4": config4}[c](args, out)


if __name__ == "__main__":
    main()
