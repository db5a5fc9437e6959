#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native roaring executor (contract: see DESIGN.md §Measurement).

Workload (BASELINE.json configs[1]): per GPU 1024 shards x 2^20 columns, 64 rows at 1 % density, one step =
    Count(Intersect(Union(Row(f=0..31)), Union(Row(f=32..63))))   over the GPU's whole shard batch
= 63 Row-level set-ops + 1 Count row per shard.  metric = set-ops/s (whole job, all GPUs); extras: Count rows/s,
columns/s, HBM GB/s vs the measured roofline.

  python bench.py --gpus 1 --steps 20 --warmup 3            # our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1     # reference arm: CPU restatement on all host cores
  torchrun ... bench.py --gpus N ...                        # one rank per GPU, shards range-partitioned, NCCL count reduce
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIELD_SEED_ID = 1
ROWS_A, ROWS_B = list(range(32)), list(range(32, 64))
SET_OPS_PER_SHARD = (len(ROWS_A) - 1) + (len(ROWS_B) - 1) + 1     # 31 + 31 unions, 1 intersect
SW = 1 << 20


def query_text():
    ua = "Union(" + ", ".join(f"Row(f={r})" for r in ROWS_A) + ")"
    ub = "Union(" + ", ".join(f"Row(f={r})" for r in ROWS_B) + ")"
    return f"Count(Intersect({ua}, {ub}))"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc, self.t = gpu, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(n_sample_shards, threads, reps):
    """the reference's algorithm (CPU restatement: oracle port, since Go is absent) on a bounded sample of the workload"""
    from featurebase_b200 import datagen as D
    from oracle import oracle as O
    shards = np.arange(n_sample_shards, dtype=np.uint64)
    bulk = D.fragments(FIELD_SEED_ID, shards, ROWS_A + ROWS_B, 0.01)
    frags = [O.Bitmap.from_bytes(bulk.fragment_bytes(i)) for i in range(n_sample_shards)]
    best, count = None, None
    times = []
    for _ in range(reps):
        count, secs = O.bench_union_intersect_count(frags, shards, ROWS_A, ROWS_B, threads)
        times.append(secs)
    return count, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = 1024                     # the whole configs[1] workload (about 2 core-seconds per step)
    _, _ = cpu_reference_run(min(sample, 64), threads, 1) if args.warmup else (None, None)
    count, times = cpu_reference_run(sample, threads, max(args.steps, 1))
    sec = float(np.median(times))
    val = SET_OPS_PER_SHARD * sample / sec
    line = {
        "impl": "reference", "metric": "set_ops_per_sec", "value": val, "unit": "set-ops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64/u16 integer", "data": "synthetic",
        "config": {"workload": "configs[1]: 1024 shards x 2^20 cols, 1% density, 64-row Union->Intersect->Count", "query": "Count(Intersect(Union(32 rows),Union(32 rows)))",
                   "sample_shards": sample, "density": 0.01},
        "count_rows_per_sec": sample / sec, "columns_per_sec": sample * SW / sec,
        "cpu_baseline": {"value": val, "unit": "set-ops/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of 1024 shards, {len(times)} reps, median; C restatement of reference algorithms (Go toolchain absent), one thread per shard range"},
        "e2e": {"value": val, "unit": "set-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "check_count": count,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--shards-per-gpu", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-memory Count merge (default) or ncclAllReduce")
    ap.add_argument("--cold", action="store_true", help="also time fragment upload + query (e2e_cold_load)")
    args = ap.parse_args()
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__
        __graft_entry__.build()          # no-op when libfbgpu.so / datagen / oracle are up to date
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from featurebase_b200 import datagen as D
    from featurebase_b200 import executor as X
    from featurebase_b200 import pql

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- synthetic shard batch of this rank: contiguous shard range [rank*S, (rank+1)*S)  
    S = args.shards_per_gpu
    shards = np.arange(rank * S, (rank + 1) * S, dtype=np.uint64)
    t0 = time.time()
    bulk = D.fragments(FIELD_SEED_ID, shards, ROWS_A + ROWS_B, 0.01)
    t_gen = time.time() - t0
    h = X.Holder(device=local)
    idx = h.create_index("i", track_existence=False)
    fld = idx.create_field("f")
    ex = X.Executor(h)
    t0 = time.time()
    h.ctx.load_fragments(idx.id, fld.id, X.VIEW_STANDARD, shards, bulk.buf, bulk.offsets)
    h.ctx.commit()
    t_load = time.time() - t0
    idx.shards.update(int(s) for s in shards)
    if world > 1:   # library-owned NCCL communicator for the count all-reduce
        uid = [h.ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.ctx.comm_init(world, rank, uid[0])
        if args.reduce == "p2p":   # fused Count merge over NVLink peer memory (mailboxes mapped through CUDA IPC)
            ok = 1
            try:
                handles = [None] * world
                dist.all_gather_object(handles, h.ctx.comm_p2p_handle())
                h.ctx.comm_p2p_open(world, rank, handles)
            except Exception as e:  # noqa: BLE001 — e.g. a peer that cannot be IPC-mapped: every rank falls back together
                print(f"[rank {rank}] peer-memory merge unavailable ({e}); using ncclAllReduce", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if ok:
                    h.ctx.comm_p2p_disable()
                args.reduce = "nccl"
    ops = ex._bitmap_call(idx, pql.parse(query_text())[0].children[0])
    payload, n_cont = h.ctx.rows_payload_bytes(idx.id, fld.id, X.VIEW_STANDARD, shards, ROWS_A + ROWS_B)
    algo_bytes = payload + 16 * n_cont + 8                       # payload + 16 B/descriptor + 8 B count
    h2d_bytes = 48 * len(ops) + 8 * len(shards)                  # fbgpu_op program + shard list (host buffers)
    ops = X.L.ops_array(ops)                                     # marshal the program once (host memory; still copied H2D by every call)
    shards = np.ascontiguousarray(np.asarray(shards, dtype=np.uint64))

    def step():
        return h.ctx.count(idx.id, ops, shards)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        expect = step()
    c0 = h.ctx.counters()["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        got = step()
        kernel_ms.append(h.ctx.counters()["last_query_gpu_ms"])   # CUDA events on the library's launching stream
        assert got == expect
    sync_all()
    wall = time.perf_counter() - t_begin
    launches = h.ctx.counters()["kernel_launches"] - c0
    kms = float(np.mean(kernel_ms))
    # max over ranks (device time of the kernels; wall time of the C-ABI calls)
    if world > 1:
        t = torch.tensor([kms, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms, wall = float(t[0]), float(t[1])
    # nvidia-smi cannot sample faster than ~10 Hz; when the timed region was shorter than that, keep the identical
    # step running (untimed; the SAME number of steps on every rank, the step contains a collective) until the sampler
    # has seen the GPU under this load, and say so
    probe_note = None
    if wall < 0.5:
        for _ in range(int(0.7 / max(wall / args.steps, 1e-5)) + 1):
            step()
        probe_note = "timed region %.0f ms < sampler period: clocks sampled over the timed region plus about 0.7 s of the identical step" % (wall * 1e3)
    clocks = sampler.stop()
    if probe_note:
        clocks["note"] = probe_note
    total_shards = S * world
    set_ops = SET_OPS_PER_SHARD * total_shards
    value = set_ops / (kms * 1e-3)
    e2e_ms = wall / args.steps * 1e3
    peak, peak_src = measured_peaks()
    achieved = algo_bytes / (kms * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of eval_kernel from the committed `ncu --set full` capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["eval_kernel"]["dram_bytes_per_launch"]
    except Exception:
        pass

    cold = None
    if args.cold and rank == 0:
        h2 = X.Holder(device=local)
        i2 = h2.create_index("i", track_existence=False)
        f2 = i2.create_field("f")
        t0 = time.perf_counter()
        h2.ctx.load_fragments(i2.id, f2.id, X.VIEW_STANDARD, shards, bulk.buf, bulk.offsets)
        n = h2.ctx.count(i2.id, ops, shards)
        cold = {"ms": (time.perf_counter() - t0) * 1e3, "h2d_bytes": int(bulk.offsets[-1]), "note": "parse + stage + H2D of every fragment, then the query"}
        assert n == expect or world > 1
        h2.ctx.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = S                     # the full per-GPU workload, ~2 core-seconds per rep
        cnt, times = cpu_reference_run(sample, threads, 5)
        sec = float(np.median(times))
        cpu = {"value": SET_OPS_PER_SHARD * sample / sec, "unit": "set-ops/s", "cores": threads, "kind": "port",
               "sample": f"{sample} of {S} shards x 5 reps (median {sec * 1e3:.1f} ms); C restatement of the reference algorithms, static shard partition over all host threads"}

    if rank == 0:
        line = {
            "metric": "set_ops_per_sec", "value": value, "unit": "set-ops/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": kms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/u16 integer", "data": "synthetic",
            "config": {"workload": "configs[1]: 1024 shards x 2^20 cols per GPU, 1% density, 64-row Union->Intersect->Count",
                       "query": "Count(Intersect(Union(Row f=0..31),Union(Row f=32..63)))", "shards_per_gpu": S, "total_shards": total_shards,
                       "density": 0.01, "l2": (f"inputs {payload / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)" if payload > 126e6 else
                                                   f"inputs {payload / 1e6:.0f} MB per GPU fit the 126 MB L2: NOT a valid bench size (use the default --shards-per-gpu)"), "parallelism": f"shard-range x{world}", "count_merge": (args.reduce if world > 1 else "none")},
            "count_rows_per_sec": total_shards / (kms * 1e-3), "columns_per_sec": total_shards * SW / (kms * 1e-3),
            "check_count": int(expect),
            "e2e": {"value": set_ops / (e2e_ms * 1e-3), "unit": "set-ops/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                    "note": "fbgpu_count() through the C ABI from host buffers (program + shard list H2D, count D2H) with fragments resident in HBM, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "eval_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src,
                         "timing": "CUDA events on the library's launching stream around the kernel, mean over the timed steps"},
            "setup": {"datagen_s": t_gen, "load_commit_s": t_load, "payload_bytes_per_gpu": int(payload), "containers": int(n_cont)},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if cold:
            line["e2e_cold_load"] = cold
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
