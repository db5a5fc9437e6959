#!/bin/bash
# round 2, call 12: GroupBy slots-per-group target (table load factor vs units), pair kernel prefetch distance and 7-warp CTAs
set -u
out=gpurun_out/r2_call12; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "config1 or padded or density_sweep or groupby" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
for t in "" 3200 1600 12288; do echo "gb[target=${t:-default}] $(FBGPU_GH_TARGET=$t python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl; done
for t in "" 2200 1100; do echo "gb512[target=${t:-default}] $(FBGPU_GH_TARGET=$t FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_gh512.so python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl; done
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
for v in pair_pf3 pair_w7; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so b)" >> $out/ab.jsonl; done
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log
