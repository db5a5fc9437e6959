#!/bin/bash
# round 2, call 11: pair kernel L2 prefetch of the next unit (A/B), bitmap loop unroll, 8-warp CTAs; groupby_shard_kernel back to per-lane walks, 512-thread CTAs
set -u
out=gpurun_out/r2_call11a; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "config1 or padded or density_sweep or mixed_encoding or groupby or container_combinations" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
for v in pair_nopf pair_w8 pair_bm4; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so b)" >> $out/ab.jsonl; done
for v in "" gh512; do echo "gb[${v:-default}] $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu${v:+_$v}.so python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl; done
for v in "" pair_bm4 pair_nopf; do FBGPU_LIB=$PWD/featurebase_b200/libfbgpu${v:+_$v}.so python bench_sweep.py --configs 5 --batched --generators uniform --densities 0.0001,0.01,0.125 2>>$out/bench_err.log | sed "s/^/${v:-default} /" >> $out/sweep5_ab.jsonl; done
ncu --set full --clock-control none --import-source on -k regex:groupby_shard -c 1 --launch-skip 2 -f -o $out/gbh python bench_sweep.py --configs 4 > $out/ncu_gbh.log 2>&1
ncu -i $out/gbh.ncu-rep --page raw --csv > $out/gbh_raw.csv 2>/dev/null
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log
