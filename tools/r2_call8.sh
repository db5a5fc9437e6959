#!/bin/bash
# round 2, call 8: two-warp teams in pair_count_kernel (CTA shapes), groupby_shard_kernel with warp-balanced entry lists, 32-byte
# slices in the word-parallel kernel; then the whole bench
set -u
out=gpurun_out/r2_call8; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "config1 or density_sweep or mixed_encoding or groupby or bsi_range or bsi_uniform or FORCE_WORDPAR or any_early or thread_safety or container_combinations or bsi_diagonal" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
for v in pair_mb3 pair_t4; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so b)" >> $out/ab.jsonl; done
for v in "" wp_narrow wp_async4; do echo "wp[${v:-default}] $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu${v:+_$v}.so python bench_sweep.py --configs 3 2>>$out/bench_err.log | head -1)" >> $out/wp.jsonl; done
for v in "" "FBGPU_GROUPBY_CTA=1 FBGPU_GROUPBY_FAST=1"; do echo "gb[$v] $(env $v python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl; done
ncu --set full --clock-control none --import-source on -k regex:pair_count -c 2 --launch-skip 70 -f -o $out/pair python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pair.log 2>&1
ncu -i $out/pair.ncu-rep --page raw --csv > $out/pair_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:groupby_shard -c 1 --launch-skip 2 -f -o $out/gbh python bench_sweep.py --configs 4 > $out/ncu_gbh.log 2>&1
ncu -i $out/gbh.ncu-rep --page raw --csv > $out/gbh_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:eval_wordpar -c 1 --launch-skip 6 -f -o $out/wp python bench_sweep.py --configs 3 > $out/ncu_wp.log 2>&1
ncu -i $out/wp.ncu-rep --page raw --csv > $out/wp_raw.csv 2>/dev/null
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2>> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
timeout 900 python bench_sweep.py --configs 5 --batched > $out/sweep5.jsonl 2>> $out/bench_err.log
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log
