#!/bin/bash
# round 2, call 10: lean one-warp-per-pair kernel (dispatch hoisted into the resolve round, ~400 instructions per pair), CTA shapes;
# groupby_shard_kernel with fixed-slot staging; eval_kernel without the directory prefetch again
set -u
out=gpurun_out/r2_call10; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "config1 or padded or density_sweep or mixed_encoding or any_early or thread_safety or container_combinations or full_size_properties_1024 or topk or groupby or archetype or kernel_table" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
for v in pair_w8 pair_w13b2; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so b)" >> $out/ab.jsonl; done
echo "gb[] $(python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl
ncu --set full --clock-control none --import-source on -k regex:pair_count -c 1 --launch-skip 165 -f -o $out/pairb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pairb.log 2>&1
ncu -i $out/pairb.ncu-rep --page raw --csv > $out/pairb_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:pair_count -c 1 --launch-skip 70 -f -o $out/pairs python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pairs.log 2>&1
ncu -i $out/pairs.ncu-rep --page raw --csv > $out/pairs_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:groupby_shard -c 1 --launch-skip 2 -f -o $out/gbh python bench_sweep.py --configs 4 > $out/ncu_gbh.log 2>&1
ncu -i $out/gbh.ncu-rep --page raw --csv > $out/gbh_raw.csv 2>/dev/null
timeout 900 python bench_sweep.py --configs 5 --batched > $out/sweep5.jsonl 2>> $out/bench_err.log
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2>> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log
