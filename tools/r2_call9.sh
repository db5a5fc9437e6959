#!/bin/bash
# round 2, call 9: pair kernel after the 64-bit share fix and the shared-memory per-pair sums; eval_kernel directory prefetch A/B
set -u
out=gpurun_out/r2_call9; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1 or padded or density_sweep or mixed_encoding or any_early or thread_safety or container_combinations or full_size_properties_1024 or topk" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
echo "default2 $(b)" >> $out/ab.jsonl
for v in eval_nopf; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so b)" >> $out/ab.jsonl; done
echo "eval_nopf2 $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_eval_nopf.so b)" >> $out/ab.jsonl
ncu --set full --clock-control none --import-source on -k regex:pair_count -c 1 --launch-skip 165 -f -o $out/pairb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pairb.log 2>&1
ncu -i $out/pairb.ncu-rep --page raw --csv > $out/pairb_raw.csv 2>/dev/null
timeout 900 python bench_sweep.py --configs 5 --batched > $out/sweep5.jsonl 2>> $out/bench_err.log
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2>> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log
