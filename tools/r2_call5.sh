#!/bin/bash
# round 2, call 5 (first call after the container was re-created: earlier round-2 outputs were lost):
# full gpu test suite, bench.py both arms, sweeps 3/4/5/X/R, ncu --set full of the four hot kernels, launch list
set -u
out=gpurun_out/r2_call5; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $out/summary.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench_err.log; echo "bench ref rc=$?" >> $out/summary.txt
timeout 900 python bench_sweep.py --configs 5 --batched > $out/sweep5.jsonl 2>> $out/bench_err.log; echo "sweep5 rc=$?" >> $out/summary.txt
timeout 600 python bench_sweep.py --configs 3,4,X,R > $out/sweep34.jsonl 2>> $out/bench_err.log; echo "sweep34 rc=$?" >> $out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -c 1 --launch-skip 2 -f -o $out/eval python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_eval.log 2>&1
ncu -i $out/eval.ncu-rep --page raw --csv > $out/eval_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:pair_count_kernel -c 2 --launch-skip 40 -f -o $out/pair python bench_sweep.py --configs 5 --batched --generators uniform --densities 0.01 > $out/ncu_pair.log 2>&1
ncu -i $out/pair.ncu-rep --page raw --csv > $out/pair_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:groupby -c 1 --launch-skip 3 -f -o $out/gbs python bench_sweep.py --configs 4 > $out/ncu_gbs.log 2>&1
ncu -i $out/gbs.ncu-rep --page raw --csv > $out/gbs_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:eval_wordpar -c 1 --launch-skip 6 -f -o $out/wp python bench_sweep.py --configs 3 > $out/ncu_wp.log 2>&1
ncu -i $out/wp.ncu-rep --page raw --csv > $out/wp_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
rm -f $out/*.ncu-rep.tmp
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_gpu.log
