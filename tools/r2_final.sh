#!/bin/bash
# round 2, final measurement call: the whole gpu test suite, bench.py both arms, all sweeps, ncu --set full of the hot kernels of the
# timed binary (-> profiles/r02_ncu_*_summary.csv, profiles/traffic.json) and the launch list of one bench run
set -u
out=gpurun_out/r2_final; mkdir -p $out
timeout 2400 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $out/summary.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench_err.log; echo "bench ref rc=$?" >> $out/summary.txt
timeout 900 python bench_sweep.py --configs 5 --batched > $out/sweep5.jsonl 2>> $out/bench_err.log
timeout 600 python bench_sweep.py --configs 3,4,X,R > $out/sweep34.jsonl 2>> $out/bench_err.log
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -c 2 --launch-skip 2 -f -o $out/eval python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_eval.log 2>&1
ncu -i $out/eval.ncu-rep --page raw --csv > $out/eval_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:pair_count -c 1 --launch-skip 165 -f -o $out/pairb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pairb.log 2>&1
ncu -i $out/pairb.ncu-rep --page raw --csv > $out/pairb_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:pair_count -c 1 --launch-skip 70 -f -o $out/pairs python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pairs.log 2>&1
ncu -i $out/pairs.ncu-rep --page raw --csv > $out/pairs_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:groupby_shard -c 1 --launch-skip 2 -f -o $out/gbh python bench_sweep.py --configs 4 > $out/ncu_gbh.log 2>&1
ncu -i $out/gbh.ncu-rep --page raw --csv > $out/gbh_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:eval_wordpar -c 1 --launch-skip 6 -f -o $out/wp python bench_sweep.py --configs 3 > $out/ncu_wp.log 2>&1
ncu -i $out/wp.ncu-rep --page raw --csv > $out/wp_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep > $out/launches_bench.log 2>&1
rm -f $out/*.ncu-rep
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_gpu.log
