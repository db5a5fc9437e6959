#!/bin/bash
# round 2, second final measurement call (after groupby_direct_kernel): ncu --set full of eval_kernel first (-> profiles/traffic.json for THIS
# kernel source, so that the bench line below carries roofline.traffic), the whole gpu test suite, bench.py both arms, the config 3/4/X/R
# sweep, ncu --set full of groupby_direct_kernel, the launch list of one bench run
set -u
out=gpurun_out/r2_final2; mkdir -p $out
ncu --set full --clock-control none -k regex:eval_kernel -c 2 --launch-skip 2 -f -o $out/eval python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_eval.log 2>&1
ncu -i $out/eval.ncu-rep --page raw --csv > $out/eval_raw.csv 2>/dev/null
python tools/make_traffic_json.py $out/eval_raw.csv 1391112042 "profiles/r02_ncu_eval_kernel_summary.csv (ncu --set full, bench.py configs[1], tools/r2_final2.sh)" > $out/traffic_make.log 2>&1; cp profiles/traffic.json $out/traffic.json
timeout 2400 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $out/summary.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench_err.log; echo "bench ref rc=$?" >> $out/summary.txt
timeout 600 python bench_sweep.py --configs 3,4,X,R > $out/sweep34.jsonl 2>> $out/bench_err.log
ncu --set full --clock-control none -k regex:groupby_direct -c 1 --launch-skip 2 -f -o $out/gbd python bench_sweep.py --configs 4 > $out/ncu_gbd.log 2>&1
ncu -i $out/gbd.ncu-rep --page raw --csv > $out/gbd_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep > $out/launches_bench.log 2>&1
rm -f $out/*.ncu-rep
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_gpu.log
