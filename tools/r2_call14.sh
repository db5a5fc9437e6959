#!/bin/bash
# round 2, call 14: groupby_direct_kernel (byte table per (shard, slot)) against groupby_shard_kernel (hash table per group of slots) on
# config 4; with and without the pipelined descriptor chain; ncu --set full of the new kernel; launch list of one config-4 run
set -u
out=gpurun_out/r2_call14; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "groupby" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
g() { python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1; }
for i in 1 2; do
echo "direct_pipe $(g)" >> $out/gb.jsonl
echo "direct_nopipe $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_gd_nopipe.so g)" >> $out/gb.jsonl
echo "hash $(FBGPU_GROUPBY_HASH=1 g)" >> $out/gb.jsonl
done
ncu --set full --clock-control none --import-source on -k regex:groupby_direct -c 1 --launch-skip 2 -f -o $out/gbd python bench_sweep.py --configs 4 > $out/ncu_gbd.log 2>&1
ncu -i $out/gbd.ncu-rep --page raw --csv > $out/gbd_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv python bench_sweep.py --configs 4 > $out/launches_run.log 2>&1
python bench.py --steps 20 --warmup 3 --extras config4 > $out/bench.json 2>>$out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log; cat $out/gb.jsonl | cut -c1-400
