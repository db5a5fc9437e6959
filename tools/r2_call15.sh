#!/bin/bash
# round 2, call 15: groupby_direct_kernel variants on config 4 (per-column loops vs eight-column steps, descriptor-chain forms, probe byte load)
set -u
out=gpurun_out/r2_call15; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "groupby" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
g() { python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1; }
for i in 1 2; do
echo "default $(g)" >> $out/gb.jsonl
for v in gd_oldpipe gd_tabload gd_oldpipe_tabload gd_chunk gd_nopipe; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so g)" >> $out/gb.jsonl; done
done
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log; cut -c1-330 $out/gb.jsonl
