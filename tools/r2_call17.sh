#!/bin/bash
# round 2, call 17: groupby_direct_kernel, depth of the descriptor-chain pipeline and the live fourth descriptor register
set -u
out=gpurun_out/r2_call17; mkdir -p $out
g() { python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1; }
for i in 1 2; do
echo "default $(g)" >> $out/gb.jsonl
for v in gd_shpf gd_keepw gd_keepw_shpf gd_deep gd_deep_keepw; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so g)" >> $out/gb.jsonl; done
done
timeout 600 env FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_gd_deep_keepw.so python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "groupby" > $out/pytest_deep.log 2>&1; echo "pytest_deep rc=$?" >> $out/summary.txt
cat $out/summary.txt; cut -c1-330 $out/gb.jsonl
