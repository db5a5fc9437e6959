#!/bin/bash
# round 2, call 16: groupby_direct_kernel, the two descriptor-chain changes of call 14 one at a time
set -u
out=gpurun_out/r2_call16; mkdir -p $out
g() { python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1; }
for i in 1 2; do
echo "default $(g)" >> $out/gb.jsonl
for v in gd_shpf gd_desc12; do echo "$v $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so g)" >> $out/gb.jsonl; done
done
cut -c1-330 $out/gb.jsonl
