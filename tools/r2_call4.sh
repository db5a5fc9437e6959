#!/bin/bash
# round 2, call 4: bench.py (fixed), new GroupBy kernel, word-parallel ring depth, any / pair_types tests, A/B of the bitmap x bitmap batched regression
set -u
out=gpurun_out/r2_call4; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node.py tests/test_zz_gpu_experimental.py -x -q -k "any_early or groupby or node or every_call or concurrent or failing or bounded or bsi or various" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench_err.log; echo "bench rc=$?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench_err.log; echo "bench ref rc=$?" >> $out/summary.txt
python bench_sweep.py --configs 3,4 > $out/sweep34.jsonl 2>> $out/bench_err.log
FBGPU_GROUPBY_CTA=1 python bench_sweep.py --configs 4 > $out/sweep4_cta.jsonl 2>> $out/bench_err.log
FBGPU_LAYOUT_SLOT_MAJOR=1 python bench_sweep.py --configs 4 > $out/sweep4_slot.jsonl 2>> $out/bench_err.log
for v in wp_ring3 wp_ring8 wp_legacy; do FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so python bench_sweep.py --configs 3 > $out/sweep3_$v.jsonl 2>> $out/bench_err.log; done
# bitmap x bitmap batched: round-1 library vs now
python bench_sweep.py --configs 5 --batched --generators uniform --densities 0.125,0.01 > $out/sweep5_now.jsonl 2>> $out/bench_err.log
FBGPU_ARRAY_SORTED=1 FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_r1.so python bench_sweep.py --configs 5 --batched --generators uniform --densities 0.125,0.01 > $out/sweep5_r1.jsonl 2>> $out/bench_err.log
ncu --set full --clock-control none -k regex:groupby_small_kernel -c 1 --launch-skip 3 -f -o $out/gbs python bench_sweep.py --configs 4 > $out/ncu_gbs.log 2>&1
ncu -i $out/gbs.ncu-rep --page raw --csv > $out/gbs_raw.csv 2>/dev/null
cat $out/summary.txt; tail -3 $out/pytest_sel.log
