#!/bin/bash
# round 2, call 6: the fused pair kernel is reachable again (program-rewrite regression), IMAD.HI address arithmetic vs the ALU form,
# L2 prefetch depth of the pair kernel, GroupBy kernels one by one, pipe-rate and pair micro-benchmarks
set -u
out=gpurun_out/r2_call6; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -x -q -m gpu -k "config1 or density_sweep or mixed_encoding or full_size_every or incremental or arena_compaction or container_combinations or groupby" > $out/pytest_sel.log 2>&1; echo "pytest_sel rc=$?" >> $out/summary.txt
b() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras north_star 2>>$out/bench_err.log | tail -1; }
echo "default $(b)" >> $out/ab.jsonl
echo "addr_alu $(FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_addr_alu.so b)" >> $out/ab.jsonl
for d in 1 2 8 15; do echo "pf$d $(FBGPU_PAIR_PF=$d b)" >> $out/ab.jsonl; done
echo "nopair $(FBGPU_NO_PAIR_KERNEL=1 b)" >> $out/ab.jsonl
for v in "" "FBGPU_GROUPBY_CTA=1" "FBGPU_GROUPBY_CTA=1 FBGPU_GROUPBY_FAST=1"; do echo "gb[$v] $(env $v python bench_sweep.py --configs 4 2>>$out/bench_err.log | tail -1)" >> $out/gb.jsonl; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:groupby -c 12 --csv --log-file $out/gb_launches.csv python bench_sweep.py --configs 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:groupby_small -c 1 --launch-skip 2 -f -o $out/gbs python bench_sweep.py --configs 4 > $out/ncu_gbs.log 2>&1
ncu -i $out/gbs.ncu-rep --page raw --csv > $out/gbs_raw.csv 2>/dev/null
ncu -i $out/gbs.ncu-rep --page source --csv > $out/gbs_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:pair_count -c 2 --launch-skip 70 -f -o $out/pair python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extras north_star > $out/ncu_pair.log 2>&1
ncu -i $out/pair.ncu-rep --page raw --csv > $out/pair_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:eval_kernel -c 1 --launch-skip 2 -f -o $out/eval python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_eval.log 2>&1
ncu -i $out/eval.ncu-rep --page raw --csv > $out/eval_raw.csv 2>/dev/null
./bench_micro/pipe_rates > $out/pipe_rates.txt 2>&1
(cd bench_micro && timeout 300 ./pair_variants 65536 0.01 > ../$out/pair_variants.txt 2>&1)
ls -la $out >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_sel.log; cat $out/pipe_rates.txt
