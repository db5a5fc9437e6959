#!/bin/bash
# First GPU call of round 2 (one gpurun invocation, about 25 GPU-minutes): run everything that was written after round 1's
# GPU budget was spent, and measure every pending performance change (one A/B each).  Results land in gpurun_out/r2_first/.
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/r2_first_call.sh'
# If the budget is tight, the sections are independent: comment out from the bottom up.
set -u
out=gpurun_out/r2_first; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/build.log 2>&1

# 1. the regular parity suite (covers the LEA.HI / hoisted-base scatter that is now the default build)
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $out/summary.txt

# 2. opt-in tests: striped arrays, RBF loader, BSI aggregates, fragment.top / filter / archetype goldens
FBGPU_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_gpu_experimental.py -q > $out/pytest_experimental.log 2>&1; echo "pytest_experimental rc=$?" >> $out/summary.txt

# 2b. (only when the call was made with `gpurun --gpus 2`) the two-GPU test: fused Count merge, ncclAllReduce of count vectors (TopN ids / GroupBy)
[ "$(nvidia-smi -L | wc -l)" -ge 2 ] && { timeout 600 python -m pytest tests/test_gpu_multi.py -q > $out/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $out/summary.txt; }

# 3. headline bench: default layout vs bank-striped arrays (same build), then the density sweep for both
bench() { python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>>$out/bench_err.log | tail -1; }
echo "default $(bench)" >> $out/bench.jsonl
echo "striped $(FBGPU_ARRAY_STRIPED=1 bench)" >> $out/bench.jsonl
python bench_sweep.py --configs 5 --batched --generators uniform > $out/sweep_default.jsonl 2>>$out/bench_err.log
FBGPU_ARRAY_STRIPED=1 python bench_sweep.py --configs 5 --batched --generators uniform > $out/sweep_striped.jsonl 2>>$out/bench_err.log

# 3b. experimental op loop of the word-parallel kernel (BSI): parity of the BSI tests with the variant library, then config 3
python tools/build_variants.py wp_unroll3 > $out/build_variants.log 2>&1
FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_wp_unroll3.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "bsi or alternative or executor_goldens" > $out/pytest_wp_unroll3.log 2>&1; echo "pytest_wp_unroll3 rc=$?" >> $out/summary.txt
python bench_sweep.py --configs 3 > $out/sweep3_default.jsonl 2>>$out/bench_err.log
FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_wp_unroll3.so python bench_sweep.py --configs 3 > $out/sweep3_wp_unroll3.jsonl 2>>$out/bench_err.log

# 3b'. pair_count_kernel without the per-pair 8 KiB wipe (-DFBGPU_PAIR_UNSCATTER): parity, then the density sweep in both array orders
python tools/build_variants.py pair_unscatter >> $out/build_variants.log 2>&1
PU=$PWD/featurebase_b200/libfbgpu_pair_unscatter.so
FBGPU_LIB=$PU timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "config1 or density_sweep or mixed_encoding or topk or executor_goldens or full_size_properties_1024" > $out/pytest_pair_unscatter.log 2>&1; echo "pytest_pair_unscatter rc=$?" >> $out/summary.txt
FBGPU_LIB=$PU python bench_sweep.py --configs 5 --batched --generators uniform > $out/sweep_pair_unscatter.jsonl 2>>$out/bench_err.log
FBGPU_LIB=$PU FBGPU_ARRAY_STRIPED=1 python bench_sweep.py --configs 5 --batched --generators uniform > $out/sweep_pair_unscatter_striped.jsonl 2>>$out/bench_err.log

# 3c. thread-per-row GroupBy passes (groupby_kernel<true>, FBGPU_GROUPBY_FAST=1): parity of every GroupBy test, then config 4 A/B
FBGPU_GROUPBY_FAST=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_experimental.py -q -k "groupby or various_queries" > $out/pytest_groupby_fast.log 2>&1; echo "pytest_groupby_fast rc=$?" >> $out/summary.txt
python bench_sweep.py --configs 4 > $out/sweep4_default.jsonl 2>>$out/bench_err.log
FBGPU_GROUPBY_FAST=1 python bench_sweep.py --configs 4 > $out/sweep4_groupby_fast.jsonl 2>>$out/bench_err.log
# (payloads of one slot's rows adjacent: tiny containers of a unit then share cache lines)
FBGPU_LAYOUT_SLOT_MAJOR=1 python bench_sweep.py --configs 4 > $out/sweep4_slot_major.jsonl 2>>$out/bench_err.log
FBGPU_LAYOUT_SLOT_MAJOR=1 FBGPU_GROUPBY_FAST=1 python bench_sweep.py --configs 4 > $out/sweep4_groupby_fast_slot_major.jsonl 2>>$out/bench_err.log

# 3d. the two result-expansion kernels (fbgpu_columns / fbgpu_extract): first timing, checked against the generator inside the script
python bench_sweep.py --configs X,R > $out/sweep_columns_extract.jsonl 2>>$out/bench_err.log

# 4. one ncu pass of the headline kernel in both layouts: shared-memory wavefronts / issue utilisation are what changed
for mode in default striped; do
  env $( [ $mode = striped ] && echo FBGPU_ARRAY_STRIPED=1 ) ncu --set full --clock-control none -k regex:eval_kernel -c 1 -f -o $out/eval_$mode \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_$mode.log 2>&1
  # keep the raw-metrics CSV (small); the .ncu-rep only if it fits the 64 MiB return budget comfortably
  ncu -i $out/eval_$mode.ncu-rep --page raw --csv > $out/eval_${mode}_raw.csv 2>/dev/null
  [ "$(stat -c %s $out/eval_$mode.ncu-rep 2>/dev/null || echo 0)" -gt 20000000 ] && rm -f $out/eval_$mode.ncu-rep
done
python - <<'PY' >> gpurun_out/r2_first/summary.txt
import json
for line in open("gpurun_out/r2_first/bench.jsonl"):
    name, _, js = line.partition(" ")
    try:
        d = json.loads(js)
        print(name, "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "count", d.get("check_count"))
    except Exception as e:
        print(name, "unparsed:", e)
PY
cat $out/summary.txt
