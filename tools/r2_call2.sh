#!/bin/bash
# round 2, call 2: eval_kernel load-depth / residency A/B (striped + sorted), staged+striped, ncu of the pair micro-kernels
set -u
out=gpurun_out/r2_call2; mkdir -p $out
bench() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>>$out/bench_err.log | tail -1; }
for v in "" eval_deep1 eval_mb6 eval_mb5 eval_mb6_deep4; do
  lib=featurebase_b200/libfbgpu${v:+_$v}.so
  echo "sorted ${v:-default} $(FBGPU_LIB=$PWD/$lib bench)" >> $out/bench.jsonl
  echo "striped ${v:-default} $(FBGPU_LIB=$PWD/$lib FBGPU_ARRAY_STRIPED=1 bench)" >> $out/bench.jsonl
done
echo "striped staged $(FBGPU_STAGED=1 FBGPU_ARRAY_STRIPED=1 bench)" >> $out/bench.jsonl
echo "striped staged3 $(FBGPU_STAGED=1 FBGPU_STAGE_CTAS=3 FBGPU_ARRAY_STRIPED=1 bench)" >> $out/bench.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $out/summary.txt
FBGPU_ARRAY_STRIPED=1 FBGPU_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_gpu_experimental.py -x -q -k stripe > $out/pytest_striped.log 2>&1; echo "pytest striped rc=$?" >> $out/summary.txt
cd bench_micro
for k in "w1 wipe 8w x3 striped" "w3 unscatter+prefetch 8w x3 striped" "c1 coop T=96 x21 striped"; do
  tag=$(echo "$k" | tr ' +=' '___')
  timeout 600 ncu --set full --clock-control none -c 1 --launch-skip 1 -f -o ../$out/micro_$tag ./pair_variants 65536 0.01 "$k" > ../$out/ncu_$tag.log 2>&1
  ncu -i ../$out/micro_$tag.ncu-rep --page raw --csv > ../$out/micro_${tag}_raw.csv 2>/dev/null
done
cd ..
python - <<'PY' >> gpurun_out/r2_call2/summary.txt
import json
for line in open("gpurun_out/r2_call2/bench.jsonl"):
    a, b, js = line.split(" ", 2)
    try:
        d = json.loads(js); print(a, b, "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "count", d.get("check_count"))
    except Exception as e:
        print(a, b, "unparsed:", e)
PY
cat $out/summary.txt
