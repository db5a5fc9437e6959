# A/B harness used during tuning: build variants with tools/build_variants.py, then e.g.
#   FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_eval_b7.so python bench.py --steps 30 --warmup 3 --no-cpu-baseline
run() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['check_count'])"; }
run default
for f in featurebase_b200/libfbgpu_*.so; do FBGPU_LIB=$PWD/$f run $f; done
