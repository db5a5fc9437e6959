run() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['e2e']['ms_per_step'],4), d['check_count'], d['clocks'])"; }
run default_prefetch
FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_nopf.so run noprefetch
