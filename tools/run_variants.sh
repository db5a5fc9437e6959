for v in v1 v3 v3_b5 v2_free; do
  FBGPU_LIB=$PWD/featurebase_b200/libfbgpu_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['check_count'])"
done
