timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_dist_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed|rror|assert" | tail -6
for cfg in "MG_LANES=1" "MG_LANES=2"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"
done
