for cfg in "MG_SKINNY_CTAS=148" "MG_SKINNY_CTAS=222" "MG_SKINNY_CTAS=296" "MG_SKINNY_CTAS=444"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --max-length 128 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['phases']['decode_step_ms_p50'])"
done
