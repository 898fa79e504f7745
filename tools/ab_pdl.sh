for cfg in "MG_NO_PDL=0 MG_SKINNY_STAGES=2" "MG_NO_PDL=1 MG_SKINNY_STAGES=2" "MG_NO_PDL=0 MG_SKINNY_STAGES=4" "MG_NO_PDL=1 MG_SKINNY_STAGES=4"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --max-length 128 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['phases']['decode_step_ms_p50'], d['roofline']['ms_per_launch'])"
done
MG_DUMP_GRAPH=gpurun_out/graph.dot timeout 300 python tools/profile_run.py --max-length 3 > /dev/null 2>&1; grep -c "PROGRAMMATIC\|programmatic" gpurun_out/graph.dot; grep -m3 -i "programmatic" gpurun_out/graph.dot; ls -la gpurun_out/graph.dot | head -2
