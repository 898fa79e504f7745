#!/bin/bash
# round-2 (second half) checks after the cross-attention rewrite / row-scale kernel / run-ahead: memcheck + synccheck on the
# tiny parity tests (incl. the run-ahead tests), memcheck on one full-size fused generate, racecheck on one tiny test with the
# -DMK_RACECHECK build; A/B of the wide-batch workloads against the previous build
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
T="tests/test_model_gpu.py tests/test_ahead_gpu.py -k tiny -x -q -p no:cacheprovider"
for tool in memcheck synccheck; do
  timeout 600 $SAN --tool $tool python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r3_sanitize_${tool}_tiny.log 2>&1
  echo "$tool tiny rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3_sanitize_${tool}_tiny.log | tail -2
done
timeout 600 $SAN --tool memcheck python -m pytest tests/test_ahead_gpu.py -x -q -p no:cacheprovider > gpurun_out/r3_sanitize_memcheck_ahead.log 2>&1
echo "memcheck ahead rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3_sanitize_memcheck_ahead.log | tail -2
timeout 600 $SAN --tool memcheck python tools/profile_run.py --max-length 4 > gpurun_out/r3_sanitize_memcheck_full.log 2>&1
echo "memcheck full rc=$?"; grep -E "ERROR SUMMARY|kernels" gpurun_out/r3_sanitize_memcheck_full.log | tail -2
if [ -f markushgrapher_b200/lib/libmg_b200_racecheck.so ]; then
  MG_B200_LIB=$PWD/markushgrapher_b200/lib/libmg_b200_racecheck.so timeout 900 $SAN --tool racecheck --racecheck-report all python -m pytest "tests/test_model_gpu.py::test_greedy_token_identical" -k "tiny and 2-12-24" -x -q -p no:cacheprovider > gpurun_out/r3_sanitize_racecheck_tiny.log 2>&1
  echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r3_sanitize_racecheck_tiny.log | tail -2
fi
for lib in prev cur; do
  L=$PWD/markushgrapher_b200/lib/libmg_b200.so; [ $lib = prev ] && L=$PWD/markushgrapher_b200/lib/libmg_b200_prev.so
  for w in gen128 beam4; do MG_B200_LIB=$L timeout 300 python bench.py --workload $w --steps 1 --warmup 1 --max-length 130 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(\"$lib $w\", round(d[\"value\"],2), \"img/s; step ms\", round(d[\"phases\"][\"decode_step_ms_mean\"],3))" | tee -a gpurun_out/r3_wide_ab.log; done
done
