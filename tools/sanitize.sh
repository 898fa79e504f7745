#!/bin/bash
# compute-sanitizer evidence (run under gpurun on one B200): memcheck / racecheck / synccheck over the oracle-parity
# tests of the tiny configuration (every kernel of the path incl. the fused persistent decode step, whose tiny shapes
# split EVERY cross-attention item over two CTAs), memcheck + synccheck over one full-size fused generate (batch 32,
# 3 decode steps).  Logs under gpurun_out/<tag>_sanitize_*.log; tools/sanitize_summary.py condenses them for profiles/.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9 --print-limit 20"
T="tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider"
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 900 $SAN --tool $tool $extra python -m pytest $T > gpurun_out/${TAG}_sanitize_${tool}_tiny.log 2>&1
  echo "$tool tiny rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_sanitize_${tool}_tiny.log | tail -3
done
for tool in memcheck synccheck; do
  timeout 900 $SAN --tool $tool python tools/profile_run.py --max-length 4 > gpurun_out/${TAG}_sanitize_${tool}_full.log 2>&1
  echo "$tool full rc=$?"; grep -E "ERROR SUMMARY|kernels" gpurun_out/${TAG}_sanitize_${tool}_full.log | tail -2
done
