#!/bin/bash
# compute-sanitizer initcheck: global-memory reads of bytes nobody wrote (a run-ahead encode into freshly allocated,
# non-zero arena chunks changed token ids once: find the reader)
mkdir -p gpurun_out
export MG_ARENA_NOZERO=1
SAN="compute-sanitizer --tool initcheck --print-limit 400 --error-exitcode 9"
timeout 900 $SAN python tools/profile_run.py --max-length 4 > gpurun_out/r2w_initcheck_full.log 2>&1
echo "full rc=$?"
grep -E "ERROR SUMMARY" gpurun_out/r2w_initcheck_full.log | tail -1
grep -A3 "Uninitialized __global__ memory read" gpurun_out/r2w_initcheck_full.log | grep -E "at .*\+0x|at mg::|at void mg::" | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
timeout 900 $SAN python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2w_initcheck_tiny.log 2>&1
echo "tiny rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2w_initcheck_tiny.log | tail -2
grep -A3 "Uninitialized __global__ memory read" gpurun_out/r2w_initcheck_tiny.log | grep -E "at .*\+0x|at mg::|at void mg::" | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
