#!/bin/bash
# round-2 third GPU call: tensor-core Swin window attention, skinny statistics rewrite, full GPU suite, bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -s -k "encoder or greedy_token" > gpurun_out/r2c_pytest_quick.log 2>&1; echo "quick rc=$?"; grep -E "passed|failed|rel err|FAILED" gpurun_out/r2c_pytest_quick.log | tail -8
timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2c_enc_mma.log 2>&1; tail -1 gpurun_out/r2c_enc_mma.log
MG_SWIN_ATTN=fma timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2c_enc_fma.log 2>&1; tail -1 gpurun_out/r2c_enc_fma.log
SAN="compute-sanitizer --error-exitcode 9 --print-limit 40"
timeout 300 $SAN --tool synccheck python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2c_sanitize_synccheck_tiny.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2c_sanitize_synccheck_tiny.log | tail -3
timeout 300 $SAN --tool memcheck python -m pytest tests/test_model_gpu.py -k "tiny" -x -q -p no:cacheprovider > gpurun_out/r2c_sanitize_memcheck_tiny.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2c_sanitize_memcheck_tiny.log | tail -3
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/r2c_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2c_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2c_bench.log | cut -c1-1800
