#!/bin/bash
# round-2 call 9: flash v3 (streamed codes), synccheck after the skinny restructure, CPU reference arm at full scale
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py "tests/test_fullsize_gpu.py::test_full_size_encoder_and_greedy_parity" tests/test_decode_paths_gpu.py -q -m gpu > gpurun_out/r2i_pytest_quick.log 2>&1; echo "quick rc=$?"; tail -2 gpurun_out/r2i_pytest_quick.log
timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2i_enc.log 2>&1; echo "enc: $(tail -1 gpurun_out/r2i_enc.log)"
SAN="compute-sanitizer --error-exitcode 9 --print-limit 40"
timeout 300 $SAN --tool synccheck python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2i_sanitize_synccheck_tiny.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2i_sanitize_synccheck_tiny.log | tail -3
timeout 300 $SAN --tool memcheck python -m pytest tests/test_model_gpu.py -k "tiny" -x -q -p no:cacheprovider > gpurun_out/r2i_sanitize_memcheck_tiny.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2i_sanitize_memcheck_tiny.log | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_flash_attn_kernel --launch-skip 5 --launch-count 1 \
  -f -o gpurun_out/r2i_enc_flash python tools/profile_run.py --max-length 3 > gpurun_out/r2i_ncu_f2.log 2>&1; echo "ncu enc_flash rc=$?"
MG_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_bench_ref.log 2>&1; echo "reference arm rc=$?"; tail -1 gpurun_out/r2i_bench_ref.log | cut -c1-900
