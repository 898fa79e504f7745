#!/bin/bash
# round-2 second GPU call: fused encoder attention, B>32 fix, kv24 beams, sanitizer re-runs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_shapes_gpu.py tests/test_golden_gpu.py tests/test_dist_gpu.py "tests/test_fullsize_gpu.py::test_full_size_encoder_and_greedy_parity" tests/test_configs_gpu.py -q -m gpu -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|parity|rel err|B=|FAILED|Error" gpurun_out/r2b_pytest.log | tail -40
MG_FLASH_GAP=0 timeout 600 python -m pytest tests/test_model_gpu.py "tests/test_fullsize_gpu.py::test_full_size_encoder_and_greedy_parity" -q -m gpu -s -k "encoder" > gpurun_out/r2b_pytest_gap0.log 2>&1; echo "pytest gap0 rc=$?"; grep -E "passed|failed|rel err|FAILED" gpurun_out/r2b_pytest_gap0.log | tail -8
timeout 300 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2b_enc_flash.log 2>&1; tail -1 gpurun_out/r2b_enc_flash.log
MG_ENC_ATTN=unfused timeout 300 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2b_enc_unfused.log 2>&1; tail -1 gpurun_out/r2b_enc_unfused.log
timeout 600 python bench.py --workload enc256 --steps 2 --warmup 1 > gpurun_out/r2b_enc256.log 2>&1; echo "enc256 rc=$?"; tail -1 gpurun_out/r2b_enc256.log | cut -c1-2500
SAN="compute-sanitizer --error-exitcode 9 --print-limit 40"
timeout 300 $SAN --tool synccheck python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2b_sanitize_synccheck_tiny.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2b_sanitize_synccheck_tiny.log | tail -3
timeout 300 $SAN --tool memcheck python -m pytest tests/test_model_gpu.py -k "tiny and (encoder or greedy)" -x -q -p no:cacheprovider > gpurun_out/r2b_sanitize_memcheck_tiny.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2b_sanitize_memcheck_tiny.log | tail -3
MG_B200_LIB=$PWD/markushgrapher_b200/lib/libmg_b200_race.so timeout 600 $SAN --tool racecheck --racecheck-report all python -m pytest "tests/test_model_gpu.py::test_greedy_token_identical[tiny-2-12-24]" -x -q -p no:cacheprovider > gpurun_out/r2b_sanitize_racecheck_tiny.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2b_sanitize_racecheck_tiny.log | tail -3
