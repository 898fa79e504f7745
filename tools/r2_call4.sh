#!/bin/bash
# round-2 fourth GPU call: synccheck without PDL, folded-FF variant (parity + A/B), configs[3]/[4] single-GPU lines
set -u
mkdir -p gpurun_out
L=$PWD/markushgrapher_b200/lib
SAN="compute-sanitizer --error-exitcode 9 --print-limit 40"
MG_NO_PDL=1 timeout 300 $SAN --tool synccheck python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2d_sanitize_synccheck_tiny_nopdl.log 2>&1; echo "synccheck(no PDL) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2d_sanitize_synccheck_tiny_nopdl.log | tail -3
MG_B200_LIB=$L/libmg_b200_foldff.so timeout 900 python -m pytest tests/test_model_gpu.py tests/test_decode_paths_gpu.py tests/test_golden_gpu.py "tests/test_parity_shapes_gpu.py::test_bench_workload_full_511_steps_rows_vs_oracle" "tests/test_parity_shapes_gpu.py::test_greedy_early_eos_ragged_finish" -q -m gpu -s > gpurun_out/r2d_pytest_foldff.log 2>&1; echo "pytest foldff rc=$?"; grep -E "passed|failed|parity|FAILED|Error" gpurun_out/r2d_pytest_foldff.log | tail -12
timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2d_ab_default.log 2>&1; tail -1 gpurun_out/r2d_ab_default.log
MG_B200_LIB=$L/libmg_b200_foldff.so timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2d_ab_foldff.log 2>&1; tail -1 gpurun_out/r2d_ab_foldff.log
timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2d_ab_default2.log 2>&1; tail -1 gpurun_out/r2d_ab_default2.log
MG_B200_LIB=$L/libmg_b200_foldff.so timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2d_ab_foldff2.log 2>&1; tail -1 gpurun_out/r2d_ab_foldff2.log
timeout 600 python bench.py --workload gen128 --steps 2 --warmup 1 > gpurun_out/r2d_gen128.log 2>&1; echo "gen128 rc=$?"; tail -1 gpurun_out/r2d_gen128.log | cut -c1-1500
timeout 900 python bench.py --workload beam4 --steps 1 --warmup 1 > gpurun_out/r2d_beam4.log 2>&1; echo "beam4 rc=$?"; tail -1 gpurun_out/r2d_beam4.log | cut -c1-1500
