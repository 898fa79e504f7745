#!/bin/bash
# round-2 call 8: flash kernel v2 (byte codes, ex2.approx, padded-row skip), Swin occupancy A/B, compaction A/B, full suite, bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py "tests/test_fullsize_gpu.py::test_full_size_encoder_and_greedy_parity" "tests/test_decode_paths_gpu.py::test_masked_memory_positions_are_dropped_without_changing_results" "tests/test_parity_shapes_gpu.py::test_edge_inputs_sep_box_1000_out_of_range_boxes_padding_row" -q -m gpu -s > gpurun_out/r2h_pytest_quick.log 2>&1; echo "quick rc=$?"; grep -E "passed|failed|rel err|FAILED|Error" gpurun_out/r2h_pytest_quick.log | tail -12
MG_FLASH_GAP=0 timeout 300 python -m pytest tests/test_model_gpu.py "tests/test_fullsize_gpu.py::test_full_size_encoder_and_greedy_parity" -q -m gpu -k "encoder" > gpurun_out/r2h_pytest_gap0.log 2>&1; echo "gap0 rc=$?"; tail -2 gpurun_out/r2h_pytest_gap0.log
for v in 2 1; do MG_SWIN_MINB=$v timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2h_enc_minb$v.log 2>&1; echo "minb $v: $(tail -1 gpurun_out/r2h_enc_minb$v.log)"; done
timeout 300 python tools/ab_env.py --settings "" "MG_COMPACT=0" "" "MG_COMPACT=0" > gpurun_out/r2h_ab_compact.log 2>&1; tail -4 gpurun_out/r2h_ab_compact.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2h_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2h_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2h_bench.log | cut -c1-1200
timeout 600 python bench.py --workload enc256 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_enc256.log 2>&1; echo "enc256 rc=$?"; tail -1 gpurun_out/r2h_enc256.log | cut -c1-400
