#!/bin/bash
# round-2 profiling call: ncu launch list + full captures of the three new / dominant kernels, synccheck diagnostic, new tests
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches.csv \
  python tools/profile_run.py --max-length 130 > gpurun_out/r2_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel --launch-skip 254 --launch-count 1 \
  -f -o gpurun_out/r2_decode_step python tools/profile_run.py --max-length 260 > gpurun_out/r2_ncu_f1.log 2>&1; echo "ncu decode_step rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_flash_attn_kernel --launch-skip 5 --launch-count 1 \
  -f -o gpurun_out/r2_enc_flash python tools/profile_run.py --max-length 3 > gpurun_out/r2_ncu_f2.log 2>&1; echo "ncu enc_flash rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attn_mma_kernel --launch-skip 6 --launch-count 1 \
  -f -o gpurun_out/r2_window_attn python tools/profile_run.py --max-length 3 > gpurun_out/r2_ncu_f3.log 2>&1; echo "ncu window_attn rc=$?"
SAN="compute-sanitizer --error-exitcode 9 --print-limit 40"
MG_NO_PDL=1 MG_B200_LIB=$PWD/markushgrapher_b200/lib/libmg_b200_nogriddep.so timeout 300 $SAN --tool synccheck python -m pytest tests/test_model_gpu.py -k tiny -x -q -p no:cacheprovider > gpurun_out/r2g_sanitize_synccheck_tiny_nogriddep.log 2>&1; echo "synccheck(no griddep) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2g_sanitize_synccheck_tiny_nogriddep.log | tail -3
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_dist_gpu.py "tests/test_parity_shapes_gpu.py::test_beam_rows_33_to_160_vs_stock_beam_search" -q -m gpu > gpurun_out/r2g_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_new.log
