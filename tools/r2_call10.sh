#!/bin/bash
# round-2 call 10: flag-array grid barrier A/B, flash v2 ncu capture
set -u
mkdir -p gpurun_out
L=$PWD/markushgrapher_b200/lib
MG_B200_LIB=$L/libmg_b200_flagbar.so timeout 300 python -m pytest tests/test_model_gpu.py tests/test_decode_paths_gpu.py -q -m gpu -k "greedy or paths or kernel_chain or masked" > gpurun_out/r2j_pytest_flagbar.log 2>&1; echo "pytest flagbar rc=$?"; tail -2 gpurun_out/r2j_pytest_flagbar.log
timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2j_ab_default.log 2>&1; tail -1 gpurun_out/r2j_ab_default.log
MG_B200_LIB=$L/libmg_b200_flagbar.so timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2j_ab_flagbar.log 2>&1; tail -1 gpurun_out/r2j_ab_flagbar.log
timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2j_ab_default2.log 2>&1; tail -1 gpurun_out/r2j_ab_default2.log
MG_B200_LIB=$L/libmg_b200_flagbar.so timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2j_ab_flagbar2.log 2>&1; tail -1 gpurun_out/r2j_ab_flagbar2.log
timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2j_enc.log 2>&1; echo "enc: $(tail -1 gpurun_out/r2j_enc.log)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_flash_attn_kernel --launch-skip 5 --launch-count 1 \
  -f -o gpurun_out/r2j_enc_flash python tools/profile_run.py --max-length 3 > gpurun_out/r2j_ncu_f2.log 2>&1; echo "ncu enc_flash rc=$?"
