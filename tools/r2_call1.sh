#!/bin/bash
# round-2 first GPU call: new parity tests, sanitizer, bench line, A/B of the two flagged builds
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests/test_parity_shapes_gpu.py tests/test_handoff_gpu.py tests/test_model_gpu.py -q -m gpu -s > gpurun_out/r2a_pytest_new.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|parity|rel err|B=" gpurun_out/r2a_pytest_new.log | tail -30
bash tools/sanitize.sh r2a
timeout 900 python bench.py > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2a_bench.log | cut -c1-3000
L=$PWD/markushgrapher_b200/lib
timeout 300 python tools/ab_env.py --settings "" "MG_MEGA_L2PF=0" > gpurun_out/r2a_ab_default.log 2>&1; tail -2 gpurun_out/r2a_ab_default.log
MG_B200_LIB=$L/libmg_b200_selfall.so timeout 300 python tools/ab_env.py --settings "" > gpurun_out/r2a_ab_selfall.log 2>&1; tail -1 gpurun_out/r2a_ab_selfall.log
MG_B200_LIB=$L/libmg_b200_pflane.so timeout 400 python tools/ab_env.py --settings "MG_MEGA_L2PF_MASK=0,MG_MEGA_L2PF=0" "MG_MEGA_L2PF_MASK=0,MG_MEGA_L2PF=512,MG_MEGA_L2PF_PIECE=4096,MG_MEGA_L2PF_GAP=400" "MG_MEGA_L2PF_MASK=0,MG_MEGA_L2PF=1024,MG_MEGA_L2PF_PIECE=8192,MG_MEGA_L2PF_GAP=400" "MG_MEGA_L2PF_MASK=0,MG_MEGA_L2PF=1536,MG_MEGA_L2PF_PIECE=16384,MG_MEGA_L2PF_GAP=500" "MG_MEGA_L2PF=384" > gpurun_out/r2a_ab_pflane.log 2>&1; tail -5 gpurun_out/r2a_ab_pflane.log
