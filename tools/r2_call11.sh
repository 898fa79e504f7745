#!/bin/bash
# round-2 call 11: flash kernel cache-policy / prefetch switches (MG_FLASH_OPT bits), encoder time per setting
set -u
mkdir -p gpurun_out
for o in 0 1 2 3 4 5 7; do MG_FLASH_OPT=$o timeout 200 python tools/profile_run.py --max-length 4 --reps 3 > gpurun_out/r2k_enc_opt$o.log 2>&1; echo "opt $o: $(tail -1 gpurun_out/r2k_enc_opt$o.log)"; done
MG_FLASH_OPT=7 timeout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -k encoder > gpurun_out/r2k_pytest_opt7.log 2>&1; echo "pytest opt7 rc=$?"; tail -1 gpurun_out/r2k_pytest_opt7.log
