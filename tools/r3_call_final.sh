#!/bin/bash
# end-of-round evidence (one B200): bench lines of the measured workloads, ncu launch list, ncu --set full of the fused step
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r3_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r3_bench.log | cut -c1-400
timeout 300 python bench.py --workload enc256 > gpurun_out/r3_enc256.log 2>&1; echo "enc256 rc=$?"; tail -1 gpurun_out/r3_enc256.log | cut -c1-300
timeout 400 python bench.py --workload gen128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_gen128.log 2>&1; echo "gen128 rc=$?"; tail -1 gpurun_out/r3_gen128.log | cut -c1-300
timeout 600 python bench.py --workload beam4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3_beam4.log 2>&1; echo "beam4 rc=$?"; tail -1 gpurun_out/r3_beam4.log | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r3_launches.csv \
  python tools/profile_run.py --max-length 130 > gpurun_out/r3_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel --launch-skip 254 --launch-count 1 \
  -f -o gpurun_out/r3_decode_step python tools/profile_run.py --max-length 260 > gpurun_out/r3_ncu_f1.log 2>&1; echo "ncu decode_step rc=$?"
timeout 200 python tools/ab_ahead.py --batches 3 > gpurun_out/r3_ahead.log 2>&1; tail -4 gpurun_out/r3_ahead.log
