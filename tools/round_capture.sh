#!/bin/bash
# End-of-round evidence on one B200 (run under gpurun): GPU test suite, the bench line, the ncu launch list of a short
# generate, one ncu --set full capture of the fused decode step at cached length 255.  Outputs under gpurun_out/.
set -u
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/profile_run.py --max-length 130 > gpurun_out/${TAG}_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel --launch-skip 254 --launch-count 1 \
  -f -o gpurun_out/${TAG}_decode_step python tools/profile_run.py --max-length 260 > gpurun_out/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_*
