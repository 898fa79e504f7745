#!/bin/bash
# ncu launch lists of the two large-batch workloads (kernel chain): configs[3] 128 images greedy, configs[4] 125 images x beam 4
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r3_launches_gen128.csv \
  python bench.py --workload gen128 --steps 1 --warmup 0 --max-length 130 --no-cpu-baseline > gpurun_out/r3_l_gen128.log 2>&1; echo "gen128 rc=$?"
python tools/summarize_launches.py gpurun_out/r3_launches_gen128.csv | head -25
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r3_launches_beam4.csv \
  python bench.py --workload beam4 --steps 1 --warmup 0 --max-length 130 --no-cpu-baseline > gpurun_out/r3_l_beam4.log 2>&1; echo "beam4 rc=$?"
python tools/summarize_launches.py gpurun_out/r3_launches_beam4.csv | head -25
