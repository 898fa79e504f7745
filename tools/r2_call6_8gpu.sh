#!/bin/bash
# round-2 8-GPU call: BASELINE.json configs[1] (8 x 32), configs[3] (8 x 128 greedy) and configs[4] (8 x 125, beam 4, <= 768 tok)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 500 $TR bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2f_bench_n8.log 2>&1; echo "n8 rc=$?"; tail -1 gpurun_out/r2f_bench_n8.log | cut -c1-700
timeout 500 $TR bench.py --gpus 8 --workload gen128 --steps 2 --warmup 1 > gpurun_out/r2f_gen128_n8.log 2>&1; echo "gen128 n8 rc=$?"; tail -1 gpurun_out/r2f_gen128_n8.log | cut -c1-700
timeout 700 $TR bench.py --gpus 8 --workload beam4 --steps 1 --warmup 1 > gpurun_out/r2f_beam4_n8.log 2>&1; echo "beam4 n8 rc=$?"; tail -1 gpurun_out/r2f_beam4_n8.log | cut -c1-700
