#!/bin/bash
# 2-GPU check after the kernel changes of the second half of round 2: dist parity + one bench line
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/dist_check.py > gpurun_out/r3_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "^rank" gpurun_out/r3_dist_check.log
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r3_bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -1 gpurun_out/r3_bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases']['decode_step_ms_mean'], d['phases']['decode_step_ms_p50'], d['config']['parallelism'])"
