#!/bin/bash
# round-2 2-GPU call: dist parity (peer stores / NCCL / beams / early EOS / handshake), per-step cost of the two exchange paths
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/dist_check.py > gpurun_out/r2e_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "^rank" gpurun_out/r2e_dist_check.log
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_bench_n2_p2p.log 2>&1; echo "bench p2p rc=$?"; tail -1 gpurun_out/r2e_bench_n2_p2p.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases']['decode_step_ms_mean'], d['phases']['decode_step_ms_p50'], d['config']['parallelism'])"
MG_DIST=nccl timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_bench_n2_nccl.log 2>&1; echo "bench nccl rc=$?"; tail -1 gpurun_out/r2e_bench_n2_nccl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases']['decode_step_ms_mean'], d['phases']['decode_step_ms_p50'], d['config']['parallelism'])"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_n1.log 2>&1; tail -1 gpurun_out/r2e_bench_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases']['decode_step_ms_mean'], d['phases']['decode_step_ms_p50'])"
