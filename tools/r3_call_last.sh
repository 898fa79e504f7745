#!/bin/bash
# last call of the round: chain-path parity after the 2-stage change, final lines, tensor-pipe evidence for configs[4]
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_shapes_gpu.py tests/test_decode_paths_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed" | tee gpurun_out/r3r_pytest.log
timeout 400 python bench.py --workload gen128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3r_gen128.log 2>&1; tail -1 gpurun_out/r3r_gen128.log | cut -c1-160
timeout 600 python bench.py > gpurun_out/r3r_bench.log 2>&1; tail -1 gpurun_out/r3r_bench.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 320 --launch-count 6 \
  -f -o gpurun_out/r3r_beam_gemm python tools/profile_run.py --batch 125 --num-beams 4 --text-len 256 --max-length 8 > gpurun_out/r3r_ncu_g.log 2>&1; echo "ncu gemm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:beam_cross_attn24_kernel --launch-skip 30 --launch-count 1 \
  -f -o gpurun_out/r3r_beam_cross python tools/profile_run.py --batch 125 --num-beams 4 --text-len 256 --max-length 8 > gpurun_out/r3r_ncu_c.log 2>&1; echo "ncu beam cross rc=$?"
