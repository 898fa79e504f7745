#!/bin/bash
# two lanes sharing every SM (2 CTAs / SM, ring 2 x 40 KB each): feasibility
mkdir -p gpurun_out
V=$PWD/markushgrapher_b200/lib/libmg_b200_occ2.so
MG_B200_LIB=$V timeout 400 python tools/ab_env.py --settings "" > gpurun_out/r2n_occ2_single32.log 2>&1
MG_B200_LIB=$V timeout 500 python tools/ab_two_lanes.py > gpurun_out/r2n_occ2_two_lanes.log 2>&1
MG_B200_LIB=$V MG_MEGA_NOCOOP=1 timeout 500 python tools/ab_two_lanes.py > gpurun_out/r2n_occ2_two_lanes_nocoop.log 2>&1
tail -n 12 gpurun_out/r2n_occ2_single32.log gpurun_out/r2n_occ2_two_lanes.log gpurun_out/r2n_occ2_two_lanes_nocoop.log
