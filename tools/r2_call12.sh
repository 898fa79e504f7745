#!/bin/bash
# two-lane feasibility (tools/ab_two_lanes.py) against the one-lane default on the same box
mkdir -p gpurun_out
timeout 400 python tools/ab_env.py --settings "" > gpurun_out/r2m_single.log 2>&1
MG_MEGA_CTAS=74 timeout 500 python tools/ab_two_lanes.py > gpurun_out/r2m_two_lanes.log 2>&1
MG_MEGA_CTAS=74 MG_MEGA_NOCOOP=1 timeout 500 python tools/ab_two_lanes.py --stagger-ms 1.1 > gpurun_out/r2m_two_lanes_nocoop.log 2>&1
MG_MEGA_CTAS=74 MG_MEGA_INFLIGHT=3 timeout 500 python tools/ab_two_lanes.py > gpurun_out/r2m_two_lanes_if3.log 2>&1
tail -n 12 gpurun_out/r2m_single.log gpurun_out/r2m_two_lanes.log gpurun_out/r2m_two_lanes_nocoop.log gpurun_out/r2m_two_lanes_if3.log
