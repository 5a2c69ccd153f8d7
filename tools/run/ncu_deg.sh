#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"usm_fused|filter2d" -s 4 -c 3 -o gpurun_out/prof_deg -f python tools/time_degrade.py > gpurun_out/ncu_deg.log 2>&1
tail -1 gpurun_out/ncu_deg.log
