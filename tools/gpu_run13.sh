#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_degrade_gpu.py -m gpu -x -q > gpurun_out/deg.log 2>&1; echo "rc=$?" >> gpurun_out/deg.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/deg.log | head -20
timeout 300 python tools/time_degrade.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_deg2.csv python tools/time_degrade.py > /dev/null 2>&1
