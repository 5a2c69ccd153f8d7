#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all.log 2>&1; echo "rc=$?" >> gpurun_out/all.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/all.log | head -10
timeout 300 python tools/time_train.py 16 64 64 2>&1 | tail -1
timeout 300 python tools/time_generator.py 64 128 128 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/time_train.py 16 64 64 > /dev/null 2>&1
