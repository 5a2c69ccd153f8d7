#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -m gpu -x -q > gpurun_out/conv.log 2>&1; echo "rc=$?" >> gpurun_out/conv.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/conv.log | head -20
timeout 300 python tools/time_generator.py 64 128 128 2>&1 | tail -2
timeout 300 python tools/time_generator.py 16 64 64 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 352 --csv --log-file gpurun_out/launches_gen3.csv python tools/time_generator.py 64 128 128 > /dev/null 2>&1
timeout 300 python tools/time_degrade.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_deg.csv python tools/time_degrade.py > /dev/null 2>&1
