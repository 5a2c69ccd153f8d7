#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/time_generator.py 64 128 128 > gpurun_out/time.log 2>&1
RESR_CONV_MODE=1 timeout 300 python tools/time_generator.py 64 128 128 >> gpurun_out/time.log 2>&1
timeout 300 python tools/time_generator.py 16 64 64 >> gpurun_out/time.log 2>&1
timeout 300 python tools/time_generator.py 1 128 128 >> gpurun_out/time.log 2>&1
cat gpurun_out/time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 352 --csv --log-file gpurun_out/launches_gen.csv python tools/time_generator.py 64 128 128 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
