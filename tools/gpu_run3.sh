#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 357 -c 5 -o gpurun_out/prof_rdb -f python tools/time_generator.py 64 128 128 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
