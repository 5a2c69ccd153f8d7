#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_degrade_gpu.py -m gpu -q -s > gpurun_out/degrade.log 2>&1; echo "rc=$?" >> gpurun_out/degrade.log
tail -60 gpurun_out/degrade.log
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/time_generator.py 64 128 128 2>&1 | tail -1
