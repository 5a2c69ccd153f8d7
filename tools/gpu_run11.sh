#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -m gpu -x -q > gpurun_out/conv.log 2>&1; echo "rc=$?" >> gpurun_out/conv.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/conv.log | head -20
timeout 300 python tools/conv_phases.py 2>&1 | tail -4 | cut -c1-190
timeout 300 python tools/time_generator.py 64 128 128 2>&1 | tail -1
