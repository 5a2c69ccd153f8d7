#!/bin/bash
# first GPU bring-up: conv parity, generator parity, quick timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > gpurun_out/conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/conv.log
tail -25 gpurun_out/conv.log
timeout 600 python -m pytest tests/test_generator_gpu.py -m gpu -q -s > gpurun_out/gen.log 2>&1; echo "gen rc=$?" >> gpurun_out/gen.log
tail -25 gpurun_out/gen.log
timeout 300 python tools/time_generator.py 64 128 128 > gpurun_out/time.log 2>&1
RESR_CONV_MODE=1 timeout 300 python tools/time_generator.py 64 128 128 >> gpurun_out/time.log 2>&1
timeout 300 python tools/time_generator.py 16 64 64 >> gpurun_out/time.log 2>&1
cat gpurun_out/time.log
