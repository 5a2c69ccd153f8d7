#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/time_generator.py 64 128 128 2>&1 | tail -4
timeout 300 python tools/time_train.py 2>&1 | tail -4
