#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh v3 v6
