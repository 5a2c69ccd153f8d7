#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do timeout 100 python tools/dbg_epi.py 2>&1 | grep "bad count" | tr '\n' ' '; done; echo
bash tools/ab.sh v6 v7
timeout 300 python tools/time_train.py 2>&1 | tail -1
