#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py -m gpu -x -q -s 2>&1 | grep -i "max-abs\|psnr\|passed\|failed\|error\|cos" | tail -20
bash tools/ab.sh v6 v8
timeout 300 python tools/time_train.py 2>&1 | tail -1
