#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q -s > gpurun_out/train.log 2>&1; echo "rc=$?" >> gpurun_out/train.log
grep -E "^E  |passed|failed|rc=|loss rel|Error|\[resr\]" gpurun_out/train.log | head -30
