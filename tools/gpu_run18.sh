#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q > gpurun_out/train.log 2>&1; echo "rc=$?" >> gpurun_out/train.log
grep -E "^E  .*(assert|Error)|passed|failed|rc=" gpurun_out/train.log | head -10
timeout 300 python tools/time_train.py 16 64 64 2>&1 | tail -1
RESR_GRAPH=0 timeout 300 python tools/time_train.py 16 64 64 2>&1 | tail -1
timeout 300 python tools/time_train.py 8 128 128 2>&1 | tail -1
