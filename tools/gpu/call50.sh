#!/bin/bash
mkdir -p gpurun_out/c50
O=gpurun_out/c50
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "oracle_autograd or 13 or replay or buckets or autograd_function" > $O/t_train.log 2>&1; echo "exit $?" >> $O/t_train.log
grep "loss rel err\|passed\|failed\|exit" $O/t_train.log | tail -n 22
RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
RESR_PREC=fp16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
