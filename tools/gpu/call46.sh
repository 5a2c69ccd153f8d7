#!/bin/bash
mkdir -p gpurun_out/c46
O=gpurun_out/c46
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "wgrad or 13 or replay or oracle_autograd" > $O/t_train.log 2>&1; tail -n 2 $O/t_train.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
RESR_PREC=bf16 timeout 300 python tools/trace_train.py > $O/trace_bf16.log 2>&1
grep "n=" $O/trace_bf16.log | head -8; grep "forward span" $O/trace_bf16.log
