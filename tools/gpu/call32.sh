#!/bin/bash
mkdir -p gpurun_out/c32
O=gpurun_out/c32
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "wgrad or 13 or replay" > $O/t_wgrad.log 2>&1; echo "exit $?" >> $O/t_wgrad.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py > $O/time_bf16.log 2>&1
RESR_PREC=bf16 timeout 300 python tools/trace_train.py > $O/trace_bf16.log 2>&1
RESR_TRAIN_ONE_STREAM=1 RESR_PREC=bf16 timeout 300 python tools/trace_train.py > $O/trace_bf16_one.log 2>&1
tail -n 3 $O/t_wgrad.log; tail -n 1 $O/time_bf16.log; sed -n 3,6p $O/trace_bf16.log; grep "n=" $O/trace_bf16.log | head -8; sed -n 3,6p $O/trace_bf16_one.log;  grep "n=" $O/trace_bf16_one.log | head -8; grep "forward span" $O/trace_bf16*.log
