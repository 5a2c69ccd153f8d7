#!/bin/bash
mkdir -p gpurun_out/c31
O=gpurun_out/c31
timeout 300 python tools/dbg_wgrad_mn.py > $O/dbg.log 2>&1; echo "exit $?" >> $O/dbg.log
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "wgrad" > $O/t_wgrad.log 2>&1; echo "exit $?" >> $O/t_wgrad.log
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s -k "not wgrad" > $O/t_train.log 2>&1; echo "exit $?" >> $O/t_train.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py > $O/time_bf16.log 2>&1
RESR_PREC=fp16 timeout 200 python tools/time_train.py > $O/time_fp16.log 2>&1
RESR_PREC=bf16 timeout 300 python tools/trace_train.py > $O/trace_bf16.log 2>&1
tail -n 12 $O/dbg.log; tail -n 3 $O/t_wgrad.log; tail -n 5 $O/t_train.log; tail -n 1 $O/time_bf16.log $O/time_fp16.log; sed -n 3,6p $O/trace_bf16.log; grep "n=" $O/trace_bf16.log | head -12
