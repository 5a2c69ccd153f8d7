#!/bin/bash
mkdir -p gpurun_out/c42
O=gpurun_out/c42
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "wgrad or 13 or replay" > $O/t_wgrad.log 2>&1; tail -n 2 $O/t_wgrad.log
for n in 4 16 64; do N=$n timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel" >> $O/probe.log; done
cat $O/probe.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
RESR_PREC=bf16 timeout 300 python tools/trace_train.py > $O/trace_bf16.log 2>&1
grep "n=" $O/trace_bf16.log | head -6; grep "forward span" $O/trace_bf16.log
