#!/bin/bash
mkdir -p gpurun_out/c37
O=gpurun_out/c37
for d in 20 532 0 512; do RESR_WGRAD_MN_DEBUG=$d timeout 120 python tools/time_wgrad_mn.py 2>&1 | grep "per-kernel\|debug=" >> $O/probe.log; done
cat $O/probe.log
timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "wgrad" > $O/t_wgrad.log 2>&1; tail -n 2 $O/t_wgrad.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
