#!/bin/bash
mkdir -p gpurun_out/c73
O=gpurun_out/c73
timeout 200 python tools/time_train_split.py 1 2>&1 | tail -n 1 >> $O/ab.log
RESR_NUM_SMS=74 timeout 200 python tools/time_train_split.py 2 2>&1 | tail -n 1 >> $O/ab.log
timeout 200 python tools/time_train_split.py 2 2>&1 | tail -n 1 >> $O/ab.log
RESR_NUM_SMS=74 timeout 200 python tools/time_train_split.py 1 2>&1 | tail -n 1 >> $O/ab.log
RESR_NUM_SMS=48 timeout 200 python tools/time_train_split.py 3 2>&1 | tail -n 1 >> $O/ab.log
cat $O/ab.log
