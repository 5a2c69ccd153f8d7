#!/bin/bash
mkdir -p gpurun_out/c72
O=gpurun_out/c72
for v in "X=0" "RESR_CONV_PAIR_SMALL64=3" "RESR_CONV_PAIR_SMALL64=1" "RESR_CONV_PAIR=2" "X=0"; do
  echo "$v" >> $O/ab.log
  env $v RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1 >> $O/ab.log
done
cat $O/ab.log
