#!/bin/bash
mkdir -p gpurun_out/c19
O=gpurun_out/c19
for v in 0 1 0 1; do
  echo "ONE_STREAM=$v" >> $O/train_ab.log
  if [ $v = 1 ]; then export RESR_TRAIN_ONE_STREAM=1; else unset RESR_TRAIN_ONE_STREAM; fi
  timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
" >> $O/train_ab.log 2>&1
done
RESR_TRAIN_ONE_STREAM=1 python tools/trace_train.py > $O/trace_one.txt 2>&1
cat $O/train_ab.log; sed -n 5,30p $O/trace_one.txt | cut -c1-150; tail -n 18 $O/trace_one.txt | cut -c1-150
