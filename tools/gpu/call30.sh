#!/bin/bash
mkdir -p gpurun_out/c30
O=gpurun_out/c30
for v in "0 0" "0 72" "0 36" "100 0" "100 48" "74 72" "0 0"; do set -- $v
  echo "CONV_SMS=$1 WGRAD_SMS=$2" >> $O/ab.log
  RESR_TRAIN_CONV_SMS=$1 RESR_WGRAD_SMS=$2 timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
" >> $O/ab.log 2>&1
done
cat $O/ab.log
