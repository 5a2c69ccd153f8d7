#!/bin/bash
mkdir -p gpurun_out/c15
O=gpurun_out/c15
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py tests/test_degrade_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 600 python -m pytest tests/test_parity_at_size_gpu.py -x -q -m gpu -k "cfg3 or cfg5" -s 2>&1 | grep -E "cfg|passed|failed" > $O/atsize.log
for v in 0 1 0 1; do
  echo "SERPENTINE=$v" >> $O/ab.log
  RESR_CONV_SERPENTINE=$v timeout 300 python bench.py --no-train --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> $O/ab.log 2>&1
done
for v in 0 1; do
  echo "ONE_STREAM=$v" >> $O/train_ab.log
  if [ $v = 1 ]; then export RESR_TRAIN_ONE_STREAM=1; else unset RESR_TRAIN_ONE_STREAM; fi
  timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
" >> $O/train_ab.log 2>&1
done
unset RESR_TRAIN_ONE_STREAM
tail -n 4 $O/tests.log; cat $O/atsize.log $O/ab.log $O/train_ab.log
