#!/bin/bash
mkdir -p gpurun_out/c29
O=gpurun_out/c29
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py tests/test_train_gpu.py tests/test_parity_at_size_gpu.py -x -q -m gpu -k "not cfg2" > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
" > $O/train.log 2>&1
tail -n 3 $O/tests.log; cat $O/train.log
