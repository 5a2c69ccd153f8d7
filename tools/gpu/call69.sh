#!/bin/bash
mkdir -p gpurun_out/c69
O=gpurun_out/c69
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_generator_gpu.py tests/test_conv_gpu.py -x -q -m gpu > $O/t.log 2>&1; echo "exit $?" >> $O/t.log
tail -n 3 $O/t.log
timeout 600 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 --no-other-precision 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'train', d['training']['ms_per_step'], d['training']['value'], d['training']['optimizer'])
"
