#!/bin/bash
mkdir -p gpurun_out/c68
O=gpurun_out/c68
timeout 2400 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "exit $?" >> $O/t_all.log
tail -n 3 $O/t_all.log
RESR_PREC=bf16 timeout 200 python tools/time_train.py 2>&1 | tail -n 1
timeout 600 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 --no-other-precision 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'train', d['training']['ms_per_step'], d['training']['value'])
"
